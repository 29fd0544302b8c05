"""Developer aid: clock64() timeline of CTA 0 of the fused layer kernel (train mode) at the bench size."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import subprocess
import torch
# the stamps are compiled in only with TIMELINE=1: build that variant of the library next to the product one
CS = os.path.join(ROOT, "extended-gan_b200", "csrc")
subprocess.check_call(f"cd {CS} && mkdir -p build_tl && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo "
                      f"-Xcompiler -fPIC --use_fast_math -DCGAT_LF_TIMELINE {os.environ.get('CGAT_TL_DEFS', '')} -c layer_fused.cu -o build_tl/layer_fused.o && "
                      f"nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcgat_b200_timeline.so "
                      f"$(ls build/*.o | grep -v layer_fused.o) build_tl/layer_fused.o -lcuda", shell=True)
os.environ["CGAT_B200_LIB"] = os.path.join(ROOT, "extended-gan_b200", "libcgat_b200_timeline.so")
from cgat import _lib
from cgat.train_step import TrainStep
from convolutional_gat.GAT3D.GATMultistream import Model

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
torch.manual_seed(369)
dev = "cuda"
m = Model(image_width=64, image_height=64, n_vertices=6, attention_type="temporal", mapping_type="conv").to(dev)
x = torch.rand(64, 64, 64, 4, 6, device=dev).bfloat16()
y = torch.rand(64, 64, 64, 4, 6, device=dev).bfloat16()
ts = TrainStep(m, x, y, use_graph=False, fuse_loss=(mode == "train"))
for _ in range(3):
    ts.run()
buf = torch.zeros(16 * 16 + 2 * 148, dtype=torch.int64, device=dev)
full_buf = buf
buf = full_buf[:256].view(16, 16)
if len(sys.argv) > 2:
    buf[15, 15] = 1  # probe mode: the MMA thread waits for its own MMAs and logs their completion
L = _lib.lib()
L.cgat_layer_debug_timeline.argtypes = [ctypes.c_void_p]
L.cgat_layer_debug_timeline.restype = None
L.cgat_layer_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
if mode == "fwd":
    with torch.no_grad():
        m(x)
else:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ts.run()
    e1.record()
torch.cuda.synchronize()
L.cgat_layer_debug_timeline(None)
t = buf.cpu()
t0 = int(t[0][t[0] > 0].min())
names = ["Ttop", "Tempty", "Mfpdone", "Mwgdone", "Mfull", "Maccfr", "Mfprop", "Mdyfull", "Mwgrad", "Awh", "Aregs", "Afwd",
         "Aexch", "Abwd", "Adone", "g0bar1"]
print("tile " + " ".join(f"{n:>8s}" for n in names))
for i in range(14):
    print(f"{i:4d} " + " ".join(f"{(int(v) - t0) if v > 0 else -1:8d}" for v in t[i][:16]))
x = t[14]
print("kernel stamps (cycles from entry): " + " ".join(f"{n}={int(x[i]) - int(x[0])}" for i, n in enumerate(
    ["entry", "init", "regs", "tiles_done", "flushed", "wgrad_done", "partials", "joined", "exit", "weights_landed", "x0_landed"])),
      "first TMA at", t0 - int(x[0]))
c = full_buf[256:].view(148, 2).cpu()
if mode != "fwd":
    g0 = int(c[:, 0].min())
    st, en = (c[:, 0] - g0).double() / 1e3, (c[:, 1] - g0).double() / 1e3
    print(f"per-CTA global timer (us): start min {st.min():.2f} median {st.median():.2f} max {st.max():.2f} | end min {en.min():.2f} "
          f"median {en.median():.2f} max {en.max():.2f} | duration min {(en - st).min():.2f} median {(en - st).median():.2f} max {(en - st).max():.2f}")
    print(f"whole eager step by CUDA events: {e0.elapsed_time(e1) * 1e3:.1f} us")
