"""Tensor-pipe throughput of the streamed tcgen05 conv kernels (csrc/conv_tc_big.cu) on the tensor-bound convs of
the stacks around the conv-GAT layer: the DCGAN discriminator convs after 2x2 regrouping (dcgan/model.py:152-160,
ndf = 64) and SmaAt-UNet-like pointwise convs.  CUDA events around `reps` back-to-back launches through the C ABI;
prints one JSON line per (shape, direction) with achieved TFLOP/s and the fraction of MEASURED_PEAKS.json's dense
bf16 figure (builder's tool, not the bench contract)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from cgat import _lib
from cgat.functional import _conv_desc, ptr, stream

dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = 20
only = sys.argv[2] if len(sys.argv) > 2 else ""  # substring filter on the shape name
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1369.9
SHAPES = [
    # name, n, h, w, cin, cout, k, pad
    ("dcgan conv2 (64->128 k4s2 as 2x2 over 256)", N, 17, 17, 256, 128, 2, 0),
    ("dcgan conv3 (128->256 k4s2 as 2x2 over 512)", N, 9, 9, 512, 256, 2, 0),
    ("dcgan conv4 (256->512 k4s2 as 2x2 over 1024)", N, 5, 5, 1024, 512, 2, 0),
    ("unet pointwise 256->256 @32x32", N, 32, 32, 256, 256, 1, 0),
    ("unet pointwise 1024->512 @16x16", N, 16, 16, 1024, 512, 1, 0),
    ("dense 3x3 256->256 @32x32", N, 32, 32, 256, 256, 3, 1),
]
lib = _lib.lib()
for name, n, h, w, cin, cout, k, pad in SHAPES:
    if only not in name:
        continue
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    d = _conv_desc(n, h, w, cin, cout, k, k, 1, pad, pad, ho, wo, _lib.BF16, 0)
    x = (torch.rand(n, h, w, cin, device=dev) - 0.5).bfloat16()
    wt = (torch.rand(cout, k, k, cin, device=dev) - 0.5).bfloat16()
    y = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.bfloat16)
    dy = (torch.rand(n, ho, wo, cout, device=dev) - 0.5).bfloat16()
    dx = torch.empty_like(x)
    dw = torch.empty(cout, k, k, cin, device=dev, dtype=torch.float32)
    flops = 2.0 * n * ho * wo * cout * cin * k * k
    ws = [torch.empty(max(16, lib.cgat_conv_workspace_bytes(ctypes.byref(d), i)), dtype=torch.uint8, device=dev) for i in range(3)]
    calls = {
        "fprop": lambda: _lib.call("cgat_conv2d_fprop", ctypes.byref(d), ptr(x), ptr(wt), None, ptr(y), 1, ptr(ws[0]), stream()),
        "dgrad": lambda: _lib.call("cgat_conv2d_dgrad", ctypes.byref(d), ptr(dy), ptr(wt), ptr(dx), 1, ptr(ws[1]), stream()),
        "wgrad": lambda: _lib.call("cgat_conv2d_wgrad", ctypes.byref(d), ptr(x), ptr(dy), ptr(dw), None, 1, ptr(ws[2]), stream()),
    }
    for which, fn in calls.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # the launches are replayed from a CUDA graph: the host side of a call (two cuTensorMapEncodeTiled + ctypes)
        # costs ~10 us, more than the small shapes run for
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for _ in range(reps):
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        tf = flops / ms / 1e9
        print(json.dumps({"shape": name, "n": n, "dir": which, "ms": round(ms, 4), "tflops": round(tf, 1),
                          "frac_of_peak": round(tf / peak, 3), "peak_tflops": peak}), flush=True)
