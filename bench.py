#!/usr/bin/env python
"""bench.py -- conv-GAT train-step throughput on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full train step of the reference loop (convolutional_gat/train.py:129-133: forward, loss,
backward, Adam) of the conv-GAT model on one synthetic batch of the data loaders' shape
``x, y : [64, 64, 64, 4, 6]`` bf16 per GPU (BASELINE.json configs[1]; weak scaling for N > 1).  Prints ONE JSON
line on rank 0.  ``value`` is device-resident throughput; ``e2e`` includes the per-step host->device copy of
x, y from pinned memory and the device->host read of the loss.  L2 is flushed between timed steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "extended-gan_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "convgat_train_samples_per_sec"
UNIT = "samples/s"
SHAPE = (64, 64, 4, 6)  # H, W, T, V: 64x64 coastal-sea crop (data_loader.py:14), 4 frames (:16), 6 KNMI regions
SEED = 369  # the reference's only seed (dcgan/train.py:181-183)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_batch(n, dtype, device="cpu", pin=False):
    g = torch.Generator().manual_seed(SEED)
    x = torch.rand((n,) + SHAPE, generator=g).to(dtype)
    y = torch.rand((n,) + SHAPE, generator=g).to(dtype)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return x.to(device), y.to(device)


# ----------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference arithmetic on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(sample_n: int, steps: int, warmup: int, attention_type: str, mapping: str):
    from oracle import spec

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(SEED)
    model = spec.SpecGATMultiHead3D(4, 4, 0.2, 3, type_=attention_type, mapping_type=mapping, n_vertices=SHAPE[3])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.01)  # train.py:212
    x, y = synth_batch(sample_n, torch.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = model(x)
        loss = spec.train_loss(out, y)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return sample_n / sec, sec, cores


def conv_tensor_roofline(pk, batch=64, reps=10):
    """Tensor-pipe side of the roofline: the streamed tcgen05 implicit-GEMM conv kernels (csrc/conv_tc_big.cu) that serve
    the dense convs of the stacks around the layer (DCGAN discriminators dcgan/model.py:152-160 after 2x2 regrouping;
    SmaAt-UNet pointwise convs), timed live: `reps` launches through the C ABI replayed from one CUDA graph, CUDA
    events around the replay.  achieved = 2*pixels*cin*cout*taps / time against MEASURED_PEAKS' burst bf16 figure (kernels timed alone)."""
    import ctypes
    import torch
    from cgat import _lib
    from cgat.functional import _conv_desc, ptr, stream

    dev = "cuda"
    shapes = [  # name, n, h, w, cin, cout, k, pad
        ("dense 3x3 256->256 @32x32", batch, 32, 32, 256, 256, 3, 1),
        ("dcgan conv3 128->256 k4s2 (2x2 over 512 regrouped channels) @8x8", batch, 9, 9, 512, 256, 2, 0),
        ("dcgan conv4 256->512 k4s2 (2x2 over 1024) @4x4", batch, 5, 5, 1024, 512, 2, 0),
        ("unet pointwise 1024->512 @16x16", batch, 16, 16, 1024, 512, 1, 0),
    ]
    lib = _lib.lib()
    out = []
    for name, n, h, w, cin, cout, k, pad in shapes:
        ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
        d = _conv_desc(n, h, w, cin, cout, k, k, 1, pad, pad, ho, wo, _lib.BF16, 0)
        x = (torch.rand(n, h, w, cin, device=dev) - 0.5).bfloat16()
        wt = (torch.rand(cout, k, k, cin, device=dev) - 0.5).bfloat16()
        y = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.bfloat16)
        dy = (torch.rand(n, ho, wo, cout, device=dev) - 0.5).bfloat16()
        dx = torch.empty_like(x)
        dw = torch.empty(cout, k, k, cin, device=dev, dtype=torch.float32)
        ws = [torch.empty(max(16, lib.cgat_conv_workspace_bytes(ctypes.byref(d), i)), dtype=torch.uint8, device=dev)
              for i in range(3)]
        calls = {
            "fprop": lambda: _lib.call("cgat_conv2d_fprop", ctypes.byref(d), ptr(x), ptr(wt), None, ptr(y), 1, ptr(ws[0]), stream()),
            "dgrad": lambda: _lib.call("cgat_conv2d_dgrad", ctypes.byref(d), ptr(dy), ptr(wt), ptr(dx), 1, ptr(ws[1]), stream()),
            "wgrad": lambda: _lib.call("cgat_conv2d_wgrad", ctypes.byref(d), ptr(x), ptr(dy), ptr(dw), None, 1, ptr(ws[2]), stream()),
        }
        flops = 2.0 * n * ho * wo * cout * cin * k * k
        for which, fn in calls.items():
            fn()
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side), torch.cuda.graph(graph, stream=side):
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
            tf = flops / (ms * 1e-3) / 1e12
            out.append({"shape": name, "n": n, "dir": which, "ms": ms, "tflops": tf, "frac": tf / pk["bf16_tflops"]})
    # headline entry: the best REFERENCE shape (the synthetic 3x3 256->256 conv, which no reference model contains, shows
    # what the kernel reaches when the problem fills the machine and is reported beside it)
    ref = [o for o in out if not o["shape"].startswith("dense 3x3")]
    best = max(ref, key=lambda o: o["tflops"])
    synth = out[0]
    return {"kernel": "conv_big_fprop_kernel", "bound": "tensor", "achieved": best["tflops"],
            "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": best["frac"], "shape": best["shape"], "dir": best["dir"],
            "frac_of_sustained_peak": best["tflops"] / pk["bf16_tflops_sustained"],
            "peak_source": pk["source"] + " (burst figure: the kernels are timed alone)", "timing": f"{reps} launches replayed from one CUDA graph, CUDA events",
            "synthetic_full_machine_shape": {"shape": synth["shape"], "dir": synth["dir"], "tflops": synth["tflops"],
                                             "frac": synth["frac"]},
            "note": "the reference's dense convs are small problems at its batch (DCGAN conv3 / conv4: 32 / 16 output tiles for "
                    "148 SMs), bound by occupancy of the machine, not by the tensor pipe of the SMs that work",
            "all": out}


def ours_rows(dev, batch=64):
    """Our modules on the workloads of the eager-GPU baseline rows R1, R2, R4 (R3 is the headline line itself)."""
    from oracle.baselines import fwd_bwd, time_fn  # timing helpers only
    from convolutional_gat.baseline_model import BaselineModel, BaselineModel2D
    from dcgan.model import FrameDiscriminator, Generator, TemporalDiscriminator
    from dcgan.train import GraphedAdversarialStep, default_criterion, make_optimizers

    torch.manual_seed(SEED)
    rows = {}
    x20 = torch.rand(4, 20, 20, 4, 6, device=dev)
    for name, cls in (("R1_BaselineModel2D_20x20_fwd_bwd", BaselineModel2D), ("R2_BaselineModel_20x20_fwd_bwd", BaselineModel)):
        m = cls(image_width=20, image_height=20, n_vertices=6).to(dev)
        sec = time_fn(fwd_bwd(m, x20), dev, 3, 10)
        rows[name] = {"samples_per_s": 4 / sec, "ms_per_step": sec * 1e3, "batch": 4, "dtype": "f32"}
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        params = {"nc": 4, "ndf": 64}
        nets = [Generator(params).to(dev), FrameDiscriminator(params).to(dev), TemporalDiscriminator(params).to(dev)]
        oG, oFD, oTD = make_optimizers(*nets, capturable=True)
        x = torch.rand(batch, 4, 64, 64, device=dev).to(dtype)
        y = torch.rand(batch, 4, 64, 64, device=dev).to(dtype)
        step = GraphedAdversarialStep(netG=nets[0], netFD=nets[1], netTD=nets[2], optimizerG=oG, optimizerFD=oFD,
                                      optimizerTD=oTD, criterion=default_criterion(), x=x, y=y)
        sec = time_fn(lambda: step(x, y), dev, 2, 10)
        rows[f"R4_dcgan_adversarial_step_{tag}"] = {"samples_per_s": batch / sec, "ms_per_step": sec * 1e3, "batch": batch,
                                                    "dtype": tag, "cuda_graph": True}
    # BASELINE config 4 (stress): UnetModel [2,128,128,4,8], one full train step from TrainStep's CUDA graph -- the V
    # per-vertex UNet passes as one batched pass with per-vertex BatchNorm statistic sets; fp32 master parameters
    from cgat.train_step import TrainStep
    from convolutional_gat.unet_model import UnetModel
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16_activations")):
        torch.manual_seed(SEED)
        m = UnetModel(image_width=128, image_height=128, n_vertices=8, attention_type="unet").to(dev)
        xu = torch.rand(2, 128, 128, 4, 8, device=dev).to(dtype)
        yu = torch.rand(2, 128, 128, 4, 8, device=dev).to(dtype)
        ts = TrainStep(m, xu, yu, lr=1e-3, use_graph=True)
        sec = time_fn(lambda: ts.run(), dev, 2, 10)
        rows[f"config4_unet_model_train_step_{tag}"] = {"samples_per_s": 2 / sec, "ms_per_step": sec * 1e3, "batch": 2,
                                                        "dtype": tag, "cuda_graph": True}
        del ts, m
    return rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 4  # BASELINE.json configs[0]: small batch on CPU, fp32
    rate, sec, cores = cpu_reference_step_rate(n, max(1, args.steps), max(1, min(args.warmup, 2)), args.type, args.mapping)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_gpu_batch=n),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} samples/step of the same workload (fp32, torch CPU, oracle/spec.py restatement; "
                                   "the reference's own GAT3D layer is absent from its tree)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, per_gpu_batch):
    H, W, T, V = SHAPE
    return {
        "workload": f"convolutional_gat train step: GATMultistream.Model(attention_type={args.type}, mapping_type={args.mapping}), "
                    f"3 heads, x,y [{per_gpu_batch},{H},{W},{T},{V}] per GPU (BASELINE.json configs[1])",
        "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * args.gpus, "image": [H, W], "time_steps": T,
        "n_vertices": V, "parallelism": f"dp{args.gpus}",
        "l2": ("inputs larger than L2: the steps rotate over 8 input slots of x, y (8 x 25.2 MB = 201 MB > 126 MB L2), K steps "
               "inside ONE event pair, no flush; x resident chunk-planar [N,T*V/8,H,W,8] as the loader kernel writes it, "
               "y [N,H,W,T,V]" if args.l2 == "rotate" else
               "flushed between timed steps (256 MiB memset outside the per-step event pairs)"),
        "optimizer": "Adam(lr=1e-3, weight_decay=0.01) fused, flat fp32 buffers", "cuda_graph": True,
        "timing": "W warm-up steps, barrier + synchronize, K steps inside one CUDA-event pair, barrier + synchronize, max over ranks"
                  + ("; N > 1: a device-side all-reduce is enqueued between the barrier and the start event so that the ranks' "
                     "timed regions start together" if args.gpus > 1 else ""),
    }


# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE is 1)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at any debug level >= VERSION; main() has pointed fd 1 at stderr, so
        # whatever native libraries print, rank 0's stdout carries ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from cgat import _lib
    from cgat.train_step import TrainStep
    from convolutional_gat.GAT3D.GATMultistream import Model

    H, W, T, V = SHAPE
    B = args.batch
    dtype = torch.bfloat16
    torch.manual_seed(SEED)
    model = Model(image_width=W, image_height=H, n_vertices=V, attention_type=args.type, mapping_type=args.mapping).to(dev)
    # synthetic batch in the LOADER'S raw format (kmni_data_loader.py:72-127): B overlapping 8-frame windows of
    # B + 7 raw uint8 frames [L, V, H, W]; x, y are what the loader derives from them (every rank: its own shard)
    from convolutional_gat.data_loaders.kmni_data_loader import gather_windows
    g = torch.Generator().manual_seed(SEED)
    frames_h = torch.randint(0, 255, (B + 2 * T - 1, V, H, W), generator=g, dtype=torch.uint8).pin_memory()
    start_h = torch.arange(B, dtype=torch.int32).pin_memory()
    x, y = gather_windows(frames_h.to(dev), start_h.to(dev), steps=T, dtype=dtype)
    xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
    launches0 = _lib.LAUNCHES
    ts = TrainStep(model, x, y, lr=1e-3, use_graph=True)
    ts.sync_params()
    p2p = ts.enable_p2p_exchange() if world > 1 else False  # gradient exchange + Adam in one peer-memory kernel
    # every step's loss reaches the host as a posted write of the step's last kernel into mapped pinned memory (a ring), not
    # as a 4-byte memcpy between two steps (--loss-readback memcpy keeps that)
    mirror = ts.fused_stream is not None and args.loss_readback == "mirror"
    if mirror:
        ts.enable_loss_mirror(4096)
    # launches of one step = those captured in the graph (one fwd+bwd pass) + Adam
    _lib.LAUNCHES = 0
    ts.graph = None
    ts._fwd_bwd()
    per_step_launches = _lib.LAUNCHES + 1  # (+1 loader gather launch per step in the e2e region)
    ts._capture()
    torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    align_buf = torch.zeros(1, device=dev) if world > 1 else None

    def align():
        """Device-side rendezvous enqueued (not synchronised) right before a start event: after barrier() the ranks' hosts
        leave the synchronisation tens of microseconds apart, which the max over ranks would count as step time of a
        20 x 0.07 ms region; the all-reduce kernel completes on every rank at the same moment."""
        if world > 1:
            dist.all_reduce(align_buf)

    def timed(step_fn, K, Wm):
        for _ in range(Wm):
            step_fn()
        barrier()
        evs = []
        for _ in range(K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident throughput ----
    if args.l2 == "rotate":
        # inputs larger than L2: eight slots of x, y (own buffers, own captured graph each; 25.2 MB of bf16 per slot); by
        # the time a slot comes round again 176 MB of other inputs have passed through the 126 MB L2.  One event pair around the K steps: the
        # ranks of a data-parallel run stay in step through the gradient exchange alone (per-step flushes end at
        # slightly different times on every rank, and the max over ranks then counts that jitter as step time).
        NS = 8
        ts.enable_prefetch(NS)
        for i, sl in enumerate(ts._slots[1:], 1):
            sl["x"].copy_(torch.roll(x, i, 0))
            sl["y"].copy_(torch.roll(y, i, 0))
            ts.refresh_planar(i)  # the resident input format of the train kernel (what the loader kernel writes)
        torch.cuda.synchronize()

        host_enqueue = [None]

        def timed_rotate(K, Wm):
            for i in range(Wm):
                ts.run_slot(i % NS)
            barrier()
            align()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            h0 = time.perf_counter()
            for i in range(K):
                ts.run_slot(i % NS)
            host_enqueue[0] = (time.perf_counter() - h0) / K * 1e3  # ms of host time per step to enqueue (no sync)
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        total_ms = timed_rotate(args.steps, max(3, args.warmup))
        # (where the exchange's time goes: tools/p2p_timeline.py -- the kernel is a node of the step's graph now)
    else:
        total_ms = timed(lambda: ts.run(), args.steps, max(3, args.warmup))
    # ---- end to end: every step copies its x, y from pinned host memory and reads the loss back to the host.
    # The copy of batch i+1 runs on a copy stream while batch i trains (two input slots, two captured graphs);
    # ONE event pair brackets all K steps, so every copy and every read is inside the timed region. ----
    loss_host = torch.empty(args.steps + 3, dtype=torch.float32).pin_memory()
    ts.enable_prefetch()

    pipelined = ts.fused_stream is not None
    if pipelined:  # raw frames: ONE graph launch per step (train step || H2D copies + gather of the next batch)
        ts.enable_raw_pipeline(frames_h, start_h)
    e2e_host = [None]

    def e2e_run(K, raw):
        if raw and pipelined:
            ts.prime_raw_pipeline()
            h0 = time.perf_counter()
            for i in range(K):
                loss = ts.run_pipelined(i & 1)
                if not mirror:
                    loss_host[i:i + 1].copy_(loss, non_blocking=True)
            e2e_host[0] = (time.perf_counter() - h0) / K * 1e3
            return
        pf = (lambda slot: ts.prefetch_raw(frames_h, start_h, slot)) if raw else (lambda slot: ts.prefetch(xh, yh, slot))
        pf(0)
        for i in range(K):
            if i + 1 < K:
                pf((i + 1) & 1)
            loss = ts.run_slot(i & 1)
            if not mirror:
                loss_host[i:i + 1].copy_(loss, non_blocking=True)

    def e2e_time(raw):
        e2e_run(3, raw)
        barrier()
        align()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        e2e_run(args.steps, raw)
        eb.record()
        barrier()
        t = torch.tensor([ea.elapsed_time(eb)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # headline: the loader's raw frames cross PCIe, windowing / normalisation / layout on the device; for comparison the
    # same steps with the finished x, y tensors (bf16) copied instead, as the reference's loader does in fp32
    e2e_xy_ms = e2e_time(False)
    e2e_ms = e2e_time(True)
    ts.graph = ts._slots[0]["graph"]
    ts.x, ts.y, ts.xp = ts._slots[0]["x"], ts._slots[0]["y"], ts._slots[0]["xp"]
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()
    final_loss = ts.loss_of_step(int(ts._mirror[1]) - 1) if mirror else float(loss_host[args.steps - 1].item())

    # ---- instrumented eager pass: CUDA-event time of every C-ABI kernel (same stream) ----
    ts.graph = None
    for _ in range(2):
        ts.run()
    torch.cuda.synchronize()
    _lib.profile_start()
    frames_d, start_d = frames_h.to(dev), start_h.to(dev)
    for _ in range(max(3, min(args.steps, 10))):
        flush.zero_()
        if ts.xp is not None:  # the e2e path's per-step loader kernel
            gather_windows(frames_d, start_d, steps=T, out=(ts.xp, ts.y), planar=True)
        else:
            gather_windows(frames_d, start_d, steps=T, out=(ts.x, ts.y))
        ts.run()
    torch.cuda.synchronize()
    prof = _lib.profile_stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    n_pix = B * H * W
    esz = 2
    heads = 3
    rec = T * V
    # ALGORITHMIC bytes / flops per launch (DESIGN.md section "roofline"): attention reads its input records and
    # writes its output records once; backward reads input + dout and writes din.
    pre = args.mapping == "conv"
    in_rec = heads * rec if pre else rec
    alg = {
        "cgat_attn_fwd": ("hbm", n_pix * (in_rec + rec) * esz),
        "cgat_attn_bwd": ("hbm", n_pix * (2 * in_rec + rec) * esz),
        "cgat_loss_fwd_bwd": ("hbm", n_pix * rec * 3 * esz),
        "cgat_loader_gather": ("hbm", n_pix * rec * 2 * esz + frames_h.numel()),  # writes x, y; reads the raw frames once
        "cgat_loader_gather_planar": ("hbm", n_pix * rec * 2 * esz + frames_h.numel()),
    }
    if pre:
        # fused conv + attention kernels: forward reads x and writes out; backward reads x and d(out) (the projected
        # features never touch HBM).  flops: the dense block-diagonal conv GEMM(s) actually executed on tcgen05.
        fl = 2.0 * n_pix * rec * heads * rec * 9
        alg["cgat_layer_fwd"] = ("hbm", n_pix * 2 * rec * esz, fl)
        alg["cgat_layer_bwd"] = ("hbm", n_pix * 2 * rec * esz, 2 * fl)
        alg["cgat_layer_train"] = ("hbm", n_pix * 2 * rec * esz, 2 * fl)  # reads x and y; out / d(out) stay on chip
        cin, cout, k = rec, heads * rec, 3
        conv_flops = 2.0 * n_pix * cin * cout * k * k  # dense block-diagonal implicit GEMM actually executed
        conv_bytes = n_pix * (cin + cout) * esz
        for nm in ("cgat_conv2d_fprop", "cgat_conv2d_fprop_packed", "cgat_conv2d_dgrad", "cgat_conv2d_dgrad_packed",
                   "cgat_conv2d_wgrad", "cgat_conv2d_wgrad_partial"):
            alg[nm] = ("hbm", conv_bytes, conv_flops)
    kernels = {}
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
        ent = {"launches_per_step": cnt / max(3, min(args.steps, 10)), "ms": ms}
        if name in alg:
            ent["gbs"] = alg[name][1] / (ms * 1e-3) / 1e9
            if len(alg[name]) > 2:
                ent["tflops"] = alg[name][2] / (ms * 1e-3) / 1e12
        kernels[name] = ent
    dom = max((n for n in prof if n in alg), key=lambda n: prof[n][0] * prof[n][1])
    achieved = alg[dom][1] / (prof[dom][1] * 1e-3) / 1e9
    traffic, ncu_entry = None, {}
    try:
        ncu_entry = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom, {})
        traffic = ncu_entry.get("bytes")
    except (OSError, ValueError):
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": alg[dom][1], "kernel_ms": prof[dom][1]}
    if len(alg[dom]) > 2:  # the kernel also runs the conv GEMMs on tcgen05: report the tensor-pipe side as well
        kms = prof[dom][1]
        # executed: the dense block-diagonal GEMMs the tensor cores run (3/4 of it multiplies structural zeros);
        # algorithmic: the shared per-node conv the layer defines (fprop + wgrad), i.e. executed / nodes
        nodes = T if args.type == "temporal" else V
        roofline["tensor_tflops_executed"] = alg[dom][2] / (kms * 1e-3) / 1e12
        roofline["tensor_tflops"] = roofline["tensor_tflops_executed"] / nodes
        roofline["tensor_frac"] = roofline["tensor_tflops"] / pk["bf16_tflops_sustained"]
        # the bound that applies: warp-instruction issue.  ncu counts the kernel's executed warp instructions
        # (smsp__inst_executed.sum, profiles/traffic.json); a B200 SM sub-partition issues at most one per cycle
        # (tools/microbench/fma_rate.cu: FFMA 0.98 / cycle, packed HFMA2 0.50), 148 SMs x 4 sub-partitions.
        inst = ncu_entry.get("warp_instructions")
        if inst and clocks and clocks.get("sm_mhz"):
            bound_ms = inst / (148 * 4) / (clocks["sm_mhz"] * 1e3)
            roofline["issue_bound"] = {"warp_instructions": inst, "bound_ms": bound_ms, "frac": bound_ms / kms,
                                       "source": ncu_entry.get("source"),
                                       "note": "kernel time if every sub-partition issued one warp instruction per cycle at "
                                               "the sampled SM clock; the packed-half instructions occupy the FP pipe for "
                                               "two cycles each, so the reachable figure is lower"}
        roofline["note"] = ("fused conv+attention kernel: neither HBM- nor tensor-bound by design (the projected features never "
                            "leave the SM); bound by warp-instruction issue of the per-pixel attention math "
                            "(roofline.issue_bound, profiles/*_sass_hot.txt)")

    cpu_rate, cpu_sec, cores = cpu_reference_step_rate(4, 3, 1, args.type, args.mapping) if world == 1 else (None, None, None)

    ms_per_step = total_ms / args.steps
    line = {
        "metric": METRIC, "value": world * B / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "dtype_detail": "bf16 activations and tensor-core operands, fp32 accumulation and parameter gradients; the per-pixel "
                        "attention math of cgat_layer_train in packed fp16 (two pixels per instruction) behind a range guard, "
                        "with an in-graph fp32 re-run of the step when the guard fires",
        "config": dict(workload_config(args, B), gradient_exchange=("none (1 GPU)" if world == 1 else
                       "cgat_p2p_allreduce_adam: push + rank-ordered sum + Adam in one kernel over NVLink peer memory"
                       if p2p else "NCCL all_reduce + cgat_adam_step")),
        "e2e": {"value": world * B / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": frames_h.numel() + start_h.numel() * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps, "host_enqueue_ms_per_step": e2e_host[0],
                "d2h": ("the step's loss (4 bytes) written by the step's last kernel (cgat_stream_finish_mirror) into a ring in "
                        "mapped pinned host memory, every step" if mirror else "4-byte cudaMemcpyAsync of the loss after every step"),
                "input": "raw uint8 frames [B+7,V,H,W] + int32 window starts from pinned host memory (the KNMI loader's "
                         "on-disk format); sliding windows, /254 and the layouts the train kernel reads (x chunk-planar [N,T*V/8,H,W,8], "
                         "y [N,H,W,T,V]) by cgat_loader_gather_planar on the device",
                "xy_tensor_copy": {"value": world * B / (e2e_xy_ms / args.steps * 1e-3), "ms_per_step": e2e_xy_ms / args.steps,
                                   "h2d_bytes_per_step": xh.numel() * xh.element_size() + yh.numel() * yh.element_size(),
                                   "input": "finished x, y bf16 tensors copied per step (PCIe-bound)"}},
        "gpu_launches": per_step_launches * args.steps, "gpu_launches_per_step": per_step_launches,
        "roofline": roofline, "kernels": kernels, "clocks": clocks, "final_loss": final_loss,
    }
    if args.l2 == "rotate":
        line["host_enqueue_ms_per_step"] = host_enqueue[0]  # host time to enqueue one step: below ms_per_step = not host-bound
    if world == 1:
        try:
            line["conv_roofline"] = conv_tensor_roofline(pk, batch=B)
        except Exception as exc:  # the headline line must not depend on the side measurement
            line["conv_roofline"] = {"error": repr(exc)}
    if cpu_rate is not None:
        line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "4 samples/step x 3 steps of the same model (fp32 torch CPU, oracle/spec.py)"}
    if world == 1 and not args.no_baselines:
        # SURVEY 8(d) last row / BASELINE.md R1-R5: stock PyTorch (ATen / cuDNN / cuBLAS) on THIS B200 for the same
        # modules -- what a user of the reference would run on the box -- and the same rows on the host cores.  Timed
        # after, and outside of, our timed region.
        try:
            from oracle import baselines
            eager = {"fp32": baselines.reference_rows(dev, batch_gpu=B),
                     "bf16_autocast": baselines.reference_rows(dev, batch_gpu=B, autocast=True),
                     "ours": dict(ours_rows(dev, B), R3_convgat_config2_train_step={
                         "samples_per_s": line["value"], "ms_per_step": ms_per_step, "batch": B, "dtype": "bf16"}),
                     "note": "eager PyTorch on the same GPU, CUDA events, batch 64 for R3/R4 and 4 for R1/R2 (the only size the "
                             "reference's 2-D layer runs at); rows are restatements through oracle/spec.py (pinned to the live "
                             "reference), which is FASTER than the reference's own diag_embed formulation"}
            line["eager_gpu_baseline"] = eager
            line["cpu_baseline"]["rows"] = baselines.reference_rows("cpu")
            line["cpu_baseline"]["rows_note"] = ("R1-R4 of BASELINE.md section 3 on the host cores, fp32, batch 4; the unmodified "
                                                 "reference's R1 (diag_embed [N,V,V,P,P]) measured 2.1 samples/s at survey time")
        except Exception as exc:  # side measurements must not take the headline line down
            line["eager_gpu_baseline"] = {"error": repr(exc)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # keep the real stdout for the JSON line only: everything else (NCCL's version banner, warnings of native libraries,
    # stray prints) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU")
    ap.add_argument("--type", default="temporal", choices=["temporal", "spatial", "multi_stream"])
    ap.add_argument("--mapping", default="conv", choices=["conv", "linear"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the eager-GPU / CPU baseline rows (R1-R5)")
    ap.add_argument("--loss-readback", default="mirror", choices=["mirror", "memcpy"],
                    help="e2e: how each step's loss reaches the host (a posted write of the last kernel, or a memcpy per step)")
    ap.add_argument("--l2", default="rotate", choices=["rotate", "flush"],
                    help="cold-L2 rule of the timed region: rotate over input slots larger than L2, or flush between steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
