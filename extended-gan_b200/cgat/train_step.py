"""The reference train step (convolutional_gat/train.py:129-133, :212) as a replayable device program.

    optimizer.zero_grad(); y_hat = model(x)
    loss = MSE(y_hat, y) - 0.0005 * sum(y_hat)/numel          (:131)
    loss.backward(); optimizer.step()                          (Adam(lr, weight_decay=0.01), :212)

* parameters that receive a gradient live in ONE flat fp32 buffer (``flat_param``), their gradients in
  another (``flat_grad``); ``p.data`` / ``p.grad`` are views, so the model's ``state_dict()`` is unchanged;
* parameters that never receive a gradient (the reference registers an ``output_layer`` it does not
  call, model.py:44-47) are left untouched, exactly as ``torch.optim.Adam`` skips ``grad is None``;
* loss + d loss/d y_hat is one kernel, Adam is one kernel over the flat buffer;
* data parallel: the batch is sharded over ranks, ``flat_grad`` is all-reduced ONCE per step (NCCL sum over
  NVLink) and the 1/world scale is folded into the Adam kernel;
* forward + loss + backward are captured in a CUDA graph (one ``cudaGraphLaunch`` per step); the all-reduce
  and Adam follow on the same stream.
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import functional
from .functional import adam_step_, loss_and_grad
from .parallel import FlatParams


class TrainStep:
    P2P_POLL_EVERY = 16  # steps between asynchronous reads of the P2P exchange's time-out marker

    def __init__(self, model: torch.nn.Module, example_x: torch.Tensor, example_y: torch.Tensor, lr: float = 1e-3,
                 weight_decay: float = 0.01, lam: float = 0.0005, betas=(0.9, 0.999), eps: float = 1e-8,
                 use_graph: bool = True, process_group: Optional[dist.ProcessGroup] = None, fuse_loss: bool = True,
                 precision: str = "auto"):
        """``precision`` of the fused train kernel: ``"auto"`` (default) = packed-fp16 attention math with the in-graph
        range guard and fp32 re-run (functional.gat_stream_train), ``"fp16x2"`` = unguarded, ``"fp32"`` = the fp32
        instantiation only."""
        if precision not in ("auto", "fp16x2", "fp32"):
            raise ValueError(f"precision must be 'auto', 'fp16x2' or 'fp32', got {precision!r}")
        self.precision = precision
        self.model = model
        self._hyper = None
        self._lr = lr
        self.weight_decay, self.lam, self.betas, self.eps = weight_decay, lam, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1
        self.x = example_x.clone()
        self.y = example_y.clone()
        self.device = example_x.device
        self._step = 0  # optimiser steps taken (host side: Adam after an NCCL all-reduce takes it by value)
        self._p2p_in_graph = False
        self._mirror = None  # (pinned host ring, device cursor): the fused step writes its loss to host memory itself
        self.graph = None
        self._slots = None  # double-buffered inputs for pipelined host->device loading (enable_prefetch)
        self.fused_stream = self._single_stream(model, example_x) if fuse_loss else None
        # the fused train kernel reads x CHUNK-PLANAR ([N, T*V/8, H, W, 8]: TMA boxes with 128-byte rows); ``xp`` is that
        # copy of ``x``.  The loader kernel writes it directly (prefetch_raw); record tensors handed to load_batch / step /
        # prefetch are converted by one small kernel outside the captured graph.
        self.xp = functional.records_to_planar(self.x) if self.fused_stream is not None else None
        self._flatten()
        # Single GPU + fused train kernel: Adam is applied by the launch that finishes the gradients (cgat_stream_finish); its
        # step counter and hyper-parameters live on the device so that the captured graph stays valid when a scheduler
        # changes lr.  With a gradient exchange (world > 1) the optimiser stays behind it (P2P kernel / all-reduce + Adam).
        self._adam_in_graph = self.fused_stream is not None and self.world == 1
        self._step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._hyper = torch.empty(6, dtype=torch.float32, device=self.device)
        self._push_hyper()
        if use_graph:
            self._capture()

    @property
    def lr(self) -> float:
        return self._lr

    @lr.setter
    def lr(self, value: float):
        changed = value != self._lr
        self._lr = value
        if changed and self._hyper is not None:
            self._push_hyper()

    def _push_hyper(self):
        self._hyper.copy_(torch.tensor([self._lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, 1.0 / self.world],
                                       dtype=torch.float32), non_blocking=False)

    @staticmethod
    def _single_stream(model, x):
        """The model's only stream when its forward is ONE conv-mapped GAT stream (the Spatial/Temporal wrappers of
        convolutional_gat/model.py:8-88 call just ``hidden_layer``) and the fused train kernel serves it, else None."""
        m = getattr(model, "net", model)
        if not hasattr(m, "hidden_layer") or getattr(m, "_use_output", True):
            return None
        stream = getattr(m.hidden_layer, "stream", None)
        if stream is None or not hasattr(stream, "train_step_supported") or not stream.train_step_supported(x):
            return None
        return stream

    # -- flat buffers ---------------------------------------------------------------------------
    @contextlib.contextmanager
    def _preserve_module_state(self):
        """Probe / warm-up passes run the model eagerly before step 1; for stateful modules (BatchNorm running statistics
        and ``num_batches_tracked`` of UnetModel / SmaAt-UNet, dropout's RNG) that must not count as training: module
        buffers and the RNG state are put back afterwards, so step 1 starts where the reference loop would."""
        bufs = [(b, b.detach().clone()) for b in self.model.buffers()]
        cpu_rng = torch.get_rng_state()
        cuda_rng = torch.cuda.get_rng_state(self.device) if self.device.type == "cuda" else None
        try:
            yield
        finally:
            with torch.no_grad():
                for b, c in bufs:
                    if not torch.equal(b, c):  # (untouched buffers keep their version counter: caches keyed on it stay valid)
                        b.copy_(c)
            torch.set_rng_state(cpu_rng)
            if cuda_rng is not None:
                torch.cuda.set_rng_state(cuda_rng, self.device)

    def _flatten(self):
        model = self.model
        for p in model.parameters():
            p.grad = None
        # probe which parameters take part in the forward (eager, outside any graph)
        with self._preserve_module_state():
            out = model(self.x)
            _, dy = loss_and_grad(out, self.y, self.lam)
            out.backward(dy)
        self.active = [(n, p) for n, p in model.named_parameters() if p.grad is not None and p.requires_grad]
        self.inactive = [n for n, p in model.named_parameters() if p.grad is None]
        self.flat = FlatParams(self.active)
        self.flat_param, self.flat_grad = self.flat.param, self.flat.grad
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        # loss scalars and the fused kernel's a / adjacency accumulators live in the gradient buffer's scratch tail
        # (scratch layout: [loss, mse, range guard, -, -, -, -, -, accumulators ... | 256: the fp32 re-run's set])
        self.loss, self.mse = self.flat.scratch[0:1], self.flat.scratch[1:2]
        self._acc = self.flat.scratch[8:256]

    @property
    def range_guard_fired(self) -> bool:
        """Whether the LAST step left the packed-fp16 kernel's range and was recomputed in fp32 (host sync; diagnostics)."""
        return bool(self.flat.scratch[2].item())

    @property
    def step_count(self) -> torch.Tensor:
        """Optimiser steps taken, as a device tensor (kept for checkpoint code written against the earlier attribute)."""
        return torch.tensor([self._step], device=self.device, dtype=torch.int64)

    def sync_params(self):
        """Broadcast rank 0's parameters (data-parallel start state)."""
        self.flat.broadcast_params(0, self.pg)

    def enable_p2p_exchange(self) -> bool:
        """Replace all-reduce + Adam by the single peer-memory kernel (``cgat_p2p_allreduce_adam``) where available.  The
        kernel then becomes the LAST NODE of the step's captured graph (``cgat_p2p_allreduce_adam_graph``: step counter and
        hyper-parameters on the device, programmatic dependent launch behind the gradient kernel), so a data-parallel
        step stays one graph launch; call it before ``enable_prefetch`` / ``enable_raw_pipeline``."""
        ok = self.flat.enable_p2p(self.pg)
        self._p2p_in_graph = bool(ok)
        if ok:
            if self._slots is not None:
                raise RuntimeError("enable_p2p_exchange must be called before enable_prefetch / enable_raw_pipeline")
            if self.graph is not None:
                self._capture()  # re-capture with the exchange node
        return ok

    def enable_loss_mirror(self, n: int = 1024) -> torch.Tensor:
        """Have the fused step's last launch write each step's loss into a ring of ``n`` floats in pinned HOST memory
        (``cgat_stream_finish_mirror``) instead of the caller copying ``loss`` back with a memcpy between two steps
        (train.py:135 reads ``loss.item()`` every batch).  Returns the ring; ``loss_of_step(k)`` reads it (after a
        synchronisation point of the caller's choice).  Call before ``enable_prefetch`` / ``enable_raw_pipeline``."""
        if self.fused_stream is None:
            raise RuntimeError("enable_loss_mirror serves the fused train step")
        if self._slots is not None:
            raise RuntimeError("enable_loss_mirror must be called before enable_prefetch / enable_raw_pipeline")
        ring = torch.zeros(n, dtype=torch.float32).pin_memory()
        self._mirror = (ring, torch.zeros(1, dtype=torch.int32, device=self.device))
        if self.graph is not None:
            self._capture()
        return ring

    def loss_of_step(self, k: int) -> float:
        """Loss of the k-th fused step taken since ``enable_loss_mirror`` (0-based), from the host ring."""
        ring = self._mirror[0]
        return float(ring[k % ring.numel()])

    def _step_launches(self):
        """Everything one optimisation step enqueues on the device: forward + loss + backward (+ Adam on one GPU), then
        -- with the peer-memory exchange -- the exchange + Adam kernel."""
        self._fwd_bwd(with_adam=True)
        if self._p2p_in_graph:
            import ctypes

            from . import _lib

            p2p = self.flat.p2p
            _lib.call("cgat_p2p_allreduce_adam_graph", ctypes.cast(p2p["ptrs"], ctypes.c_void_p), p2p["rank"], p2p["world"],
                      _lib.ptr(self.flat_grad), _lib.ptr(self.flat_param), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                      _lib.ptr(self._step_dev), _lib.ptr(self._hyper), self.flat_param.numel(), _lib.stream())

    def _exchange_and_update(self):
        """Gradient mean over ranks + Adam(lr, weight_decay): one P2P kernel, or NCCL all-reduce + fused Adam."""
        self._step += 1
        if self._adam_in_graph:
            return  # applied by cgat_stream_finish inside the step's launch sequence
        p2p = getattr(self.flat, "p2p", None)
        if p2p is not None:
            import ctypes

            from . import _lib

            # a rank that gave up waiting for a peer skipped its update (csrc/p2p_kernels.cu): that is fatal, not a
            # warning -- the replicas no longer hold the same parameters.  Polled without a stream sync.
            if self.flat.p2p_poll_timeout(refresh=self._step % self.P2P_POLL_EVERY == 0):
                raise RuntimeError(
                    f"cgat_p2p_allreduce_adam timed out waiting for a peer around step {self._step}: the ranks are out "
                    "of step (every rank must take the same number of steps) or a peer died; parameters were NOT updated "
                    "on this rank for that step.  Restart from a checkpoint, or train with the NCCL exchange.")
            if self._p2p_in_graph:
                return  # the exchange + Adam kernel was the last launch of the step (_step_launches)

            _lib.call("cgat_p2p_allreduce_adam", ctypes.cast(p2p["ptrs"], ctypes.c_void_p), p2p["rank"], p2p["world"],
                      _lib.ptr(self.flat_grad), _lib.ptr(self.flat_param), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                      None, self._step, self.flat_param.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, _lib.stream())
            return
        self.flat.all_reduce_grads(self.pg)
        adam_step_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self._step, self.lr,
                   self.betas[0], self.betas[1], self.eps, self.weight_decay, 1.0 / self.world)

    # -- one step -------------------------------------------------------------------------------
    def _fwd_bwd(self, with_adam: bool = False):
        """Forward + loss + backward into the flat gradient buffer; ``with_adam``: the fused path also applies the optimiser
        step in its last launch (``run`` asks for it where ``_adam_in_graph``; probes and warm-ups never do)."""
        if self.fused_stream is not None and self.fused_stream.train_step_supported(self.x):
            # forward + loss + backward in one kernel (cgat_layer_train); gradients, loss scalars and accumulators are
            # cleared by the step's first launch (cgat_stream_prepare_clear), not by a memset node of their own
            clear = self.flat.grad_all if self.flat.grad_all.data_ptr() % 16 == 0 and self.flat.grad_all.numel() % 4 == 0 else None
            if clear is None:
                self.flat.zero_grad()
            adam = None
            if with_adam and self._adam_in_graph:
                adam = (self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self._step_dev, self._hyper)
            self.fused_stream.fused_train_step(self.x, self.y, self.lam, self.loss, self.mse, self._acc, x_planar=self.xp,
                                               scratch=self.flat.scratch, adam=adam, clear=clear, mirror=self._mirror,
                                               precision="fp32" if self.precision == "fp32" else
                                               ("fp16x2" if self.precision == "auto" else "fp16x2-unguarded"))
            return
        self.flat.zero_grad()  # gradients, loss scalars and accumulators: one memset
        prev, functional.DIRECT_GRAD = functional.DIRECT_GRAD, True  # param-grad kernels add into flat_grad views
        try:
            out = self.model(self.x)
            _, dy = loss_and_grad(out, self.y, self.lam, loss_out=self.loss, mse_out=self.mse)
            out.backward(dy)
        finally:
            functional.DIRECT_GRAD = prev

    def _capture(self):
        cursor0 = self._mirror[1].clone() if self._mirror is not None else None  # warm-ups below must not count as steps
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with self._preserve_module_state():
            with torch.cuda.stream(s):
                for _ in range(2):  # warm up allocator / lazy init on the side stream
                    self._fwd_bwd()
            torch.cuda.current_stream().wait_stream(s)  # (before the buffers are restored on the current stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_launches()
        if cursor0 is not None:
            self._mirror[1].copy_(cursor0)

    def load_batch(self, x: torch.Tensor, y: torch.Tensor):
        """Copy one batch (host-pinned or device) into the static input buffers."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.refresh_planar()

    def refresh_planar(self, slot: Optional[int] = None):
        """Rebuild the chunk-planar copy after ``x`` (of ``slot``, or the current buffers) was written in record layout."""
        if self.xp is None:
            return
        if slot is None:
            functional.records_to_planar(self.x, out=self.xp)
        else:
            functional.records_to_planar(self._slots[slot]["x"], out=self._slots[slot]["xp"])

    def run(self) -> torch.Tensor:
        """One optimisation step on the loaded batch; returns the (device) loss of this rank's shard."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_launches()
        self._exchange_and_update()
        return self.loss

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """One optimisation step on ``(x, y)``.  A batch of the captured shape replays the graph; any other batch size
        (the short last batch of a file, kmni_data_loader.py:113) runs the same launches eagerly."""
        if x.shape == self.x.shape:
            self.load_batch(x, y)
            return self.run()
        keep = (self.x, self.y, self.graph, self.xp)
        self.x, self.y, self.graph = x.to(self.x.dtype).contiguous(), y.to(self.y.dtype).contiguous(), None
        if self.xp is not None:
            self.xp = functional.records_to_planar(self.x)
        try:
            return self.run()
        finally:
            self.x, self.y, self.graph, self.xp = keep

    # -- pipelined loading: batch i+1 crosses PCIe on a copy stream while batch i trains ------------------
    def enable_prefetch(self, n_slots: int = 2):
        """``n_slots`` input slots, each with its own captured graph, plus a copy stream and the events that order them."""
        if self._slots is not None:
            if len(self._slots) >= n_slots:
                return
            raise RuntimeError("enable_prefetch was already called with fewer slots")
        if self.graph is None:
            raise RuntimeError("enable_prefetch needs use_graph=True")
        slots = [dict(x=self.x, y=self.y, graph=self.graph, xp=self.xp)]
        for _ in range(1, n_slots):
            self.x, self.y = torch.empty_like(self.x), torch.empty_like(self.y)
            self.x.copy_(slots[0]["x"])
            self.y.copy_(slots[0]["y"])
            if self.xp is not None:
                self.xp = slots[0]["xp"].clone()
            self._capture()
            slots.append(dict(x=self.x, y=self.y, graph=self.graph, xp=self.xp))
        self._slots = slots
        for s in self._slots:
            s["ready"] = torch.cuda.Event()   # inputs of this slot have landed
            s["free"] = torch.cuda.Event()    # the step that read this slot has finished
            s["ready"].record()
            s["free"].record()
        self.copy_stream = torch.cuda.Stream()

    def prefetch(self, x_host: torch.Tensor, y_host: torch.Tensor, slot: int):
        """Asynchronous pinned-host -> device copy of one batch into ``slot`` on the copy stream."""
        s = self._slots[slot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["free"])
            s["x"].copy_(x_host, non_blocking=True)
            s["y"].copy_(y_host, non_blocking=True)
            if s["xp"] is not None:
                functional.records_to_planar(s["x"], out=s["xp"])
            s["ready"].record(self.copy_stream)

    def prefetch_raw(self, frames_host: torch.Tensor, start_host: torch.Tensor, slot: int, *, normalizing_max=254.0,
                     power=1.0):
        """The loader's raw format end to end: pinned uint8 frames ``[L, V, H, W]`` and int32 window starts ``[N]`` cross
        PCIe on the copy stream, then ``cgat_loader_gather`` builds this slot's x and y on the device
        (kmni_data_loader.py:72-127)."""
        from convolutional_gat.data_loaders.kmni_data_loader import gather_windows

        s = self._slots[slot]
        if "frames" not in s or s["frames"].shape != frames_host.shape:
            s["frames"] = torch.empty(frames_host.shape, dtype=torch.uint8, device=self.device)
            s["start"] = torch.empty(start_host.shape, dtype=torch.int32, device=self.device)
            s.pop("raw_key", None)

        def enqueue():
            s["frames"].copy_(frames_host, non_blocking=True)
            s["start"].copy_(start_host, non_blocking=True)
            N, H, W, T, V = s["x"].shape
            if s["xp"] is not None:  # x straight into the train kernel's chunk-planar format (the record copy is not needed)
                gather_windows(s["frames"], s["start"], crop=H, steps=T, normalizing_max=normalizing_max, power=power,
                               out=(s["xp"], s["y"]), planar=True)
            else:
                gather_windows(s["frames"], s["start"], crop=H, steps=T, normalizing_max=normalizing_max, power=power,
                               out=(s["x"], s["y"]))

        # A loader that refills the SAME pinned staging buffers every batch (the usual arrangement) gets the two copies
        # and the gather kernel as one captured graph per slot: one launch instead of ~8 host calls per batch -- the
        # end-to-end loop is otherwise bound by the host's enqueue rate, not by PCIe or the GPU.
        if start_host.numel() != s["x"].shape[0]:
            raise RuntimeError(f"prefetch_raw: {start_host.numel()} window starts for a slot of {s['x'].shape[0]} samples "
                               "(short last batches go through TrainStep.step)")
        key = (frames_host.data_ptr(), start_host.data_ptr(), tuple(frames_host.shape), tuple(start_host.shape),
               float(normalizing_max), float(power))
        graphable = frames_host.is_pinned() and start_host.is_pinned() and not os.environ.get("CGAT_NO_RAW_GRAPH")
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["free"])
            if graphable and s.get("raw_key") == key:
                s["raw_graph"].replay()
            else:
                enqueue()
                if graphable:
                    self.copy_stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.copy_stream):
                        enqueue()
                    s["raw_graph"], s["raw_key"] = g, key
            s["ready"].record(self.copy_stream)

    def run_slot(self, slot: int) -> torch.Tensor:
        """One optimisation step on the batch previously prefetched into ``slot``."""
        s = self._slots[slot]
        cur = torch.cuda.current_stream()
        cur.wait_event(s["ready"])
        s["graph"].replay()
        s["free"].record(cur)
        self._exchange_and_update()
        return self.loss

    # -- the raw-frame pipeline as ONE graph launch per step ------------------------------------------------------------
    def enable_raw_pipeline(self, frames_host: torch.Tensor, start_host: torch.Tensor, *, normalizing_max=254.0, power=1.0):
        """End-to-end loop for a loader that refills the SAME pinned staging buffers every batch (raw uint8 frames
        ``[L, V, H, W]`` + int32 window starts ``[N]``, kmni_data_loader.py:72-127).  Two input slots; for each ONE captured
        graph holding two parallel branches: the train step on this slot's batch, and -- on a forked stream -- the
        host-to-device copies plus ``cgat_loader_gather_planar`` of the NEXT batch into the other slot.  ``run_pipelined``
        is then one graph launch per step (the two-graph / event version of ``prefetch_raw`` + ``run_slot`` costs the host
        more than the GPU needs for the step).  Call ``prime_raw_pipeline()`` after filling the staging buffers with the
        first batch."""
        from convolutional_gat.data_loaders.kmni_data_loader import gather_windows

        if not (frames_host.is_pinned() and start_host.is_pinned()):
            raise RuntimeError("enable_raw_pipeline needs pinned staging buffers")
        if self.fused_stream is None or self.xp is None:
            raise RuntimeError("enable_raw_pipeline serves the fused train step (planar x)")
        if start_host.numel() != self.x.shape[0]:
            raise RuntimeError(f"{start_host.numel()} window starts for a step of {self.x.shape[0]} samples")
        self.enable_prefetch(2)
        N, H, W, T, V = self.x.shape
        for s in self._slots[:2]:
            s["frames"] = torch.empty(frames_host.shape, dtype=torch.uint8, device=self.device)
            s["start"] = torch.empty(start_host.shape, dtype=torch.int32, device=self.device)
        self._raw_host = (frames_host, start_host, float(normalizing_max), float(power))

        def fill(slot):
            slot["frames"].copy_(frames_host, non_blocking=True)
            slot["start"].copy_(start_host, non_blocking=True)
            gather_windows(slot["frames"], slot["start"], crop=H, steps=T, normalizing_max=normalizing_max, power=power,
                           out=(slot["xp"], slot["y"]), planar=True)

        self._raw_fill = fill
        side = torch.cuda.Stream()
        keep = (self.x, self.y, self.xp, self.graph)
        torch.cuda.synchronize()
        for k in range(2):
            cur_slot, other = self._slots[k], self._slots[1 - k]
            self.x, self.y, self.xp = cur_slot["x"], cur_slot["y"], cur_slot["xp"]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    fill(other)
                self._step_launches()
                cur.wait_stream(side)
            cur_slot["pipe_graph"] = g
        self.x, self.y, self.xp, self.graph = keep

    def prime_raw_pipeline(self):
        """Load the batch currently in the staging buffers into slot 0 (before the first ``run_pipelined(0)``)."""
        self._raw_fill(self._slots[0])

    def run_pipelined(self, slot: int) -> torch.Tensor:
        """One optimisation step on ``slot``'s batch while the staging buffers' batch is loaded into the other slot."""
        self._slots[slot]["pipe_graph"].replay()
        self._exchange_and_update()
        return self.loss
