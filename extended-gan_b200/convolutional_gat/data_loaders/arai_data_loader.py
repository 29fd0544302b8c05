"""``DataLoader`` / ``get_loaders`` of ``convolutional_gat/data_loaders/arai_data_loader.py`` (reference :14-224) with the
windowing and the layout change on the GPU (``cgat_loader_gather_f32``, csrc/loader_kernels.cu).

Same constructor arguments and attributes (``total_length``, ``n_regions``, ``downsample_size``, ``batch_size``,
``time_steps``, ``files`` sorted numerically, ``item_count``), same ``__len__`` formula (:53-55) and the same batches:
a file ``[L, regions, 1, H, W]`` of floats is cropped to ``downsample_size`` (:144), cut into windows ``i .. i+7``
(:74-84), served ``batch_size`` windows at a time with the short tail and without merging across files (:166-182), each
batch permuted to ``[N, H, W, T, V]`` (:89-96).  The reference stacks every overlapping window on the host (8x the file)
and copies finished batches; here a file's frames are uploaded once and a batch is one kernel launch.

End of the pass: after the reference has read its last file no further batch is prepared (:110-115), so of the last file
only the first batch exists -- and whether ``__next__`` hands it out is a race between the consumer and the reader
thread (:99-104).  This loader always serves it (the outcome pinned by tests/golden/arai_loader.pt).  The reference's
prefetch ``Thread`` has no counterpart: the file upload is an asynchronous copy from pinned memory.
"""
from __future__ import annotations

import json
import os

import torch as t

from cgat import _lib


def gather_windows_f32(frames: t.Tensor, start: t.Tensor, *, downsample_size=(256, 256), steps: int = 4, dtype=t.float32):
    """``frames [L, V, H, W]`` fp32 (device), ``start [N]`` int32 (device) -> ``(x, y)`` ``[N, H', W', steps, V]``."""
    _lib.require_cuda(frames, start)
    if frames.dtype != t.float32 or start.dtype != t.int32:
        raise RuntimeError("gather_windows_f32 takes fp32 frames and int32 window starts")
    frames = frames.contiguous()
    L, V, H, W = frames.shape
    ch, cw = min(downsample_size[0], H), min(downsample_size[1], W)
    n = start.numel()
    x = t.empty(n, ch, cw, steps, V, device=frames.device, dtype=dtype)
    y = t.empty_like(x)
    _lib.call("cgat_loader_gather_f32", _lib.ptr(frames), L, _lib.ptr(start), _lib.ptr(x), _lib.ptr(y), n, V, H, W, ch, cw,
              steps, _lib.dtype_tag(x), _lib.stream())
    return x, y


class DataLoader:
    def __init__(self, batch_size: int, folder: str, device, *, total_length: int, n_regions: int = 5, time_steps: int = 4,
                 norm_max=None, norm_min=None, downsample_size=(256, 256), dtype=t.float32):
        self.total_length = total_length
        self.n_regions = n_regions
        self.downsample_size = downsample_size
        self.folder = folder
        self.device = t.device(device)
        self.norm_max = norm_max  # stored and never applied, as in the reference (:32-33)
        self.norm_min = norm_min
        self.batch_size = batch_size
        self.time_steps = time_steps
        self.dtype = dtype
        self.file_index = 0
        self.should_stop_iteration = False
        self.files = sorted(os.listdir(folder), key=lambda x: int(x.split(".")[0]))  # :44-46
        self.item_count = 86 * len(self.files)  # :49
        self._frames = None
        self._next_window = 0
        self._n_windows = 0

    def __len__(self):
        tot = self.total_length - (self.time_steps - 1) * (len(self.files) + 1)  # :54
        return tot // self.batch_size

    def __iter__(self):
        return self

    def __read_next_file(self):
        data = t.load(os.path.join(self.folder, self.files[self.file_index]))  # [L, regions, 1, H, W]
        self.file_index += 1
        if data.dim() != 5 or data.shape[2] != 1:
            raise RuntimeError(f"ARAI file of shape {tuple(data.shape)}: expected [L, regions, 1, H, W]")
        frames = data[:, :, 0].to(t.float32).contiguous()
        self._frames = frames.pin_memory().to(self.device, non_blocking=True)
        self._n_windows = max(0, len(data) - 2 * self.time_steps + 1)  # :74-77
        self._next_window = 0
        self._last_file = self.file_index == len(self.files)

    def __next__(self):
        if self.should_stop_iteration:
            raise StopIteration
        if self._next_window >= self._n_windows:
            self.__read_next_file()
        first = self._next_window
        n = min(self.batch_size, self._n_windows - first)
        self._next_window += n
        if self._last_file:  # of the last file only the first batch exists (:156-157, :110-115)
            self.should_stop_iteration = True
        if n <= 0:
            raise StopIteration
        start = t.arange(first, first + n, dtype=t.int32).to(self.device)
        return gather_windows_f32(self._frames, start, downsample_size=self.downsample_size, steps=self.time_steps,
                                  dtype=self.dtype)


def get_loaders(train_batch_size: int, test_batch_size: int, preprocessed_folder: str, device, *,
                downsample_size=(256, 256)):
    """reference :194-231 (validation and test loaders both read the ``validation`` folder there too)."""
    with open(os.path.join(preprocessed_folder, "metadata.json")) as f:
        metadata = json.load(f)
    mk = lambda bs, sub: DataLoader(bs, os.path.join(preprocessed_folder, sub), device,
                                    total_length=metadata[sub]["length"], downsample_size=downsample_size,
                                    n_regions=metadata["n_regions"])
    return mk(train_batch_size, "training"), mk(test_batch_size, "validation"), mk(test_batch_size, "validation")
