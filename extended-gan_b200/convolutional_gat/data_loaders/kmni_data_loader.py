"""``DataLoader`` of ``convolutional_gat/data_loaders/kmni_data_loader.py`` (reference :15-127) with the windowing,
normalisation and layout change on the GPU (``cgat_loader_gather``, csrc/loader_kernels.cu).

Same constructor arguments, attributes (``power``, ``normalizing_max``, ``batch_size``, ``files``) and iteration
protocol; a batch is the same ``(x, y)`` pair of ``[N, H, W, T, V]`` tensors.  What changes is where the work
happens: the reference builds every overlapping 8-frame window on the host in fp32 (:72-96) and copies
``2*N*4*V*H*W`` floats per batch (:114); here a file's raw frames ``[L, V, H, W]`` cross PCIe once as uint8 and each
batch is one kernel launch writing straight into the pixel-record layout the conv-GAT kernels read.  Batch
composition follows the reference exactly: files in (shuffled) order, a file truncated to a multiple of 8 frames (:74),
windows ``i .. i+7`` for ``i <= len-8`` (:79-85), ``batch_size`` windows per batch with the remainder carried to the
next call (:109-113), and a per-batch ``randperm`` when ``shuffle`` (:115-117).  ``merge_nodes`` (:97-108) is not on
the conv-GAT path and is rejected.
"""
from __future__ import annotations

import ctypes
import os

import torch as t

from cgat import _lib


def gather_windows(frames_u8: t.Tensor, start: t.Tensor, *, crop=None, steps: int = 4, normalizing_max: float = 254.0,
                   power: float = 1.0, dtype=t.float32, out=None, planar: bool = False):
    """``frames_u8 [L, V, H, W]`` uint8 (device), ``start [N]`` int32 (device) -> ``(x, y)`` ``[N, H', W', steps, V]``.

    ``planar=True`` (bf16 only): x is written PADDED CHUNK-PLANAR ``[N, steps*V/8, H', padded_width(W'), 8]`` -- the fused
    layer kernels' input format (``cgat_loader_gather_planar``; ``cgat.functional.planar_zeros`` allocates it: the padding
    columns are never written); y keeps the record layout."""
    _lib.require_cuda(frames_u8, start)
    if frames_u8.dtype != t.uint8 or start.dtype != t.int32:
        raise RuntimeError("gather_windows takes uint8 frames and int32 window starts")
    frames_u8 = frames_u8.contiguous()
    L, V, H, W = frames_u8.shape
    ch = H if crop is None else min(crop, H)
    cw = W if crop is None else min(crop, W)
    n = start.numel()
    if planar:
        from cgat.functional import planar_shape, planar_zeros

        if out is None:
            x = planar_zeros((n, ch, cw, steps, V), frames_u8.device)
            y = t.empty(n, ch, cw, steps, V, device=frames_u8.device, dtype=t.bfloat16)
        else:
            x, y = out
        if x.dtype != t.bfloat16 or y.dtype != t.bfloat16 or tuple(x.shape) != planar_shape((n, ch, cw, steps, V)):
            raise RuntimeError("gather_windows(planar=True) writes bf16 x [N, steps*V/8, H, padded_width(W), 8] (zeroed "
                               "padding) and y [N, H, W, steps, V]")
        _lib.call("cgat_loader_gather_planar", _lib.ptr(frames_u8), L, _lib.ptr(start), _lib.ptr(x), _lib.ptr(y), n, V, H, W,
                  ch, cw, steps, float(normalizing_max), float(power), _lib.stream())
        return x, y
    if out is None:
        x = t.empty(n, ch, cw, steps, V, device=frames_u8.device, dtype=dtype)
        y = t.empty_like(x)
    else:
        x, y = out
    _lib.call("cgat_loader_gather", _lib.ptr(frames_u8), L, _lib.ptr(start), _lib.ptr(x), _lib.ptr(y), n, V, H, W, ch, cw,
              steps, float(normalizing_max), float(power), _lib.dtype_tag(x), _lib.stream())
    return x, y


class DataLoader:
    def __init__(self, batch_size: int, folder: str, device, *, time_steps: int = 4, crop=None, shuffle: bool = True,
                 merge_nodes: bool = False, power: float = 1.0, dtype=t.float32):
        if merge_nodes:
            raise NotImplementedError("merge_nodes is not on the conv-GAT path (kmni_data_loader.py:97-108)")
        self.power = t.tensor(power)
        self.data_folder = folder
        self.normalizing_max = 254
        self.merge_nodes = merge_nodes
        self.crop = crop
        self.device = t.device(device)
        self.batch_size = batch_size
        self.time_steps = time_steps
        self.dtype = dtype
        self.file_index = 0
        self.folder = folder
        self.files = tuple(os.path.join(folder, fn) for fn in sorted(os.listdir(folder)))
        self.shuffle = shuffle
        if self.shuffle:
            rand_indices = t.randperm(len(self.files))
            self.files = tuple(self.files[i] for i in rand_indices)
        self._frames = None  # raw frames of the current file on the device (uint8)
        self._next_window = 0
        self._n_windows = 0
        self.__read_next_file()
        self.file_length = self._n_windows * 2

    def __read_next_file(self):
        if self.file_index == len(self.files):
            raise StopIteration
        data = t.load(self.files[self.file_index])
        self.file_index += 1
        data = data[: (len(data) // 8) * 8]  # :74
        if data.numel() and (int(data.min()) < 0 or int(data.max()) > 255):
            raise RuntimeError("raw KNMI frames are expected in 0..255")
        self._frames = data.to(t.uint8).pin_memory().to(self.device, non_blocking=True)
        self._n_windows = max(0, len(data) - 2 * self.time_steps + 1)  # windows i..i+7 with len(el) == 8 (:79-85)
        self._next_window = 0

    def __next__(self):
        if self._next_window >= self._n_windows:
            self.__read_next_file()
        first = self._next_window
        n = min(self.batch_size, self._n_windows - first)
        self._next_window += n
        start = t.arange(first, first + n, dtype=t.int32)
        if self.shuffle:
            start = start[t.randperm(n)]
        return gather_windows(self._frames, start.to(self.device), crop=self.crop, steps=self.time_steps,
                              normalizing_max=self.normalizing_max, power=float(self.power), dtype=self.dtype)

    def __iter__(self):
        return self


def get_loaders(train_batch_size: int, test_batch_size: int, data_folder: str, device, crop: int = None,
                shuffle: bool = True, merge_nodes: bool = False):
    """reference :130-166 (val and test loaders both read the ``test`` folder there too)."""
    mk = lambda bs, sub: DataLoader(bs, os.path.join(data_folder, sub), device, crop=crop, shuffle=shuffle,
                                    merge_nodes=merge_nodes)
    return mk(train_batch_size, "train"), mk(test_batch_size, "test"), mk(test_batch_size, "test")
