"""The model registry and metric helpers of ``convolutional_gat/utils.py`` (reference :13-22, :128-167); plotting and
visualisation there are out of scope.  ``get_metrics`` keeps the reference signature and return values but counts on
the device (``cgat_val_metrics``) instead of cloning both tensors to the CPU."""
import torch as t

from cgat import _lib

from .GAT3D.GATMultistream import Model as GatModel
from .unet_model import UnetModel

model_classes = {
    "unet": UnetModel,
    "temporal": GatModel,
    "spatial": GatModel,
    "multi_stream": GatModel,
}


def get_number_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def update_history(history, data):  # reference :128-132
    for key, val in data.items():
        if key not in history:
            history[key] = []
        history[key].append(val)


def val_batch_sums(y, y_hat, threshold, *, power=1.0, normalizing_max=254.0):
    """One pass over a batch: ``[sum sq err, sum denormalised sq err, TP, FP, FN, #equal]`` (float64, device)."""
    _lib.require_cuda(y, y_hat)
    y, y_hat = y.contiguous(), y_hat.contiguous().to(y.dtype)
    out = t.zeros(6, dtype=t.float64, device=y.device)
    _lib.call("cgat_val_metrics", _lib.ptr(y), _lib.ptr(y_hat), y.numel(), float(power), float(threshold),
              float(normalizing_max), _lib.dtype_tag(y), _lib.ptr(out), _lib.stream())
    return out


def get_metrics(y, y_hat, mean):
    """``(acc, prec, rec)`` of reference :135-167: accuracy = #equal / numel of ONE sample, precision and recall
    multiplied by the batch length (the caller divides by the number of samples seen, train.py:86-88)."""
    s = val_batch_sums(y, y_hat, mean).cpu()
    tp, fp, fn, eq = s[2], s[3], s[4], s[5]
    n = len(y)
    acc = (eq / y[0].numel()).float()
    prec = ((tp / (tp + fp)) * n).float()
    rec = ((tp / (tp + fn)) * n).float()
    return acc, prec, rec
