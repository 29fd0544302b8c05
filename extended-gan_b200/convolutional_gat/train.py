"""``test`` / ``train_single_epoch`` / ``train`` of ``convolutional_gat/train.py`` (reference :28-265) over the B200 path.

Same signatures, history keys (``train_loss``, ``val_loss``, ``val_acc``, ``val_prec``, ``val_rec``,
``val_denorm_mse``), ``history.json`` and best-``model.pt`` behaviour, batches of size 1 skipped (:52, :127), StepLR /
ReduceLROnPlateau choice (:212-221).  Differences: the step runs as ``cgat.train_step.TrainStep`` (one CUDA graph +
fused Adam; the learning rate is re-read from the scheduler every epoch), the validation sums are one kernel per batch
(``cgat_val_metrics``) instead of CPU clones, loaders come from our ``data_loaders`` package, and the plotting /
visualisation calls of the reference (:245-258) are dropped (out of scope).
"""
import json
import os

import torch as t
import torch.nn as nn

from cgat.train_step import TrainStep
from .data_loaders.get_loaders import get_loaders
from .utils import get_number_parameters, model_classes, update_history, val_batch_sums


def test(model: nn.Module, device, loader, flag="val"):
    """reference :28-91.  Returns the same dict of per-sample averages."""
    model.eval()
    sums = t.zeros(5, dtype=t.float64)
    total_length = 0
    power = float(getattr(loader, "power", 1.0))
    nmax = float(getattr(loader, "normalizing_max", 254))
    with t.no_grad():
        for x, y in loader:
            if len(x) > 1:
                y_hat = model(x)
                yp = t.pow(y, 1 / power) if power != 1.0 else y
                unique = t.unique(yp)  # :60-61: the median of the distinct target values is the threshold
                threshold = unique[int(len(unique) * (1 / 2))].item()
                s = val_batch_sums(y, y_hat, threshold, power=power, normalizing_max=nmax).cpu()
                per_sample = y[0].numel()
                n = len(x)
                total_length += n
                prec = (s[2] / (s[2] + s[3])) * n
                rec = (s[2] / (s[2] + s[4])) * n
                sums += t.stack([s[0] / per_sample, s[5] / per_sample, t.nan_to_num(prec, nan=0.0),
                                 t.nan_to_num(rec, nan=0.0), s[1] / per_sample])
    model.train()
    avg = (sums / max(total_length, 1)).tolist()
    return {"val_loss": avg[0], "val_acc": avg[1], "val_prec": avg[2], "val_rec": avg[3], "val_denorm_mse": avg[4]}


def train_single_epoch(epoch, optimizer, criterion, scheduler, model, train_batch_size, test_batch_size,
                       preprocessed_folder, device, dataset, downsample_size, history, output_path, *, step=None):
    """reference :94-155.  ``optimizer`` is the ``TrainStep`` holder's torch optimiser shim (lr source); ``step`` the
    ``TrainStep`` that owns parameters, gradients and Adam state."""
    train_loader, val_loader, test_loader = get_loaders(
        train_batch_size=train_batch_size, test_batch_size=test_batch_size, preprocessed_folder=preprocessed_folder,
        device=device, dataset=dataset, downsample_size=downsample_size, merge_nodes=False)
    model.train()
    print(f"\nEpoch: {epoch}")
    running = t.zeros(1, dtype=t.float64, device=device)
    total_length = 0
    for param_group in optimizer.param_groups:
        print(f"LR: {param_group['lr']}")
        step.lr = param_group["lr"]
    for x, y in train_loader:
        if len(x) > 1:
            step.step(x.to(step.x.dtype), y.to(step.y.dtype))  # forward, loss, backward, Adam (:129-133)
            total_length += len(x)
            running += step.mse.double() * len(x)  # sum((y_hat-y)^2)/numel(sample) of :135-139 = MSE * N
    train_loss = (running / max(total_length, 1)).item()
    print(f"Train loss: {round(train_loss, 6)}")
    history["train_loss"].append(train_loss)
    test_result = test(model, device, val_loader)
    scheduler.step(test_result["val_loss"]) if isinstance(scheduler, t.optim.lr_scheduler.ReduceLROnPlateau) \
        else scheduler.step()
    print(json.dumps(test_result, indent=4))
    update_history(history, test_result)
    with open(os.path.join(output_path, "history.json"), "w") as f:
        json.dump(history, f, indent=4)
    if (len(history["val_loss"]) == 1) or test_result["val_loss"] < min(history["val_loss"][:-1]):
        print("Saving model.")
        t.save(model.state_dict(), os.path.join(output_path, "model.pt"))
    return step


def train(*, model_type, optimizer, mapping_type, output_path, train_batch_size, test_batch_size, epochs, learning_rate,
          lr_step, gamma, plot=True, criterion=nn.MSELoss(), downsample_size=(256, 256), preprocessed_folder="",
          dataset="kmni", test_first=False, reduce_lr_on_plateau=False, dtype=t.bfloat16):
    """reference :158-261 (``optimizer`` must be ``torch.optim.Adam``: the step kernel implements Adam, :212)."""
    if not t.cuda.is_available():
        raise RuntimeError("the conv-GAT path needs a CUDA device (no CPU fallback)")
    if optimizer is not t.optim.Adam:
        raise NotImplementedError("the fused step implements torch.optim.Adam (convolutional_gat/train.py:212)")
    device = t.device("cuda")
    history = {"train_loss": []}
    train_loader, val_loader, test_loader = get_loaders(
        train_batch_size=train_batch_size, test_batch_size=test_batch_size, preprocessed_folder=preprocessed_folder,
        device=device, dataset=dataset, downsample_size=downsample_size, merge_nodes=False)
    for x, y in val_loader:
        _, image_width, image_height, steps, n_vertices = x.shape
        break
    model = model_classes[model_type](image_width=image_width, image_height=image_height, n_vertices=n_vertices,
                                      attention_type=model_type, mapping_type=mapping_type).to(device)
    print(f"Number of parameters: {get_number_parameters(model)}")
    print(f"Using mapping: {model.mapping_type}")
    opt = optimizer(model.parameters(), lr=learning_rate, weight_decay=0.01)  # lr / scheduler bookkeeping only
    if not reduce_lr_on_plateau:
        scheduler = t.optim.lr_scheduler.StepLR(opt, step_size=lr_step, gamma=gamma)
    else:
        scheduler = t.optim.lr_scheduler.ReduceLROnPlateau(opt, "min", patience=0, factor=0.5)
    x0, y0 = next(iter(train_loader))
    step = TrainStep(model, x0.to(dtype), y0.to(dtype), lr=learning_rate, weight_decay=0.01)
    if test_first:
        result = test(model, device, train_loader)
        history["train_loss"].append(result["val_loss"])
        result = test(model, device, test_loader)
        update_history(history, result)
    for epoch in range(1, epochs + 1):
        step = train_single_epoch(epoch, opt, criterion, scheduler, model, train_batch_size, test_batch_size,
                                  preprocessed_folder, device, dataset, downsample_size, history, output_path, step=step)
    return history
