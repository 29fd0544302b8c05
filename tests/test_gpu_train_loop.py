"""GPU parity of the train / validation loop pieces: the validation loop and
metrics against the reference's own ``test`` / ``get_metrics`` functions (tests/golden/val_metrics.pt, compiled from
the reference source by oracle/make_golden.py), and ``train()`` end to end on a synthetic KNMI folder."""
import json
import os

import pytest
import torch

from util import close, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _AffineModel(torch.nn.Module):
    def forward(self, x):
        return 0.9 * x + 0.02


@pytest.mark.parametrize("tag,power", [("p1", 1.0), ("p05", 0.5)])
def test_validation_loop_matches_reference(tag, power):
    from convolutional_gat.train import test as run_test
    from convolutional_gat.utils import get_metrics

    fx = golden("val_metrics")

    class Loader(list):
        pass

    loader = Loader([(fx[f"{tag}.x{i}"].to(DEV), fx[f"{tag}.y{i}"].to(DEV)) for i in range(3)])
    loader.power = torch.tensor(power)
    loader.normalizing_max = 254
    res = run_test(_AffineModel(), DEV, loader)
    for k in ("val_loss", "val_acc", "val_prec", "val_rec", "val_denorm_mse"):
        close(torch.tensor(res[k]), fx[f"{tag}.{k}"], rtol=2e-5, atol=1e-7, msg=k)
    acc, prec, rec = get_metrics(loader[0][1], _AffineModel()(loader[0][0]), 0.05)
    close(torch.stack([acc, prec, rec]), fx[f"{tag}.gm"], rtol=1e-6, atol=1e-7, msg="get_metrics")


def test_train_function_end_to_end(tmp_path):
    """``train(**cfg)`` (train.py:158-261) on a synthetic KNMI folder: history keys, history.json, best model.pt, a loss
    that goes down, short last batches and batches of size 1 handled."""
    from convolutional_gat.train import train

    g = torch.Generator().manual_seed(5)
    base = torch.rand(1, 6, 16, 16, generator=g)
    for split, nfiles in (("train", 2), ("test", 1)):
        os.makedirs(tmp_path / "data" / split)
        for i in range(nfiles):
            L = 24 if split == "train" else 16
            frames = (base + 0.05 * torch.rand(L, 6, 16, 16, generator=g)).clamp(0, 1)
            torch.save((frames * 254).round().to(torch.int64), tmp_path / "data" / split / f"{i:010d}.pt")
    out = tmp_path / "out"
    os.makedirs(out)
    torch.manual_seed(369)
    hist = train(model_type="temporal", optimizer=torch.optim.Adam, mapping_type="conv", output_path=str(out),
                 train_batch_size=8, test_batch_size=4, epochs=3, learning_rate=1e-2, lr_step=1, gamma=0.9,
                 downsample_size=(16, 16), preprocessed_folder=str(tmp_path / "data"), dataset="kmni")
    assert set(hist) == {"train_loss", "val_loss", "val_acc", "val_prec", "val_rec", "val_denorm_mse"}
    assert len(hist["train_loss"]) == 3 and len(hist["val_loss"]) == 3
    assert hist["train_loss"][-1] < hist["train_loss"][0]
    assert json.load(open(out / "history.json")) == hist
    sd = torch.load(out / "model.pt")
    assert any(k.endswith("conv.weight") for k in sd)
