"""The adversarial step of ``dcgan/train.py`` (reference :97-160) on the B200 nets of ``dcgan/model.py``.

Host logic mirrors the reference line by line (same variable names, same order of backward passes and optimiser
steps, BCELoss criterion :224, Adam(lr=2e-4, betas=(0.5, 0.999)) :227-236); every convolution of the three nets --
forward, dgrad and wgrad, nine passes through the discriminators and three through the generator per step -- runs
in the CUDA conv kernels (tcgen05 implicit GEMM where ``cgat_conv_tc_supported``, the direct kernel otherwise).
Metrics, printing, checkpointing and the data loader of the reference loop are out of scope.
"""
import torch as t
from torch import nn, optim


def make_optimizers(netG, netFD, netTD, lr=0.0002, beta1=0.5, capturable=False):
    """``optim.Adam(net.parameters(), lr=params["lr"], betas=(params["beta1"], 0.999))`` (reference :227-236).
    ``capturable=True`` keeps the step counters on the device so the step can live in a CUDA graph."""
    def mk(net):
        params = list(net.parameters())
        # graph-resident steps on the GPU use torch's FUSED Adam (one multi-tensor kernel per step instead of ~6 foreach
        # kernels: 0.35 -> 0.06 ms of the 5 ms step for the three optimisers); same update rule
        fused = bool(capturable and params and all(p.is_cuda for p in params))
        return optim.Adam(params, lr=lr, betas=(beta1, 0.999), capturable=capturable, **({"fused": True} if fused else {}))

    return mk(netG), mk(netFD), mk(netTD)


def adversarial_step(*, netG, netFD, netTD, optimizerG, optimizerFD, optimizerTD, criterion, x, y):
    """One batch of ``train_single_epoch`` (reference :97-160).  ``x, y`` are ``[N, nc, 64, 64]`` (the reference
    squeezes a singleton dim 2 first, :98-99).  Returns ``(errFD, errTD, errG, fake_data)`` as device tensors."""
    data = x
    b_size = data.size(0)
    device = data.device
    netTD.zero_grad()
    netFD.zero_grad()
    real_label = t.zeros(b_size, device=device) + 1
    fake_label = t.zeros(b_size, device=device)

    pred_real_frame_label = netFD(y)
    pred_real_temp_label = netTD(t.cat((data, y), dim=1))
    errFD_real = criterion(pred_real_frame_label.float(), real_label)
    errTD_real = criterion(pred_real_temp_label.float(), real_label)
    errFD_real.backward()
    errTD_real.backward()

    fake_data = netG(data)
    fake_data_detached = fake_data.detach()
    pred_fake_frame_label = netFD(fake_data_detached)
    pred_fake_temp_label = netTD(t.cat((data, fake_data_detached), dim=1))
    errFD_fake = criterion(pred_fake_frame_label.float(), fake_label)
    errTD_fake = criterion(pred_fake_temp_label.float(), fake_label)
    errFD_fake.backward()
    errTD_fake.backward()

    errFD = errFD_real + errFD_fake
    errTD = errTD_real + errTD_fake
    optimizerFD.step()
    optimizerTD.step()

    netG.zero_grad()
    pred_frame_label = netFD(fake_data).view(-1)
    pred_temp_label = netTD(t.cat((data, fake_data), dim=1)).view(-1)
    errG = criterion(pred_frame_label.float(), real_label) + criterion(pred_temp_label.float(), real_label)
    errG.backward()
    optimizerG.step()
    return errFD.detach(), errTD.detach(), errG.detach(), fake_data.detach()


class GraphedAdversarialStep:
    """``adversarial_step`` captured once in a CUDA graph and replayed per batch.

    The eager step issues ~1 500 small launches (nine discriminator and three generator passes, BatchNorm, dropout,
    three Adam steps); with the convs on tensor cores the GPU finishes them faster than Python can enqueue them, so
    the step is host-bound.  Replay removes the host from the loop; the arithmetic is the same launches in the same
    order.  Inputs are copied into static buffers; the returned tensors are the graph's static outputs (clone them
    to keep a value across steps).  Optimisers must be built with ``make_optimizers(..., capturable=True)`` and be
    fresh (no steps taken): construction runs warm-up steps on the example batch and then puts the nets and the Adam
    state back where they were."""

    def __init__(self, *, netG, netFD, netTD, optimizerG, optimizerFD, optimizerTD, criterion, x, y, warmup=3):
        self.x, self.y = x.clone(), y.clone()
        kw = dict(netG=netG, netFD=netFD, netTD=netTD, optimizerG=optimizerG, optimizerFD=optimizerFD,
                  optimizerTD=optimizerTD, criterion=criterion, x=self.x, y=self.y)
        # the warm-up steps (they allocate gradients and optimiser state before capture) must not count as training:
        # parameters and BatchNorm statistics are restored and the Adam state is reset IN PLACE afterwards (the graph
        # holds the addresses), so the first replay is the first optimisation step from the caller's state
        nets, opts = (netG, netFD, netTD), (optimizerG, optimizerFD, optimizerTD)
        saved = [(v, v.clone()) for net in nets for v in net.state_dict().values()]
        side = t.cuda.Stream()
        side.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(side):
            for _ in range(max(1, warmup)):
                adversarial_step(**kw)
        t.cuda.current_stream().wait_stream(side)
        self.graph = t.cuda.CUDAGraph()
        with t.cuda.graph(self.graph):
            self.out = adversarial_step(**kw)
        with t.no_grad():
            for v, c in saved:
                v.copy_(c)
            for opt in opts:
                for st in opt.state.values():
                    for v in st.values():
                        if t.is_tensor(v):
                            v.zero_()

    def __call__(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.out


def default_criterion():
    return nn.BCELoss()  # reference :224
