"""Loss / utility surface of the reference's util.py with the same signatures and error behaviour.

The fused training iteration (step.TrainStep) does not go through these functions; they exist so code
written against the reference (`from util import ...`) keeps working on the drop-in modules."""
import os
import random

import numpy as np
import torch
from torch.nn import init

from . import _C
from ._C import F, LL, ptr

# reference util.py:8-11 seeds every RNG at import
torch.manual_seed(21)
if torch.cuda.is_available():
    torch.cuda.manual_seed_all(21)
random.seed(21)
np.random.seed(21)


def set_requires_grad(nets, requires_grad=False):
    """reference util.py:14-20"""
    for net in nets if isinstance(nets, list) else [nets]:
        if net is not None:
            for p in net.parameters():
                p.requires_grad = requires_grad


def init_weights(net, init_type='normal', gain=0.02):
    """reference util.py:23-34: N(0, gain) on Conv*/Linear weights, zero biases, BatchNorm2d N(1, gain)."""
    def visit(m):
        name = m.__class__.__name__
        if hasattr(m, 'weight') and m.weight is not None and ('Conv' in name or 'Linear' in name):
            init.normal_(m.weight.data, 0.0, gain)
            if getattr(m, 'bias', None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif 'BatchNorm2d' in name:
            init.normal_(m.weight.data, 1.0, gain)
            init.constant_(m.bias.data, 0.0)

    net.apply(visit)


def mkdir(path):
    if not os.path.exists(path):
        os.makedirs(path)


def _l1_mean(a, b):
    out = torch.zeros(1, device=a.device)
    a, b = a.detach().contiguous().float(), b.detach().contiguous().float()
    _C.call("l1_loss", ptr(a), ptr(b), LL(a.numel()), F(1.0), ptr(out), None)
    return out[0]


def pan_loss(real_features, fake_features, mode='normal', loss_type='l1', weights=[1, 1, 1, 1]):
    """reference util.py:41-70. The features the discriminator hands out are detached, so the value
    carries no gradient (as in the reference); 'normal'/'l1' -- the combination train.py uses -- runs
    on the tg_l1_loss kernel, the unused gram / l2 variants on plain tensor ops."""
    if mode not in ('normal', 'gram'):
        raise ValueError('mode must be normal or gram')
    if loss_type not in ('l1', 'l2'):
        raise ValueError('loss_type must be l1 or l2')
    if len(weights) != 4:
        raise ValueError('weights must be a list of 4 numbers')
    w = np.array(weights) / np.sum(weights)
    total = 0.0
    for i in range(4):
        r, f = real_features[i], fake_features[i]
        if mode == 'gram':
            r = r.reshape(r.shape[0], r.shape[1], -1)
            f = f.reshape(f.shape[0], f.shape[1], -1)
            r, f = r @ r.permute(0, 2, 1), f @ f.permute(0, 2, 1)
        if loss_type == 'l1' and r.is_cuda and not r.requires_grad and not f.requires_grad:
            lo = _l1_mean(r, f)
        elif loss_type == 'l1':
            lo = torch.nn.functional.l1_loss(r, f)
        else:
            lo = torch.nn.functional.mse_loss(r, f)
        total = total + lo * w[i]
    return total


class _GradientPenaltyFn(torch.autograd.Function):
    """Penalty value + its parameter gradients in one fused pass (forward, input-gradient backward and
    the double backward all run inside PatchDInstance); backward() scales the stored gradients."""

    @staticmethod
    def forward(ctx, inst, out, nparams, *params):
        st = inst.store
        ctx.grads = [st.grad_as_torch(i) for i in range(nparams)]   # snapshot: the arena is shared
        return out.clone()

    @staticmethod
    def backward(ctx, g):
        return (None, None, None, *[x * g for x in ctx.grads])


def gradient_penalty(disc, real_img, real_mask, fake_mask, device, ver=2, type='mixed', constant=1.0,
                     lambda_gp=1.0):
    """reference util.py:72-97: lambda * mean_n (|| d disc(real_img, x_n) / d x_n + 1e-16 ||_2 - constant)^2."""
    if lambda_gp <= 0.0:
        return 0.0
    from .bridge import disc_instance
    n, _, h, w = real_mask.shape
    a = real_img.detach().contiguous().float()
    rm, fm = real_mask.detach().contiguous().float(), fake_mask.detach().contiguous().float()
    inst = disc_instance(disc, n, h, w, True, True, "gp")
    if type == 'real':
        inst.pack_input(a, rm)
    elif type == 'fake':
        inst.pack_input(a, fm)
    elif type == 'mixed':
        alpha = torch.rand(n, 1, device=device)
        if ver == 2:
            alpha = (alpha + 1) / 2
        al = alpha.view(-1).float().contiguous()
        inst.pack_input(a, rm, wa=al, b2=fm, wb=(1 - al).contiguous())
    else:
        raise NotImplementedError(f'{type} not implemented')
    out = torch.zeros(1, device=a.device)
    inst.store.zero_grad()
    inst.forward()
    inst.gp_first_backward()
    inst.gp_penalty(real_img.shape[1], real_mask.shape[1], float(lambda_gp), float(constant), out)
    inst.gp_second_backward()
    params = list(disc.parameters())
    return _GradientPenaltyFn.apply(inst, out, len(params), *params)[0]
