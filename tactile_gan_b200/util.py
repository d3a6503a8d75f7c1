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


class _VGGLossFn(torch.autograd.Function):
    """Value and input-gradients of the perceptual term in one fused pass through engine.VGGFeatEngine."""

    @staticmethod
    def forward(ctx, inp, target, owner, weights, feature_layers):
        n, c, h, w = inp.shape
        e_in = owner._engine(n, c, h, w, "input", inp.requires_grad)
        e_tg = owner._engine(n, c, h, w, "target", target.requires_grad)
        xi, xt = inp.detach().contiguous().float(), target.detach().contiguous().float()
        e_in.forward(xi)
        e_tg.forward(xt)
        out = torch.zeros(1, device=inp.device)
        spare = torch.zeros(1, device=inp.device)
        gi = gt = None
        counted = False
        if target.requires_grad:
            e_tg.loss_and_seed(e_in, weights, 1.0, out, feature_layers)
            gt = e_tg.backward(torch.zeros(n, c, h, w, device=inp.device))
            counted = True
        if inp.requires_grad:
            e_in.loss_and_seed(e_tg, weights, 1.0, spare if counted else out, feature_layers)
            gi = e_in.backward(torch.zeros(n, c, h, w, device=inp.device))
            counted = True
        if not counted:
            e_tg.loss_and_seed(e_in, weights, 1.0, out, feature_layers)
        ctx.grads = (gi, gt)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        gi, gt = ctx.grads
        return (None if gi is None else gi * g, None if gt is None else gt * g, None, None, None)


class VGGPerceptualLoss(torch.nn.Module):
    """reference util.py:100-144 (the --version 1 perceptual term): L1 distances between the outputs of four
    frozen VGG16 slices. Same constructor / forward signature; the slices run on the sm_100a engine. The
    reference loads torchvision's ImageNet weights; without network access they cannot be downloaded, so
    the slices fall back to torchvision's random init (load a state_dict with keys blocks.<b>.<i>.* to
    use real weights). The gram / style branch (style_layers, unused by train.py) is not built."""

    def __init__(self, resize=True):
        super().__init__()
        import torchvision
        try:
            feats = torchvision.models.vgg16(weights=torchvision.models.VGG16_Weights.IMAGENET1K_V1).features
        except Exception as e:   # no network / no cached checkpoint
            import warnings
            warnings.warn(f"VGG16 ImageNet weights unavailable ({type(e).__name__}); using random-init VGG16")
            feats = torchvision.models.vgg16(weights=None).features
        blocks = [feats[:4], feats[4:9], feats[9:16], feats[16:23]]
        for bl in blocks:
            for p in bl.parameters():
                p.requires_grad = False
        self.blocks = torch.nn.ModuleList(blocks).eval()
        self.resize = resize
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        if torch.cuda.is_available():
            self.cuda()

    def _engine(self, n, c, h, w, role, backward):
        from .engine import VGGFeatEngine
        cache = self.__dict__.setdefault("_tg_engines", {})
        key = (n, c, h, w, role)
        if key not in cache or (backward and not cache[key].with_backward):
            cache[key] = VGGFeatEngine(self.blocks, n, h, w, src_channels=c, resize=self.resize, backward=backward)
        return cache[key]

    def forward(self, input, target, feature_layers=[0, 1, 2, 3], style_layers=[], weights=[0.25, 0.25, 0.25, 0.25]):
        if len(style_layers):
            raise NotImplementedError("the gram / style branch of VGGPerceptualLoss is not built (unused by train.py)")
        if not input.is_cuda:
            raise _C.TgError("tactile_gan_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if input.shape[1] not in (1, 3):
            raise ValueError("VGGPerceptualLoss expects 1- or 3-channel images")
        return _VGGLossFn.apply(input, target, self, [float(v) for v in weights], tuple(feature_layers))
