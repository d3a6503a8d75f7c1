"""Device-side input pipeline: what the reference's DataLoader workers do per sample on the CPU
(datasets/PairedDataset.py:30-44,80-92) -- albumentations HorizontalFlip(p=.5) + Affine(translate_percent=.1,
scale=(.8,1.2), rotate=(-15,15), p=.5), then ToTensor (+ Normalize(.5,.5) on the source) -- done for the whole batch
by one kernel (tg_augment_pair) on uint8 HWC tensors that crossed PCIe at a quarter of the fp32 size.

albumentations is a third-party, unpinned dependency that is not installed here, so its exact sampling order and
matrix conventions cannot be checked: PARITY UNPINNED against it. The transform itself is fully specified below
(and restated in oracle.augment_pair): parameters are drawn on the host, the warp is the inverse map in 16.16 fixed
point, so which source pixels every output pixel reads is exact integer arithmetic on both sides."""
import math

import torch

from . import _C

FIX = 1 << 16


def sample_params(n, h, w, generator=None, p_flip=0.5, p_affine=0.5, translate=0.1, scale=(0.8, 1.2),
                  rotate=(-15.0, 15.0)):
    """(n, 8) int64: {flip, a00, a01, a02, a10, a11, a12, 0}. The forward transform is
    T = Translate(c + t) . Rotate(theta) . Scale(sx, sy) . Translate(-c) about the image centre c = ((w-1)/2, (h-1)/2),
    t uniform in +-translate * (w, h), sx / sy independent in `scale`, theta in `rotate` degrees; output(x) =
    input(T^-1 x). Each sample applies the flip with p_flip and the affine with p_affine (else identity)."""
    u = torch.rand(n, 7, generator=generator, dtype=torch.float64)
    out = torch.zeros(n, 8, dtype=torch.int64)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    for i in range(n):
        flip = u[i, 0].item() < p_flip
        if u[i, 1].item() < p_affine:
            tx = (2 * u[i, 2].item() - 1) * translate * w
            ty = (2 * u[i, 3].item() - 1) * translate * h
            sx = scale[0] + u[i, 4].item() * (scale[1] - scale[0])
            sy = scale[0] + u[i, 5].item() * (scale[1] - scale[0])
            th = math.radians(rotate[0] + u[i, 6].item() * (rotate[1] - rotate[0]))
            c, s = math.cos(th), math.sin(th)
            # T^-1 = Translate(c) . Scale^-1 . Rotate(-theta) . Translate(-(c + t))
            a00, a01, a10, a11 = c / sx, s / sx, -s / sy, c / sy
            a02 = cx - (a00 * (cx + tx) + a01 * (cy + ty))
            a12 = cy - (a10 * (cx + tx) + a11 * (cy + ty))
        else:
            a00, a01, a02, a10, a11, a12 = 1.0, 0.0, 0.0, 0.0, 1.0, 0.0
        out[i] = torch.tensor([int(flip)] + [int(round(v * FIX)) for v in (a00, a01, a02, a10, a11, a12)] + [0])
    return out


def identity_params(n):
    out = torch.zeros(n, 8, dtype=torch.int64)
    out[:, 1] = FIX
    out[:, 5] = FIX
    return out


def augment_pair(img_u8, mask_u8, params):
    """img_u8 (N,H,W,ca) / mask_u8 (N,H,W,cb) uint8 on the device, params (N,8) int64 ->
    (real_A fp32 NCHW in [-1,1], real_B fp32 NCHW in [0,1])."""
    if not (img_u8.is_cuda and mask_u8.is_cuda):
        raise _C.TgError("the device-side input pipeline runs on CUDA (sm_100a); there is no CPU fallback")
    assert img_u8.dtype == torch.uint8 and mask_u8.dtype == torch.uint8 and img_u8.dim() == 4
    n, h, w, ca = img_u8.shape
    cb = mask_u8.shape[3]
    assert mask_u8.shape[:3] == (n, h, w) and params.shape == (n, 8)
    img_u8, mask_u8 = img_u8.contiguous(), mask_u8.contiguous()
    q = params.to(img_u8.device, torch.int64).contiguous()
    a = torch.empty(n, ca, h, w, device=img_u8.device)
    b = torch.empty(n, cb, h, w, device=img_u8.device)
    _C.call("augment_pair", _C.ptr(img_u8), _C.ptr(mask_u8), _C.ptr(q), _C.ptr(a), _C.ptr(b), n, h, w, ca, cb)
    return a, b
