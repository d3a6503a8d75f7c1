"""Input pipeline against albumentations' DOCUMENTED semantics for the reference's transform list
(datasets/PairedDataset.py:30-44): HorizontalFlip(p=.5) then Affine(translate_percent=.1, scale=(.8, 1.2),
rotate=(-15, 15), fit_output=False, p=.5) with the library defaults -- image cv2.INTER_LINEAR, mask
cv2.INTER_NEAREST, BORDER_CONSTANT with fill 0, independent x / y scales (keep_ratio=False), independent x / y
translations in +-10 % of width / height, the transform taken about the image centre ((w-1)/2, (h-1)/2) as
Translate . Rotate . Scale. albumentations is unpinned in requirements.txt and not installed here, so:

  * CPU (always): the sampler is tested statistically -- rates, ranges, uniformity (Kolmogorov-Smirnov) and
    independence of the decoded parameters -- and the fixed-point matrices are decoded back to those parameters;
  * GPU (always): the warp kernel against an INDEPENDENT float implementation of the documented resampling
    (torch grid_sample, bilinear / nearest, zero padding), not against the repo's own fixed-point restatement;
  * GPU + albumentations importable: fixed-parameter A.Affine / A.HorizontalFlip outputs against the kernel.
"""
import math

import numpy as np
import pytest
import torch

from tactile_gan_b200.augment import FIX, identity_params, sample_params


def decode(q, h, w):
    """(flip, applied, sx, sy, theta_deg, tx / w, ty / h) from the 16.16 inverse-map rows of sample_params."""
    a = q[:, 1:7].double() / FIX
    a00, a01, a02, a10, a11, a12 = a.unbind(1)
    sx, sy = 1 / torch.sqrt(a00 ** 2 + a01 ** 2), 1 / torch.sqrt(a10 ** 2 + a11 ** 2)
    th = torch.atan2(a01, a00)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    # a02 = cx - (a00 (cx+tx) + a01 (cy+ty)); a12 = cy - (a10 (cx+tx) + a11 (cy+ty))  -> solve for (cx+tx, cy+ty)
    m = torch.stack([torch.stack([a00, a01], 1), torch.stack([a10, a11], 1)], 1)
    rhs = torch.stack([cx - a02, cy - a12], 1).unsqueeze(2)
    ct = torch.linalg.solve(m, rhs).squeeze(2)
    applied = (q[:, 1:7] != identity_params(1)[0, 1:7]).any(1)
    return q[:, 0].bool(), applied, sx, sy, torch.rad2deg(th), (ct[:, 0] - cx) / w, (ct[:, 1] - cy) / h


def ks_uniform(x, lo, hi):
    """Kolmogorov-Smirnov distance of samples x to U[lo, hi]."""
    x = np.sort((np.asarray(x) - lo) / (hi - lo))
    n = len(x)
    return max(np.max(np.arange(1, n + 1) / n - x), np.max(x - np.arange(0, n) / n))


def test_sampler_follows_the_documented_parameter_distributions():
    n, h, w = 6000, 256, 256
    q = sample_params(n, h, w, generator=torch.Generator().manual_seed(123))
    flip, applied, sx, sy, th, tx, ty = decode(q, h, w)
    sig = 0.5 / math.sqrt(n)
    assert abs(flip.float().mean().item() - 0.5) < 4 * sig          # HorizontalFlip(p=0.5)
    assert abs(applied.float().mean().item() - 0.5) < 4 * sig       # Affine(p=0.5)
    # flip and affine are drawn independently (Compose applies each transform with its own p)
    both = (flip & applied).float().mean().item()
    assert abs(both - 0.25) < 4 * math.sqrt(0.25 * 0.75 / n)
    k = int(applied.sum())
    crit = 1.95 / math.sqrt(k)                                      # KS critical value at alpha ~ 0.001
    sel = lambda t: t[applied].numpy()
    for name, v, lo, hi in (("scale_x", sel(sx), 0.8, 1.2), ("scale_y", sel(sy), 0.8, 1.2),
                            ("rotate", sel(th), -15.0, 15.0), ("translate_x", sel(tx), -0.1, 0.1),
                            ("translate_y", sel(ty), -0.1, 0.1)):
        assert v.min() >= lo - 1e-3 and v.max() <= hi + 1e-3, name  # documented ranges (16.16 rounding slack)
        assert ks_uniform(v, lo, hi) < crit, (name, ks_uniform(v, lo, hi), crit)
    # keep_ratio=False: x and y scales independent; translations independent; nothing tied to the rotation
    cols = np.stack([sel(sx), sel(sy), sel(th), sel(tx), sel(ty)])
    cc = np.corrcoef(cols)
    assert np.abs(cc - np.eye(5)).max() < 4 / math.sqrt(k)
    # samples without the affine are the exact identity map
    assert torch.equal(q[~applied][:, 1:7], identity_params(int((~applied).sum()))[:, 1:7])


def _float_reference(img_u8, mask_u8, q):
    """Documented resampling, in float: output(x) = input(T^-1 x) -- image bilinear, mask nearest, zeros outside
    (cv2.warpAffine with BORDER_CONSTANT 0), the horizontal flip applied to the input first; then ToTensor (+
    Normalize(.5, .5) on the source image, PairedDataset.py:52-58,86)."""
    import torch.nn.functional as F
    n, h, w, _ = img_u8.shape
    a = q[:, 1:7].double() / FIX
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
    outs_a, outs_b = [], []
    for i in range(n):
        sx = a[i, 0] * xs + a[i, 1] * ys + a[i, 2]
        sy = a[i, 3] * xs + a[i, 4] * ys + a[i, 5]
        if q[i, 0]:
            sx = (w - 1) - sx
        grid = torch.stack([2 * sx / (w - 1) - 1, 2 * sy / (h - 1) - 1], -1).unsqueeze(0)
        im = img_u8[i].permute(2, 0, 1).unsqueeze(0).double()
        mk = mask_u8[i].permute(2, 0, 1).unsqueeze(0).double()
        oa = F.grid_sample(im, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0] / 255
        ob = F.grid_sample(mk, grid, mode="nearest", padding_mode="zeros", align_corners=True)[0] / 255
        outs_a.append(((oa - 0.5) / 0.5).float())
        outs_b.append(ob.float())
    return torch.stack(outs_a), torch.stack(outs_b)


@pytest.mark.gpu
def test_warp_kernel_matches_an_independent_float_resampler():
    from tactile_gan_b200.augment import augment_pair
    g = torch.Generator().manual_seed(5)
    n, h, w = 6, 96, 128
    # smooth image (so sub-pixel errors show up as small value errors) + blocky mask (so nearest ties are rare)
    img = (torch.rand(n, h // 8, w // 8, 3, generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2) * 255).byte()
    mask = (torch.rand(n, h // 4, w // 4, 3, generator=g).repeat_interleave(4, 1).repeat_interleave(4, 2) > 0.5).byte() * 255
    q = sample_params(n, h, w, generator=g, p_flip=0.5, p_affine=1.0)
    q[0, 0], q[1, 0] = 0, 1
    ra, rb = _float_reference(img, mask, q)
    a, b = augment_pair(img.cuda(), mask.cuda(), q)
    a, b = a.cpu(), b.cpu()
    # 16.16 fixed-point coordinates: <= 2^-16 * (h + w) pixels off -> far below one grey level on an 8-bit image
    assert (a - ra).abs().max().item() < 2.0 / 255 * 2
    assert (a - ra).abs().mean().item() < 1e-4
    # nearest: identical except where the float coordinate sits within 2^-15 of a rounding tie or the border
    assert (b != rb).float().mean().item() < 2e-3


@pytest.mark.gpu
def test_against_albumentations_when_installed():
    A = pytest.importorskip("albumentations")
    from tactile_gan_b200.augment import augment_pair
    g = torch.Generator().manual_seed(9)
    h = w = 128
    img = (torch.rand(1, h // 8, w // 8, 3, generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2) * 255).byte()
    mask = (torch.rand(1, h // 4, w // 4, 3, generator=g).repeat_interleave(4, 1).repeat_interleave(4, 2) > 0.5).byte() * 255
    for s, r, t, flip in ((1.0, 0.0, 0.0, True), (1.1, 0.0, 0.0, False), (1.0, 10.0, 0.0, False), (0.9, -12.0, 0.05, True)):
        tf = A.Compose([A.HorizontalFlip(p=1.0 if flip else 0.0),
                        A.Affine(translate_percent=(t, t), scale=(s, s), rotate=(r, r), fit_output=False, p=1.0)])
        out = tf(image=img[0].numpy(), mask=mask[0].numpy())
        th = math.radians(r)
        c, sn = math.cos(th), math.sin(th)
        cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
        a00, a01, a10, a11 = c / s, sn / s, -sn / s, c / s
        a02 = cx - (a00 * (cx + t * w) + a01 * (cy + t * h))
        a12 = cy - (a10 * (cx + t * w) + a11 * (cy + t * h))
        q = torch.tensor([[int(flip)] + [int(round(v * FIX)) for v in (a00, a01, a02, a10, a11, a12)] + [0]])
        a, b = augment_pair(img.cuda(), mask.cuda(), q)
        ref_a = (torch.from_numpy(out["image"]).permute(2, 0, 1).float() / 255 - 0.5) / 0.5
        ref_b = torch.from_numpy(out["mask"]).permute(2, 0, 1).float() / 255
        # cv2.warpAffine interpolates with 1/32-pixel fixed-point weights and rounds to uint8: a few grey levels on edges
        assert (a[0].cpu() - ref_a).abs().mean().item() < 3.0 / 255, (s, r, t, flip)
        assert (b[0].cpu() != ref_b).float().mean().item() < 0.02, (s, r, t, flip)
