"""-m gpu: properties of the path at BASELINE.json's full size (UNet++ nf=64 + PatchDiscriminator, batch 32,
256x256, configs[1]) -- the oracle takes minutes there, so parity is checked through size-independent facts:

* the forward pass is deterministic (no atomics on that path): two runs are bit-identical;
* samples are independent (InstanceNorm is per sample): the batch-32 output equals two batch-16 runs bit for bit,
  and the batch-32 gradient equals the sum of the two half-batch gradients (fp32 atomics / bf16 storage: 2 %);
* InstanceNorm does what it says at 256^2: every (n, c) plane of a unit's normalised output has mean beta_c and
  variance gamma_c^2 before the ReLU (checked through the stored raw / mean / rstd);
* the logged L1 and LSGAN terms equal torch reductions of the tensors the step produced;
* one whole step at full size leaves finite weights and moves every parameter tensor.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
B, S, NF = 32, 256, 64


def _nets(seed=11):
    from tactile_gan_b200.discriminators.discriminators import create_disc
    from tactile_gan_b200.generators.generators import create_gen
    from tactile_gan_b200.util import init_weights
    torch.manual_seed(seed)
    g, d = create_gen("UNet++", 3, 3, NF, True), create_disc("patch", 3, 3, NF, True, True)
    init_weights(g)
    init_weights(d)
    return g.cuda(), d.cuda()


def _batch(seed=3):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda(), torch.rand(B, 3, S, S, generator=g).cuda()


def test_full_size_forward_is_deterministic_and_sample_independent():
    from tactile_gan_b200.engine import build_generator_engine
    netG, _ = _nets()
    a, _ = _batch()
    full = build_generator_engine("unet++", netG, B, S, S, False)
    y1 = full.forward(a).clone()
    y2 = full.forward(a).clone()
    assert torch.equal(y1, y2)
    half = build_generator_engine("unet++", netG, B // 2, S, S, False)
    lo = half.forward(a[:B // 2].contiguous()).clone()
    hi = half.forward(a[B // 2:].contiguous()).clone()
    assert torch.equal(torch.cat([lo, hi]), y1)
    assert torch.isfinite(y1).all() and y1.abs().max() <= 1.0          # tanh head
    # InstanceNorm at full size, from the stored conv output and statistics of the first block's second unit
    u = full.X[0, 0][1]
    raw = u.raw[..., :u.c_valid].float()
    mean, rstd = u.mr[:, :u.c_valid, 0], u.mr[:, :u.c_valid, 1]
    xhat = (raw - mean[:, None, None, :]) * rstd[:, None, None, :]
    assert xhat.mean((1, 2)).abs().max() < 2e-3
    assert (xhat.var((1, 2), unbiased=False) - 1).abs().max() < 2e-3
    del full, half
    torch.cuda.empty_cache()


def test_full_size_step_gradients_add_over_half_batches_and_losses_match_torch():
    from tactile_gan_b200.step import TrainStep
    a, b = _batch()
    g = torch.Generator().manual_seed(5)
    alpha = torch.rand(B, 1, generator=g)

    def run(sl, n):
        netG, netD = _nets()
        ts = TrainStep(netG, netD, n, S, S, lr=0.0, label_smoothing=False)
        ts.step(a[sl].contiguous(), b[sl].contiguous(), regularize=True, alpha=alpha[sl])
        torch.cuda.synchronize()
        out = dict(gG=ts.G.store.grad_arena.clone(), gD=ts.DA.store.grad_arena.clone(), loss=ts.loss_dict(),
                   fake=ts.fake_B.clone(), pred=ts.S1.pred[..., 0].float().clone())
        del ts, netG, netD
        torch.cuda.empty_cache()
        return out

    full = run(slice(0, B), B)
    lo, hi = run(slice(0, B // 2), B // 2), run(slice(B // 2, B), B // 2)
    # mean losses: full = average of the halves; gradients of mean losses likewise
    for k in ("loss_D", "gp", "G_GAN", "L1", "per"):
        assert full["loss"][k] == pytest.approx((lo["loss"][k] + hi["loss"][k]) / 2, rel=5e-3, abs=1e-6), k
    for key, tol in (("gD", 2e-2), ("gG", 5e-2)):
        want = (lo[key] + hi[key]) / 2
        err = float((full[key] - want).norm() / want.norm())
        assert err < tol, (key, err)
    # logged terms against torch reductions of what the step produced (train.py:142-146; label smoothing off)
    assert full["loss"]["L1"] == pytest.approx(float((full["fake"] - b).abs().mean()), rel=1e-4)
    assert full["loss"]["G_GAN"] == pytest.approx(float(((full["pred"] - 1) ** 2).mean()), rel=2e-2)


def test_full_size_step_updates_every_parameter():
    from tactile_gan_b200.step import TrainStep
    netG, netD = _nets()
    a, b = _batch()
    before = [p.detach().clone() for p in list(netG.parameters()) + list(netD.parameters())]
    ts = TrainStep(netG, netD, B, S, S)
    for _ in range(2):
        ts.step(a, b)
    torch.cuda.synchronize()
    losses = ts.loss_dict()
    assert all(v == v and abs(v) < 1e3 for v in losses.values())
    for p0, p in zip(before, list(netG.parameters()) + list(netD.parameters())):
        assert torch.isfinite(p).all()
        assert not torch.equal(p0, p.detach())
