"""Shared by tests/test_baseline_shapes_gpu.py and tools/parity_report.py: run ONE G+D iteration of the CUDA path
(through TrainStep -> the C-ABI) and of oracle.train_step on the same weights / batch / label / alpha draws, and
return the measured differences.

Two oracle runs per case:
  * "fp32": the oracle as it stands (fp32 CPU restatement of train.py:99-168) -- losses, fake_B, D gradients and the
    flat G gradient are compared with it directly;
  * "matched": the oracle with bf16 rounding at the CUDA path's storage points (oracle.QUANT) and with the
    generator's ReLU masks / max-pool routes FORCED to the ones the CUDA forward took. A 30-layer ReLU network's
    gradient is piecewise linear in its activation pattern: a 0.4 % bf16 perturbation flips the masks of
    near-zero pre-activations, so the fp32 comparison measures conditioning, not kernel logic. With the pattern
    pinned both sides differentiate the SAME linear map and the tolerance on the flat gradient is a number.
"""
import time
from collections import OrderedDict

import torch
import torch.nn.functional as Fn

from gpu_util import cos, rel


class MaskFeed:
    """Stands in for oracle.ACT['relu']: the k-th call multiplies by the k-th recorded mask."""

    def __init__(self, masks):
        self.masks, self.i = masks, 0

    def __call__(self, t):
        m = self.masks[self.i]
        self.i += 1
        assert m.shape == t.shape, (self.i, m.shape, t.shape)
        return t * m.to(t.dtype)


class PoolFeed:
    """Stands in for oracle.POOL['max']: routes through the arg-max positions the CUDA forward saw."""

    def __init__(self, idxs):
        self.idxs, self.i = idxs, 0

    def __call__(self, t):
        idx = self.idxs[self.i]
        self.i += 1
        n, c, h, w = t.shape
        return t.flatten(2).gather(2, idx.flatten(2)).view(n, c, h // 2, w // 2)


def nchw(act_buf, c):
    """NHWC bf16 engine buffer -> fp32 NCHW on the host, real channels only."""
    return act_buf[..., :c].permute(0, 3, 1, 2).float().cpu()


def forward_pattern(engine):
    """ReLU masks (norm units, execution order == the oracle's call order) and max-pool arg-max indices of the
    CUDA generator forward that just ran."""
    from tactile_gan_b200.engine import ConvUnit
    masks, pools = [], []
    for u in engine.units:
        if not isinstance(u, ConvUnit) or not u.norm:
            continue
        y = nchw(u.y.buf, u.c_valid)
        masks.append(y > 0)
        if u.pool is not None and u.pool_mode == 2:
            pools.append(Fn.max_pool2d(y, 2, 2, return_indices=True)[1])
    return masks, pools


def flat(grads, ref):
    a, b = [], []
    for k, v in ref.items():
        if v is None or k.startswith("clstm"):
            continue
        a.append(grads[k].flatten().float().cpu())
        b.append(v.flatten().float())
    return torch.cat(a), torch.cat(b)


def build_nets(gen, nf, loss, seed, in_nc=3, out_nc=3):
    from tactile_gan_b200.discriminators.discriminators import create_disc
    from tactile_gan_b200.generators.generators import create_gen
    from tactile_gan_b200.util import init_weights
    torch.manual_seed(seed)
    act = loss == "ls"                                   # train.py:33
    netG = create_gen(gen, in_nc, out_nc, nf, act)
    netD = create_disc("patch", in_nc, out_nc, nf, True, act)
    init_weights(netG)                                   # the reference's N(0, .02) start (util.py:23-34)
    init_weights(netD)
    sd_g = OrderedDict((k, v.detach().clone()) for k, v in netG.state_dict().items())
    sd_d = OrderedDict((k, v.detach().clone()) for k, v in netD.state_dict().items())
    return netG, netD, sd_g, sd_d


def run_step_case(gen="UNet++", batch=2, size=256, nf=64, loss="ls", regularize=True, lambda_gp=0.01, lambda_per=1.0,
                  label_smoothing=True, seed=21, matched=True, binary_target=False):
    """-> dict of measured differences (see the keys at the bottom)."""
    import oracle as orc
    from tactile_gan_b200 import _C
    from tactile_gan_b200.step import TrainStep
    netG, netD, sd_g, sd_d = build_nets(gen, nf, loss, seed)
    g = torch.Generator().manual_seed(seed + 1)
    a, b = orc.synthetic_batch(g, batch, size)
    if binary_target:                                    # --target ch: three stacked grayscale masks (PairedDataset.py:73-76)
        b = (b > 0.7).float()
    alpha = torch.rand(batch, 1, generator=g)
    netG, netD = netG.cuda(), netD.cuda()
    ts = TrainStep(netG, netD, batch, size, size, loss=loss, lambda_gp=lambda_gp, lambda_per=lambda_per,
                   label_smoothing=label_smoothing)
    label = ts.ensure_label(generator=g)
    label = label.cpu() if label is not None else torch.ones(1)
    t0 = time.time()
    ts.step(a.cuda(), b.cuda(), regularize=regularize, alpha=alpha)
    got = ts.loss_dict()
    assert _C.error_flag() == 0
    fake = ts.fake_B.cpu()
    gd = ts.DA.store.grads_by_name()
    gg = ts.G.store.grads_by_name()
    masks, pools = forward_pattern(ts.G) if matched else (None, None)
    d_after = OrderedDict((k, v.detach().cpu().clone()) for k, v in netD.state_dict().items())
    t_cuda = time.time() - t0
    cfg = orc.StepConfig(gen=gen, loss=loss, lambda_gp=lambda_gp, lambda_per=lambda_per, regularize=regularize)
    clone = lambda sd: OrderedDict((k, v.clone()) for k, v in sd.items())
    t0 = time.time()
    ref = orc.train_step(clone(sd_g), clone(sd_d), {}, {}, a, b, label, alpha, cfg)
    t_ref = time.time() - t0
    out = {"case": f"{gen} B={batch} {size}^2 nf={nf} {loss} reg={int(regularize)} gp={lambda_gp} per={lambda_per} "
                   f"smooth={int(label_smoothing)}", "t_cuda_s": t_cuda, "t_oracle_s": t_ref}
    for k in ("loss_D", "gp", "G_GAN", "L1", "per"):
        out["loss:" + k] = (got[k], ref[k])
    out["fake_B"] = rel(fake, ref["fake_B"])
    fa, fb = flat(gd, ref["grads_D"])
    out["gradD_fp32"] = rel(fa, fb)
    fa, fb = flat(gg, ref["grads_G"])
    out["gradG_fp32"], out["gradG_fp32_cos"] = rel(fa, fb), cos(fa, fb)
    if matched:
        orc.QUANT["on"] = True
        orc.ACT["relu"] = MaskFeed(masks)
        orc.POOL["max"] = PoolFeed(pools)
        cfg.sd_d_after = d_after          # G step from the SAME updated discriminator (see oracle.StepConfig)
        try:
            refm = orc.train_step(clone(sd_g), clone(sd_d), {}, {}, a, b, label, alpha, cfg)
        finally:
            orc.QUANT["on"] = False
            orc.ACT["relu"] = Fn.relu
            orc.POOL["max"] = orc.max_pool_2x2
        out["fake_B_matched"] = rel(fake, refm["fake_B"])
        fa, fb = flat(gd, refm["grads_D"])
        out["gradD_matched"] = rel(fa, fb)
        fa, fb = flat(gg, refm["grads_G"])
        out["gradG_matched"], out["gradG_matched_cos"] = rel(fa, fb), cos(fa, fb)
        worst = ("", 0.0)
        for k, v in refm["grads_G"].items():
            if v is None or v.numel() < 4096 or k.startswith("clstm"):
                continue
            r = rel(gg[k], v)
            if r > worst[1]:
                worst = (k, r)
        out["gradG_matched_worst_tensor"] = worst
    del ts
    torch.cuda.empty_cache()
    return out


def check_losses(out, rel_tol=0.03, abs_tol=2e-3, overrides=None):
    for k in ("loss_D", "gp", "G_GAN", "L1", "per"):
        got, ref = out["loss:" + k]
        r, a_ = (overrides or {}).get(k, (rel_tol, abs_tol))
        assert abs(got - ref) <= r * abs(ref) + a_, (out["case"], k, got, ref)


def fmt(out):
    lines = [out["case"] + f"  (cuda {out['t_cuda_s']:.1f}s, oracle {out['t_oracle_s']:.1f}s)"]
    for k, v in out.items():
        if k.startswith("loss:"):
            lines.append(f"    {k[5:]:7s} cuda {v[0]:.6f}  oracle {v[1]:.6f}  rel {abs(v[0] - v[1]) / (abs(v[1]) + 1e-12):.2e}")
    for k, v in out.items():
        if k.startswith(("fake_B", "grad")):
            lines.append(f"    {k:28s} {v if isinstance(v, tuple) else round(v, 5)}")
    return "\n".join(lines)


def _sync_from_oracle(net, sd, opt):
    """Weights and Adam state of the oracle -> the CUDA module's parameters / ParamStore arenas."""
    net.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    st = net._tg_store
    step = 0
    for i, name in enumerate(st.names):
        s = opt.get(name)
        if s is None:
            st.m_views[i].zero_()
            st.v_views[i].zero_()
            continue
        st.m_views[i].copy_(s["exp_avg"])
        st.v_views[i].copy_(s["exp_avg_sq"])
        step = max(step, int(s["step"]))
    st.step_count = step


def run_trajectory(gen="UNet++", nf=16, size=64, batch=2, steps=8, resync=True, seed=5, loss="ls", reg_every=1):
    """`steps` consecutive G+D iterations (train.py:99-168 incl. both Adam updates) on both sides.
    resync=True: before every step the CUDA side is loaded with the oracle's weights and Adam moments, so each
    step is compared from an identical state (per-step tolerance); resync=False: both run free from the same
    start (drift bound). -> list of per-step dicts."""
    import oracle as orc
    from tactile_gan_b200 import _C
    from tactile_gan_b200.step import TrainStep
    netG, netD, sd_g, sd_d = build_nets(gen, nf, loss, seed)
    netG, netD = netG.cuda(), netD.cuda()
    ts = TrainStep(netG, netD, batch, size, size, loss=loss)
    g = torch.Generator().manual_seed(seed + 1)
    label = ts.ensure_label(generator=g)
    label = label.cpu() if label is not None else torch.ones(1)
    opt_g, opt_d = {}, {}
    cfg = orc.StepConfig(gen=gen, loss=loss)
    rows = []
    for k in range(steps):
        a, b = orc.synthetic_batch(g, batch, size)
        alpha = torch.rand(batch, 1, generator=g)
        reg = (k % reg_every) == 0
        if resync and k > 0:
            _sync_from_oracle(netG, sd_g, opt_g)
            _sync_from_oracle(netD, sd_d, opt_d)
        ts.step(a.cuda(), b.cuda(), regularize=reg, alpha=alpha)
        got = ts.loss_dict()
        fake = ts.fake_B.cpu()
        gd = ts.DA.store.grads_by_name()
        cfg.regularize = reg
        ref = orc.train_step(sd_g, sd_d, opt_g, opt_d, a, b, label, alpha, cfg)
        row = {"step": k + 1, "fake_B": rel(fake, ref["fake_B"])}
        for name in ("loss_D", "gp", "G_GAN", "L1", "per"):
            row[name] = (got[name], ref[name])
        fa, fb = flat(gd, ref["grads_D"])
        row["gradD"] = rel(fa, fb)
        wd = torch.cat([(netD.state_dict()[n].cpu() - v).flatten() for n, v in sd_d.items()])
        row["weightsD_maxabs"] = float(wd.abs().max())
        row["weightsD_meanabs"] = float(wd.abs().mean())
        rows.append(row)
    assert _C.error_flag() == 0
    del ts
    torch.cuda.empty_cache()
    return rows


def fmt_traj(rows):
    out = []
    for r in rows:
        ls = "  ".join(f"{k} {r[k][0]:.5f}/{r[k][1]:.5f}" for k in ("loss_D", "gp", "G_GAN", "L1", "per"))
        out.append(f"    step {r['step']}: {ls}  fake_B {r['fake_B']:.4f} gradD {r['gradD']:.4f} "
                   f"dW_D max {r['weightsD_maxabs']:.2e} mean {r['weightsD_meanabs']:.2e}")
    return "\n".join(out)


def run_forward_case(gen, batch, size, nf=64, seed=3, samples=None, binary=False):
    """Generator inference forward (test.py:202-203) vs the fp32 oracle; `samples`: subset of the batch the oracle
    evaluates (InstanceNorm is per sample, so any sample of a large batch can be checked on its own)."""
    import oracle as orc
    from tactile_gan_b200.generators.generators import create_gen
    from tactile_gan_b200.util import init_weights
    torch.manual_seed(seed)
    net = create_gen(gen, 3, 3, nf, True)
    init_weights(net)
    sd = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(x.cuda()).cpu()
    idx = list(range(batch)) if samples is None else list(samples)
    with torch.no_grad():
        ref = orc.gen_forward(gen, sd, x[idx], True)
    del net
    torch.cuda.empty_cache()
    return rel(y[idx], ref), float((y[idx] - ref).abs().max())
