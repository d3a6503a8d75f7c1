"""GPU bring-up checks for the tcgen05 kernels: each case compares against torch fp32 math on the
same bf16-rounded operands and prints error statistics. Run one case group per process:
    python tools/bringup.py conv | wgrad | tail
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tactile_gan_b200 import _C  # noqa: E402

dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def pad64(c):
    return (c + 63) // 64 * 64


def nhwc_pad(x, cpad=None):
    """fp32 NCHW -> bf16 NHWC with channels zero-padded to a multiple of 64."""
    n, c, h, w = x.shape
    cp = cpad or pad64(c)
    out = torch.zeros(n, h, w, cp, dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def pack_w(w, opad=None, ipad=None):
    """torch conv weight [O][I][kh][kw] -> bf16 [taps][opad][ipad]."""
    o, i, kh, kw = w.shape
    op, ip = opad or pad64(o), ipad or pad64(i)
    out = torch.zeros(kh * kw, op, ip, dtype=torch.bfloat16, device=w.device)
    out[:, :o, :i] = w.permute(2, 3, 0, 1).reshape(kh * kw, o, i).to(torch.bfloat16)
    return out


def conv_taps(kh, kw, pad):
    return [(r - pad, s - pad, r * kw + s) for r in range(kh) for s in range(kw)]


def report(name, got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    ok = rel < 2e-2
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max_abs_err={err:.4e} ref_max={scale:.4e} rel_l2={rel:.3e}", flush=True)
    return ok


def case_conv(name, n, cins, cout, h, w, k, stride, pad, bias=False, act=_C.ACT_NONE, stats=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs = [torch.randn(n, c, h, w, generator=g).to(dev) for c in cins]
    wt = (torch.randn(cout, sum(cins), k, k, generator=g) * 0.05).to(dev)
    b = torch.randn(cout, generator=g).to(dev) if bias else None
    xs_b = [x.to(torch.bfloat16).float() for x in xs]
    wt_b = wt.to(torch.bfloat16).float()
    ref = F.conv2d(torch.cat(xs_b, 1), wt_b, b, stride=stride, padding=pad)
    ho, wo = ref.shape[2], ref.shape[3]
    if act == _C.ACT_LRELU:
        ref_a = F.leaky_relu(ref, 0.2)
    elif act == _C.ACT_SIGMOID:
        ref_a = torch.sigmoid(ref)
    else:
        ref_a = ref
    cop = pad64(cout)
    out = torch.full((n, ho, wo, cop), 7.0, dtype=torch.bfloat16, device=dev)
    srcs = []
    koff = 0
    # per-source packed weights share one [taps][cop][sum ipad] tensor (virtual concat)
    ipads = [pad64(c) for c in cins]
    wp = torch.zeros(k * k, cop, sum(ipads), dtype=torch.bfloat16, device=dev)
    ci = 0
    for c, ip in zip(cins, ipads):
        wp[:, :cout, koff:koff + c] = wt[:, ci:ci + c].permute(2, 3, 0, 1).reshape(k * k, cout, c).to(torch.bfloat16)
        koff += ip
        ci += c
    koff = 0
    for x, ip in zip(xs, ipads):
        srcs.append(dict(act=nhwc_pad(x), wgt=wp, k_off=koff))
        koff += ip
    bias_p = None
    if bias:
        bias_p = torch.zeros(cop, device=dev)
        bias_p[:cout] = b
    sp = None
    if stats:
        th, tw, tn, tpi = _C.conv_query_tiles(n, ho, wo, True)
        sp = torch.zeros(n, tpi, cop, 2, device=dev)
    plan = _C.conv_plan(srcs, out, conv_taps(k, k, pad), stride=stride, bias=bias_p, stats_partial=sp, act=act)
    plan.run()
    torch.cuda.synchronize()
    flag = _C.error_flag()
    if flag:
        print(f"[BAD] {name}: device error flag {flag}")
        return False
    got = out[..., :cout].permute(0, 3, 1, 2)
    ok = report(name, got, ref_a)
    if cop > cout:
        padmax = out[..., cout:].float().abs().max().item()
        if bias is False and act == _C.ACT_NONE and padmax != 0:
            print(f"      pad channels not zero: {padmax}")
    if stats:
        s = sp.sum(1)  # [n][c][2]
        gb = out[..., :cout].float()
        ref_s = gb.sum((1, 2))
        ref_q = (gb * gb).sum((1, 2))
        ok &= report(name + " stats.sum", s[:, :cout, 0], ref_s)
        ok &= report(name + " stats.sumsq", s[:, :cout, 1], ref_q)
    # timing
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        plan.run()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    flops = 2.0 * n * ho * wo * cout * sum(cins) * k * k
    print(f"      {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s (useful)", flush=True)
    return ok


_idx = -1


def _sel(only, fn, *a, **k):
    global _idx
    _idx += 1
    if only is not None and _idx != only:
        return True
    try:
        return fn(*a, **k)
    except Exception as e:  # CUDA faults poison the context: report and stop this process
        print(f"[BAD] case {_idx} {a[0]}: exception {type(e).__name__}: {e}", flush=True)
        sys.exit(3)


def group_conv(only=None):
    ok = True
    global _idx
    _idx = -1
    ok &= _sel(only, case_conv, "1x1 64->64 16x16", 2, [64], 64, 16, 16, 1, 1, 0)
    ok &= _sel(only, case_conv, "3x3 64->64 32x32 p1", 2, [64], 64, 32, 32, 3, 1, 1)
    ok &= _sel(only, case_conv, "3x3 128->128 32x32 p1 stats", 2, [128], 128, 32, 32, 3, 1, 1, stats=True)
    ok &= _sel(only, case_conv, "3x3 256->256 16x16 p1", 3, [256], 256, 16, 16, 3, 1, 1)
    ok &= _sel(only, case_conv, "3x3 512->512 16x16 p1", 2, [512], 512, 16, 16, 3, 1, 1)
    ok &= _sel(only, case_conv, "3x3 concat 64+64+128->64 64x64", 2, [64, 64, 128], 64, 64, 64, 3, 1, 1, stats=True)
    ok &= _sel(only, case_conv, "3x3 3->64 64x64 p1 (pad cin)", 2, [3], 64, 64, 64, 3, 1, 1)
    ok &= _sel(only, case_conv, "3x3 p0 s1 odd 61->59 128->256", 2, [128], 256, 61, 61, 3, 1, 0, stats=True)
    ok &= _sel(only, case_conv, "3x3 p0 s2 6->64 64x64 bias lrelu", 2, [6], 64, 64, 64, 3, 2, 0, bias=True, act=_C.ACT_LRELU)
    ok &= _sel(only, case_conv, "3x3 p0 s2 64->128 127->63", 2, [64], 128, 127, 127, 3, 2, 0, stats=True)
    ok &= _sel(only, case_conv, "3x3 512->1 59->57 bias sigmoid", 2, [512], 1, 59, 59, 3, 1, 0, bias=True, act=_C.ACT_SIGMOID)
    ok &= _sel(only, case_conv, "4x4 s2 p1 64->128 64x64", 2, [64], 128, 64, 64, 4, 2, 1)
    ok &= _sel(only, case_conv, "big 3x3 64->64 256x256 n8", 8, [64], 64, 256, 256, 3, 1, 1, stats=True)
    ok &= _sel(only, case_conv, "big 3x3 384->64 256x256 n4", 4, [64, 64, 64, 64, 128], 64, 256, 256, 3, 1, 1)
    ok &= _sel(only, case_conv, "big 3x3 512->512 32x32 n32", 32, [512], 512, 32, 32, 3, 1, 1)
    ok &= _sel(only, case_conv, "big 3x3 256->512 59x59 n32 p0", 32, [256], 512, 61, 61, 3, 1, 0)
    return ok


def case_wgrad(name, n, cins, cout, h, w, k, stride, pad, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs = [torch.randn(n, c, h, w, generator=g).to(dev) for c in cins]
    xcat = torch.cat([x.to(torch.bfloat16).float() for x in xs], 1).requires_grad_(False)
    wt = torch.zeros(cout, sum(cins), k, k, device=dev, requires_grad=True)
    y = F.conv2d(xcat, wt, None, stride=stride, padding=pad)
    dy = (torch.randn(y.shape, generator=g) * 0.1).to(dev)
    dy_b = dy.to(torch.bfloat16).float()
    (ref,) = torch.autograd.grad(y, wt, dy_b)
    ipads = [pad64(c) for c in cins]
    cop = pad64(cout)
    dw = torch.zeros(k * k, cop, sum(ipads), device=dev)
    plan = _C.wgrad_plan([nhwc_pad(x) for x in xs], nhwc_pad(dy), conv_taps(k, k, pad), dw, stride=stride)
    plan.run()
    torch.cuda.synchronize()
    flag = _C.error_flag()
    if flag:
        print(f"[BAD] {name}: device error flag {flag}")
        return False
    # unpack dw [tap][co][ci_padded concat] -> [O][I][kh][kw]
    parts = []
    koff = 0
    for c, ip in zip(cins, ipads):
        parts.append(dw[:, :cout, koff:koff + c])
        koff += ip
    got = torch.cat(parts, 2).reshape(k, k, cout, sum(cins)).permute(2, 3, 0, 1)
    ok = report(name, got, ref)
    dw.zero_()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        plan.run()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    flops = 2.0 * n * y.shape[2] * y.shape[3] * cout * sum(cins) * k * k
    print(f"      {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s (useful)", flush=True)
    return ok


def group_wgrad(only=None):
    ok = True
    global _idx
    _idx = -1
    ok &= _sel(only, case_wgrad, "wgrad 1x1 64->64 16x16", 2, [64], 64, 16, 16, 1, 1, 0)
    ok &= _sel(only, case_wgrad, "wgrad 3x3 64->64 32x32", 2, [64], 64, 32, 32, 3, 1, 1)
    ok &= _sel(only, case_wgrad, "wgrad 3x3 128->256 32x32", 2, [128], 256, 32, 32, 3, 1, 1)
    ok &= _sel(only, case_wgrad, "wgrad 3x3 concat 64+64+128->64 64x64", 2, [64, 64, 128], 64, 64, 64, 3, 1, 1)
    ok &= _sel(only, case_wgrad, "wgrad 3x3 3->64 64x64", 2, [3], 64, 64, 64, 3, 1, 1)
    ok &= _sel(only, case_wgrad, "wgrad 3x3 p0 odd 61 128->256", 2, [128], 256, 61, 61, 3, 1, 0)
    ok &= _sel(only, case_wgrad, "wgrad 3x3 p0 s2 64->128 127", 2, [64], 128, 127, 127, 3, 2, 0)
    ok &= _sel(only, case_wgrad, "wgrad 4x4 s2 p1 64->128 64", 2, [64], 128, 64, 64, 4, 2, 1)
    ok &= _sel(only, case_wgrad, "wgrad big 3x3 64->64 256x256 n8", 8, [64], 64, 256, 256, 3, 1, 1)
    ok &= _sel(only, case_wgrad, "wgrad big 3x3 512->512 32x32 n32", 32, [512], 512, 32, 32, 3, 1, 1)
    ok &= _sel(only, case_wgrad, "wgrad big 3x3 384->64 256x256 n4", 4, [64, 64, 64, 64, 128], 64, 256, 256, 3, 1, 1)
    return ok


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "conv"
    only = int(sys.argv[2]) if len(sys.argv) > 2 else None
    t = time.time()
    ok = {"conv": group_conv, "wgrad": group_wgrad}[which](only)
    print(f"== {which}: {'ALL OK' if ok else 'FAILURES'} ({time.time()-t:.1f}s)")
    sys.exit(0 if ok else 1)
