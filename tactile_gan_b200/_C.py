"""ctypes binding of libtactile_gan_b200.so (the C-ABI declared in include/tactile_gan_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is
raised -- the product path never routes through PyTorch ops or the CPU oracle.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_float, c_int, c_int8, c_longlong,
                    c_void_p)

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TG_LIB_PATH") or os.path.join(_HERE, "libtactile_gan_b200.so")   # TG_LIB_PATH: A/B builds

TG_MAX_SRC = 6
TG_MAX_TAPS = 16
ACT_NONE, ACT_LRELU, ACT_SIGMOID, ACT_RELU, ACT_TANH = 0, 1, 2, 3, 4
GAN_MODES = {"ls": 0, "ce": 1, "w": 2, "hinge": 3}


class TgError(RuntimeError):
    pass


class View(Structure):
    _fields_ = [("ptr", c_void_p), ("n", c_int), ("h", c_int), ("w", c_int), ("c", c_int),
                ("sn", c_longlong), ("sh", c_longlong), ("sw", c_longlong)]


class ConvSrc(Structure):
    _fields_ = [("act", View), ("wgt", c_void_p), ("wgt_taps", c_int), ("wgt_rows", c_int),
                ("wgt_k", c_int), ("k_off", c_int), ("row_off", c_int)]


class ConvDesc(Structure):
    _fields_ = [("num_src", c_int), ("src", ConvSrc * TG_MAX_SRC), ("out", View), ("taps", c_int),
                ("stride", c_int), ("tap_dy", c_int8 * TG_MAX_TAPS), ("tap_dx", c_int8 * TG_MAX_TAPS),
                ("tap_w", c_int8 * TG_MAX_TAPS), ("bias", c_void_p), ("bias_len", c_int),
                ("stats_partial", c_void_p), ("stats_tiles_total", c_int), ("stats_tile_off", c_int),
                ("act", c_int), ("slope", c_float), ("pool_out", c_int), ("splitk_ws", c_void_p),
                ("splitk_ws_bytes", c_longlong)]


class WgradDesc(Structure):
    _fields_ = [("num_src", c_int), ("p", View * TG_MAX_SRC), ("q", View), ("taps", c_int),
                ("stride", c_int), ("tap_dy", c_int8 * TG_MAX_TAPS), ("tap_dx", c_int8 * TG_MAX_TAPS),
                ("tap_w", c_int8 * TG_MAX_TAPS), ("dw", c_void_p), ("dw_rows", c_int), ("dw_cols", c_int)]


_lib = None


def lib():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TgError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no PyTorch/CPU fallback for the tactile-gan hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.tg_last_error.restype = c_char_p
        _lib.tg_error_flag_device_ptr.restype = c_void_p
        _lib.tg_plan_destroy.restype = None
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise TgError(f"{what}: {lib().tg_last_error().decode()}")


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def view(t, c=None):
    """View of a (N,H,W,C) bf16 tensor (possibly strided in N/H/W; C contiguous)."""
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.stride(3) == 1, (t.dtype, t.shape, t.stride())
    n, h, w, cc = t.shape
    return View(t.data_ptr(), n, h, w, cc if c is None else c, t.stride(0), t.stride(1), t.stride(2))


def _taps(arr, vals):
    for i, v in enumerate(vals):
        arr[i] = int(v)


# launch accounting (bench.py): number of kernels launched, and optional per-plan CUDA-event timing
COUNTERS = {"launches": 0}
TIMING = {"on": False, "records": [], "tail": False, "tail_records": []}
# records: (kind, flops, start_event, end_event, tag); tail_records: (name, start_event, end_event)
_KERNELS_PER_CALL = {"in_bwd2": 2}


class Plan:
    """Owns a tg_plan*; keeps the tensors it points at alive."""

    def __init__(self, handle, keep, kind="conv", flops=0.0, tag=""):
        self.handle = handle
        self.keep = keep
        self.kind = kind
        self.flops = flops
        self.tag = tag

    def run(self):
        COUNTERS["launches"] += 1
        if TIMING["on"]:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib().tg_plan_run(self.handle, stream_ptr()), "tg_plan_run")
            e1.record()
            TIMING["records"].append((self.kind, self.flops, e0, e1, self.tag))
            return
        check(lib().tg_plan_run(self.handle, stream_ptr()), "tg_plan_run")

    def __del__(self):
        try:
            if self.handle:
                lib().tg_plan_destroy(self.handle)
        except Exception:
            pass


def conv_query_tiles(n, ho, wo, want_stats):
    out = (c_int * 4)()
    check(lib().tg_conv_query_tiles(n, ho, wo, int(want_stats), out))
    return tuple(out)  # th, tw, tn, tiles_per_img


SPLITK_WS_BYTES = 256 << 20
SPLITK_SLOT = 0          # plans created while this is k > 0 use the k-th extra workspace (launch chains on side streams)
_splitk_ws = {}


def splitk_workspace(device, slot=0):
    """One fp32 workspace per device for the split-K convolutions (tg_conv_desc.splitk_ws): the conv launches of a
    step are ordered on one stream, so they can share it. Engines that run chains of units on side streams
    (UNetPPEngine's small inference batches) give every stream its own, smaller one (`slot` > 0)."""
    dev = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    key = (dev, slot)
    if key not in _splitk_ws:
        nbytes = SPLITK_WS_BYTES if slot == 0 else SPLITK_WS_BYTES // 8
        _splitk_ws[key] = torch.empty(nbytes // 4, dtype=torch.float32, device=torch.device("cuda", dev))
    return _splitk_ws[key]


def conv_plan(srcs, out, taps, stride=1, bias=None, stats_partial=None, act=ACT_NONE, slope=0.2,
              stats_tiles_total=0, stats_tile_off=0, cout_real=None, flops=None, pool_out=False):
    """srcs: list of dict(act=tensor NHWC, wgt=packed bf16 [taps][rows][k], k_off=0, row_off=0)
    taps: list of (dy, dx, w_index)."""
    d = ConvDesc()
    d.num_src = len(srcs)
    keep = [out, bias, stats_partial]
    for i, s in enumerate(srcs):
        w = s["wgt"]
        assert w.dtype == torch.bfloat16 and w.dim() == 3 and w.is_contiguous()
        d.src[i].act = view(s["act"])
        d.src[i].wgt = w.data_ptr()
        d.src[i].wgt_taps, d.src[i].wgt_rows, d.src[i].wgt_k = w.shape
        d.src[i].k_off = s.get("k_off", 0)
        d.src[i].row_off = s.get("row_off", 0)
        keep += [s["act"], w]
    d.out = view(out)
    d.taps = len(taps)
    d.stride = stride
    _taps(d.tap_dy, [t[0] for t in taps])
    _taps(d.tap_dx, [t[1] for t in taps])
    _taps(d.tap_w, [t[2] for t in taps])
    d.bias = bias.data_ptr() if bias is not None else None
    d.bias_len = bias.numel() if bias is not None else 0
    d.stats_partial = stats_partial.data_ptr() if stats_partial is not None else None
    if stats_partial is not None and not stats_tiles_total:
        stats_tiles_total = stats_partial.shape[1]   # [n][tile slots][c][2]
    d.stats_tiles_total = stats_tiles_total
    d.stats_tile_off = stats_tile_off
    d.act = act
    d.slope = slope
    d.pool_out = int(pool_out)
    if os.environ.get("TG_SPLITK", "1") != "0":
        ws = splitk_workspace(out.device, SPLITK_SLOT)
        d.splitk_ws, d.splitk_ws_bytes = ws.data_ptr(), ws.numel() * 4
        keep.append(ws)
    h = c_void_p()
    check(lib().tg_conv_plan_create(byref(d), byref(h)), "tg_conv_plan_create")
    # algorithmic FLOPs of this launch: real (unpadded) channels, every output pixel, every tap
    n, ho, wo, _ = out.shape
    if pool_out:
        ho, wo = 2 * ho, 2 * wo
    if flops is None:
        flops = 0.0
        for s in srcs:
            flops += 2.0 * n * ho * wo * len(taps) * s.get("c_real", s["act"].shape[3]) * (cout_real or out.shape[3])
    tag = "conv n%d %dx%d cin[%s] cout%d taps%d s%d%s" % (
        n, ho, wo, ",".join(str(s["act"].shape[3]) for s in srcs), out.shape[3], len(taps), stride,
        (" stats" if stats_partial is not None else "") + (" pool" if pool_out else ""))
    return Plan(h, keep, "conv", flops, tag)


def wgrad_plan(p_srcs, q, taps, dw, stride=1, p_real=None, q_real=None, dw_row_off=0, flops=None):
    """dw: fp32 [taps_total][rows >= q.C][cols == sum p.C]."""
    d = WgradDesc()
    d.num_src = len(p_srcs)
    for i, t in enumerate(p_srcs):
        d.p[i] = view(t)
    d.q = view(q)
    d.taps = len(taps)
    d.stride = stride
    _taps(d.tap_dy, [t[0] for t in taps])
    _taps(d.tap_dx, [t[1] for t in taps])
    _taps(d.tap_w, [t[2] for t in taps])
    assert dw.dtype == torch.float32 and dw.dim() == 3 and dw.is_contiguous()
    d.dw = dw.data_ptr() + dw_row_off * dw.shape[2] * 4   # row slice of every tap plane
    d.dw_rows, d.dw_cols = dw.shape[1], dw.shape[2]
    h = c_void_p()
    check(lib().tg_wgrad_plan_create(byref(d), byref(h)), "tg_wgrad_plan_create")
    n, ho, wo, qc = q.shape
    pc = p_real if p_real is not None else sum(t.shape[3] for t in p_srcs)
    if flops is None:
        flops = 2.0 * n * ho * wo * len(taps) * pc * (q_real if q_real is not None else qc)
    tag = "wgrad n%d %dx%d p[%s] q%d taps%d s%d" % (
        n, ho, wo, ",".join(str(t.shape[3]) for t in p_srcs), qc, len(taps), stride)
    return Plan(h, list(p_srcs) + [q, dw], "wgrad", flops, tag)


def error_flag():
    """Value of the device-side pipeline-timeout flag (synchronises)."""
    return int(lib().tg_error_flag_read())


def call(name, *args, nbytes=None):
    """Call a tail kernel launcher `tg_<name>(..., stream)`. `nbytes`: algorithmic HBM bytes of the launch
    (what it must read + write once), recorded with its CUDA-event timing while bench.py measures."""
    fn = getattr(lib(), "tg_" + name)
    COUNTERS["launches"] += _KERNELS_PER_CALL.get(name, 1)
    if TIMING["on"] and nbytes is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args, stream_ptr()), "tg_" + name)
        e1.record()
        TIMING["records"].append(("tail:" + name, float(nbytes), e0, e1, name))
        return
    if TIMING["tail"]:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args, stream_ptr()), "tg_" + name)
        e1.record()
        shape = ",".join(str(a) for a in args if isinstance(a, int))
        TIMING["tail_records"].append((name + "(" + shape + ")", e0, e1))
        return
    check(fn(*args, stream_ptr()), "tg_" + name)


F = c_float
I = c_int
LL = c_longlong
