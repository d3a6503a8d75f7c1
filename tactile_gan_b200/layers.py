"""Device-side parameter plumbing shared by the generator / discriminator engines.

* ConvLayer   -- one nn.Conv2d / nn.ConvTranspose2d weight: fp32 master (the nn.Parameter, torch
                 layout), bf16 packs in the two layouts the implicit-GEMM kernels read, an fp32
                 gradient in the packed layout, and factories for the forward / dgrad / wgrad plans.
* ParamStore  -- all parameters of one network: flat gradient arena (one NCCL allreduce), Adam
                 moments in torch layout (so optimizer.state_dict() matches torch.optim.Adam), and the
                 device table the fused Adam + re-pack kernel walks (tg_adam_step).
"""
import ctypes
from ctypes import Structure, c_int, c_longlong, c_void_p

import torch

from . import _C
from ._C import ACT_NONE


def pad64(c):
    return (c + 63) // 64 * 64


class AdamRow(Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("m", c_void_p), ("v", c_void_p),
                ("pack_fwd", c_void_p), ("pack_bwd", c_void_p), ("numel", c_longlong),
                ("kind", c_int), ("kh", c_int), ("kw", c_int), ("dim1", c_int), ("o_pad", c_int),
                ("i_pad", c_int), ("nseg", c_int), ("seg_end", c_int * 6), ("seg_shift", c_int * 6),
                ("pad_", c_int)]


assert ctypes.sizeof(AdamRow) == 136


def conv_taps(kh, kw, pad):
    """Forward taps of a stride-s Conv2d: input = out*s + (r - pad); weight tap index r*kw+s."""
    return [(r - pad, s - pad, r * kw + s) for r in range(kh) for s in range(kw)]


def dgrad_taps_s1(kh, kw, pad):
    """dX[y] = sum_r' dY[y + r' - (k-1-pad)] * Wflip[r'] (pack_bwd stores taps mirrored)."""
    return [(r - (kh - 1 - pad), s - (kw - 1 - pad), r * kw + s) for r in range(kh) for s in range(kw)]


def phase_taps(kh, kw, pad, stride, py, px, flipped):
    """Taps of output phase (py,px) of a transposed (fractionally strided) convolution:
    out[stride*a + py] += in[a + (py + pad - r)/stride] * W[r] for r with (py+pad-r) % stride == 0.
    `flipped`: weight tap axis stored mirrored (pack_bwd of a Conv2d) or not (pack_fwd of a ConvT)."""
    taps = []
    for r in range(kh):
        if (py + pad - r) % stride:
            continue
        for s in range(kw):
            if (px + pad - s) % stride:
                continue
            widx = ((kh - 1 - r) * kw + (kw - 1 - s)) if flipped else (r * kw + s)
            taps.append(((py + pad - r) // stride, (px + pad - s) // stride, widx))
    return taps


class ConvLayer:
    """kind 'conv': weight [O][I][kh][kw]; kind 'convT': weight [I][O][kh][kw] (torch layouts).
    `in_split`: channel counts of the concatenated inputs (each padded to 64 separately)."""

    def __init__(self, name, weight, bias, kind, stride, pad, in_split, device):
        self.name, self.weight, self.bias, self.kind = name, weight, bias, kind
        self.stride, self.pad = stride, pad
        if kind in ("conv", "head", "cols"):
            self.O, self.I, self.kh, self.kw = weight.shape
        else:
            self.I, self.O, self.kh, self.kw = weight.shape
        assert sum(in_split) == self.I, (name, in_split, self.I)
        self.in_split = list(in_split)
        self.seg_pad = [pad64(c) for c in in_split]
        self.k_off = [sum(self.seg_pad[:i]) for i in range(len(in_split))]
        self.i_pad = sum(self.seg_pad)
        self.o_pad = pad64(self.O)
        self.taps = self.kh * self.kw
        bf = dict(dtype=torch.bfloat16, device=device)
        if kind == "head":
            # single-output conv as a 1x1 GEMM whose output channels are the taps (tg_head_gather sums them)
            assert self.O == 1 and self.taps <= 16 and pad == 0 and stride == 1
            self.pack_shape, self.pack_bwd_shape = (1, 64, self.i_pad), (1, self.i_pad, 64)
        elif kind == "cols":
            # thin-input conv on im2col rows written by tg_im2col_pack: a 1x1 GEMM with K = taps * Cin <= 64
            assert self.taps * self.I <= 64 and pad == 0 and len(in_split) == 1
            self.pack_shape, self.pack_bwd_shape = (1, self.o_pad, 64), (1, 64, self.o_pad)
        elif kind == "conv":
            self.pack_shape = (self.taps, self.o_pad, self.i_pad)
            self.pack_bwd_shape = (self.taps, self.i_pad, self.o_pad)
        else:
            self.pack_shape = (self.taps, self.o_pad, self.i_pad)
            self.pack_bwd_shape = (self.taps, self.i_pad, self.o_pad)
        self.pack_fwd = torch.zeros(*self.pack_shape, **bf)
        self.pack_bwd = torch.zeros(*self.pack_bwd_shape, **bf)
        # fp32 gradient, in the layout the weight-gradient kernel emits: pack_shape for conv / head / cols,
        # [taps][i_pad][o_pad] for convT
        self.grad_shape = self.pack_bwd_shape if kind == "convT" else self.pack_shape
        self.grad = None
        self.bias_grad = None
        self._scratch = {}

    # ---- plan factories -------------------------------------------------------------------
    def out_hw(self, h, w):
        k, s, p = self.kh, self.stride, self.pad
        if self.kind == "cols":
            return h, w            # the source already is the im2col grid
        if self.kind in ("conv", "head"):
            return (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        return (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k

    def fwd_plans(self, srcs, out, stats_partial=None, act=ACT_NONE, slope=0.2, use_bias=True):
        """List of launches computing the layer's forward: one for Conv2d, stride^2 phase launches for
        ConvTranspose2d (each output phase is a small stride-1 conv written to a strided view)."""
        if self.kind == "conv":
            return [self.fwd_plan(srcs, out, stats_partial, act, slope, use_bias)]
        bias_t = self.bias.detach() if (use_bias and self.bias is not None) else None
        if self.kind == "cols":
            n, ho, wo, _ = out.shape
            src = [dict(act=srcs[0], wgt=self.pack_fwd, k_off=0, c_real=self.taps * self.I)]
            return [_C.conv_plan(src, out, [(0, 0, 0)], bias=bias_t, stats_partial=stats_partial, act=act, slope=slope,
                                 cout_real=self.O)]
        if self.kind == "head":
            x = srcs[0]
            n, hi, wi, _ = x.shape
            hc = self._buf(("hc", out.data_ptr()), (n, hi, wi, 64), x.device)
            ho, wo = hi - self.kh + 1, wi - self.kw + 1
            gemm = _C.conv_plan([dict(act=x, wgt=self.pack_fwd, k_off=0)], hc, [(0, 0, 0)],
                                flops=2.0 * n * ho * wo * self.taps * self.I)
            return [gemm, TailCall("head_gather", (hc, bias_t, out), lambda: (
                _C.ptr(hc), _C.ptr(bias_t), _C.ptr(out), n, hi, wi, self.kh, out.shape[3], act, _C.F(slope)))]
        assert len(srcs) == len(self.in_split)
        s = self.stride
        srcd = [dict(act=t, wgt=self.pack_fwd, k_off=o, c_real=c) for t, o, c in zip(srcs, self.k_off, self.in_split)]
        plans = []
        bias = self.bias.detach() if (use_bias and self.bias is not None) else None
        tiles_total = stats_partial.shape[1] if stats_partial is not None else 0
        slot = tiles_total // (s * s) if stats_partial is not None else 0
        ph = 0
        for py in range(s):
            for px in range(s):
                taps = phase_taps(self.kh, self.kw, self.pad, s, py, px, flipped=False)
                sub = out[:, py::s, px::s, :]
                assert taps, "every phase of the supported transposed convs has taps"
                plans.append(_C.conv_plan(srcd, sub, taps, bias=bias, stats_partial=stats_partial, act=act,
                                          slope=slope, cout_real=self.O, stats_tiles_total=tiles_total,
                                          stats_tile_off=ph * slot))
                ph += 1
        return plans

    def fwd_plan(self, srcs, out, stats_partial=None, act=ACT_NONE, slope=0.2, use_bias=True):
        """srcs: NHWC bf16 tensors in concat order (conv) -- Conv2d forward."""
        tiles_total = stats_partial.shape[1] if stats_partial is not None else 0
        assert self.kind == "conv" and len(srcs) == len(self.in_split)
        s = [dict(act=t, wgt=self.pack_fwd, k_off=o, c_real=c) for t, o, c in zip(srcs, self.k_off, self.in_split)]
        return _C.conv_plan(s, out, conv_taps(self.kh, self.kw, self.pad), stride=self.stride,
                            bias=self.bias.detach() if (use_bias and self.bias is not None) else None,
                            stats_partial=stats_partial, act=act, slope=slope, cout_real=self.O,
                            stats_tiles_total=tiles_total)

    def wgrad_plans(self, srcs, dy):
        """Conv2d: one launch (P = concat inputs, Q = dY). ConvTranspose2d: the strided operand is dY
        (P) and the fixed one the input (Q), one launch per concat segment into its row slice of
        grad[taps][i_pad][o_pad]."""
        if self.kind == "conv":
            return [_C.wgrad_plan(list(srcs), dy, conv_taps(self.kh, self.kw, self.pad), self.grad,
                                  stride=self.stride, p_real=self.I, q_real=self.O)]
        if self.kind == "cols":
            return [_C.wgrad_plan([srcs[0]], dy, [(0, 0, 0)], self.grad, p_real=self.taps * self.I, q_real=self.O)]
        if self.kind == "head":
            x = srcs[0]
            n, hi, wi, _ = x.shape
            dzc = self._buf(("dzc", dy.data_ptr()), (n, hi, wi, 64), x.device)
            ho, wo = hi - self.kh + 1, wi - self.kw + 1
            return [TailCall("head_scatter", (dy, dzc), lambda: (_C.ptr(dy), _C.ptr(dzc), n, hi, wi, self.kh,
                                                                 dy.shape[3])),
                    _C.wgrad_plan([x], dzc, [(0, 0, 0)], self.grad, flops=2.0 * n * ho * wo * self.taps * self.I)]
        taps = conv_taps(self.kh, self.kw, self.pad)
        return [_C.wgrad_plan([dy], x, taps, self.grad, stride=self.stride, p_real=self.O, q_real=c, dw_row_off=o)
                for x, o, c in zip(srcs, self.k_off, self.in_split)]

    def _buf(self, key, shape, device):
        """Scratch tensors of the restated layers (one per call site, keyed by the tensor they belong to)."""
        if key not in self._scratch:
            self._scratch[key] = torch.zeros(*shape, dtype=torch.bfloat16, device=device)
        return self._scratch[key]

    def dgrad_src(self, dy, seg):
        """Operand descriptor for the input-gradient of concat segment `seg`."""
        return dict(act=dy, wgt=self.pack_bwd, k_off=0, row_off=self.k_off[seg], c_real=self.O)

    def dgrad_plans(self, dy, out, seg=0):
        """Input gradient w.r.t. concat segment `seg` -> list of plans (one per output phase)."""
        if self.kind == "cols":
            return [_C.conv_plan([dict(act=dy, wgt=self.pack_bwd, k_off=0, c_real=self.O)], out, [(0, 0, 0)],
                                 cout_real=self.taps * self.I)]
        if self.kind == "head":
            n, hi, wi, _ = out.shape
            dzc = self._buf(("dzc", dy.data_ptr()), (n, hi, wi, 64), out.device)
            ho, wo = hi - self.kh + 1, wi - self.kw + 1
            return [TailCall("head_scatter", (dy, dzc), lambda: (_C.ptr(dy), _C.ptr(dzc), n, hi, wi, self.kh,
                                                                 dy.shape[3])),
                    _C.conv_plan([dict(act=dzc, wgt=self.pack_bwd, k_off=0)], out, [(0, 0, 0)],
                                 flops=2.0 * n * ho * wo * self.taps * self.I)]
        if self.kind == "convT":
            # d(input) of a transposed conv is an ordinary strided conv over dY with the [tap][ci][co] pack
            return [_C.conv_plan([self.dgrad_src(dy, seg)], out, conv_taps(self.kh, self.kw, self.pad),
                                 stride=self.stride, cout_real=self.in_split[seg])]
        if self.stride == 1:
            return [_C.conv_plan([self.dgrad_src(dy, seg)], out, dgrad_taps_s1(self.kh, self.kw, self.pad),
                                 cout_real=self.in_split[seg])]
        plans = []
        s = self.stride
        for py in range(s):
            for px in range(s):
                taps = phase_taps(self.kh, self.kw, self.pad, s, py, px, flipped=True)
                sub = out[:, py::s, px::s, :]
                if sub.shape[1] == 0 or sub.shape[2] == 0:
                    continue
                if not taps:
                    sub.zero_()
                    continue
                plans.append(_C.conv_plan([self.dgrad_src(dy, seg)], sub, taps, cout_real=self.in_split[seg]))
        return plans


class TailCall:
    """A bandwidth kernel that sits in a plan list next to the implicit-GEMM plans (same .run() protocol)."""

    def __init__(self, name, keep, args):
        self.name, self.keep, self.args = name, keep, args

    def run(self):
        _C.call(self.name, *self.args())


def plan_buckets(layout, bucket_elems, tail_elems=0):
    """Split an arena layout [(param, offset, size), ...] (completion order) into contiguous buckets of at
    least `bucket_elems` elements: returns [(start, end, last_param), ...]; a bucket may be all-reduced as soon
    as the gradient of `last_param` is final. tail_elems > 0: the parameters that complete last (at most that many
    elements) get a bucket of their own -- it is the only collective nothing is left to overlap with, so it is kept
    small (the bucket before it is launched earlier, while the last units still run)."""
    layout = [(i, off, size) for i, off, size in layout if size]
    cut = len(layout)
    if tail_elems > 0 and len(layout) > 1:
        acc = 0
        while cut > 1 and acc + layout[cut - 1][2] <= tail_elems:
            cut -= 1
            acc += layout[cut][2]
    buckets, start, last = [], None, None
    for k, (i, off, size) in enumerate(layout):
        if k == cut and start is not None:          # close the running bucket in front of the tail bucket
            buckets.append((start, off, last))
            start = None
        if start is None:
            start = off
        last = i
        if k < cut and off + size - start >= bucket_elems:
            buckets.append((start, off + size, last))
            start = None
    if start is not None:
        i, off, size = layout[-1]
        buckets.append((start, off + size, last))
    return buckets


def multi_dgrad_plan(consumers, out, pool_out=False):
    """d(out) = sum over consumers (layer, dY, segment) of the stride-1 input gradients -- one
    multi-source implicit GEMM (the K loop walks the consumers). pool_out: `out` is the half-resolution
    tensor and the kernel stores the 2x2 sum (gradient through a nearest-upsampled copy)."""
    l0 = consumers[0][0]
    for layer, _, _ in consumers:
        assert layer.stride == 1 and (layer.kh, layer.kw, layer.pad) == (l0.kh, l0.kw, l0.pad)
    srcs = [layer.dgrad_src(dy, seg) for layer, dy, seg in consumers]
    return _C.conv_plan(srcs, out, dgrad_taps_s1(l0.kh, l0.kw, l0.pad), cout_real=l0.in_split[consumers[0][2]],
                        pool_out=pool_out)


class ParamStore:
    """Owns gradients / Adam state / packs for every parameter of one nn.Module, in
    module.parameters() order (the order torch.optim.Adam indexes its state by)."""

    def __init__(self, module, device):
        self.module = module
        self.device = device
        self.params = list(module.parameters())
        self.names = [n for n, _ in module.named_parameters()]
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self.conv_of = {}      # param index -> ConvLayer
        self.bias_of = {}      # param index -> ConvLayer whose bias this is
        self.layer_cache = {}  # name -> ConvLayer (shared by every engine built on this module)
        self.table = None
        self.step_count = 0
        self.generation = 0    # bumped whenever parameter storage moved: captured CUDA graphs hold raw pointers

    def register_conv(self, layer):
        self.conv_of[self.index[id(layer.weight)]] = layer
        if layer.bias is not None:
            self.bias_of[self.index[id(layer.bias)]] = layer

    def finalize(self, order=None):
        """Allocate arenas once every conv is registered. `order`: parameter indices in the order their
        gradients become final during backward; the gradient arena is laid out in that order, so the buckets
        of the overlapped data-parallel allreduce are contiguous ranges that complete front to back."""
        dev = self.device
        sizes = []
        for i, p in enumerate(self.params):
            if i in self.conv_of:
                l = self.conv_of[i]
                sizes.append(l.grad_shape[0] * l.grad_shape[1] * l.grad_shape[2])
            elif i in self.bias_of:
                sizes.append(self.bias_of[i].o_pad)
            else:
                sizes.append(p.numel())
        seen = set()
        layout = [i for i in (order or []) if not (i in seen or seen.add(i))]
        layout += [i for i in range(len(self.params)) if i not in seen]
        offs, total = {}, 0
        for i in layout:
            offs[i] = total
            total += (sizes[i] + 3) // 4 * 4
        self.arena_layout = [(i, offs[i], (sizes[i] + 3) // 4 * 4) for i in layout]   # (param, offset, padded size)
        self.grad_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad_views = []
        for i, p in enumerate(self.params):
            g = self.grad_arena[offs[i]:offs[i] + sizes[i]]
            if i in self.conv_of:
                l = self.conv_of[i]
                l.grad = g.view(*l.grad_shape)
            elif i in self.bias_of:
                self.bias_of[i].bias_grad = g
            self.grad_views.append(g)
        pn = [p.numel() for p in self.params]
        po = [0]
        for s in pn:
            po.append(po[-1] + (s + 3) // 4 * 4)
        self.exp_avg = torch.zeros(po[-1], dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(po[-1], dtype=torch.float32, device=dev)
        self.m_views = [self.exp_avg[po[i]:po[i] + pn[i]].view_as(p) for i, p in enumerate(self.params)]
        self.v_views = [self.exp_avg_sq[po[i]:po[i] + pn[i]].view_as(p) for i, p in enumerate(self.params)]
        self._ptrs = None
        self._versions = None
        self.max_numel = max(pn)
        self.refresh(force=True)

    # ---- table -------------------------------------------------------------------------------
    def _build_table(self, with_grad):
        rows = (AdamRow * len(self.params))()
        for i, p in enumerate(self.params):
            r = rows[i]
            r.param = p.data_ptr()
            r.grad = self.grad_views[i].data_ptr() if with_grad else None
            r.m = self.m_views[i].data_ptr()
            r.v = self.v_views[i].data_ptr()
            r.numel = p.numel()
            if i in self.conv_of:
                l = self.conv_of[i]
                r.kind = {"conv": 1, "convT": 2, "head": 3, "cols": 4}[l.kind]
                r.kh, r.kw = l.kh, l.kw
                r.dim1 = p.shape[1]
                r.o_pad, r.i_pad = l.o_pad, l.i_pad
                r.pack_fwd = l.pack_fwd.data_ptr()
                r.pack_bwd = l.pack_bwd.data_ptr()
                r.nseg = len(l.in_split)
                end = 0
                for s, c in enumerate(l.in_split):
                    end += c
                    r.seg_end[s] = end
                    r.seg_shift[s] = l.k_off[s] - (end - c)
            else:
                r.kind = 0
        raw = bytes(rows)
        return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)

    def refresh(self, force=False):
        """Rebuild the device table / re-pack bf16 weights if parameters were touched from outside
        (load_state_dict, init_weights, .to())."""
        ptrs = [p.data_ptr() for p in self.params]
        vers = [p._version for p in self.params]
        if force or ptrs != self._ptrs:
            for p in self.params:
                assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous(), "parameters must be fp32 CUDA"
            self.table = self._build_table(True)
            self.table_nograd = self._build_table(False)
            self._ptrs = ptrs
            self._versions = None
            self.generation += 1
        if vers != self._versions:
            self.repack()
            self._versions = vers

    def repack(self):
        _C.call("adam_step", _C.ptr(self.table_nograd), len(self.params), _C.LL(self.max_numel),
                _C.F(0.0), _C.F(0.0), _C.F(0.0), _C.F(1.0), 1, _C.F(1.0))

    def grad_of(self, param):
        """fp32 gradient buffer of a non-conv parameter (bias / affine / head), torch layout."""
        return self.grad_views[self.index[id(param)]]

    def zero_grad(self):
        self.grad_arena.zero_()

    def adam_step(self, lr, beta1, beta2=0.99, eps=1e-8, grad_scale=1.0):
        self.step_count += 1
        _C.call("adam_step", _C.ptr(self.table), len(self.params), _C.LL(self.max_numel), _C.F(lr),
                _C.F(beta1), _C.F(beta2), _C.F(eps), self.step_count, _C.F(grad_scale))

    @staticmethod
    def adam_hyper(lr, beta1, step, beta2=0.99, eps=1e-8, grad_scale=1.0):
        """The seven floats tg_adam_step_dev reads (computed like tg_adam_step does on the host)."""
        return [lr, beta1, beta2, eps, 1.0 - beta1 ** step, 1.0 - beta2 ** step, grad_scale]

    def adam_step_dev(self, hyper_dev):
        """Adam step whose scalars live in device memory (graph-captured steps); the caller advances step_count."""
        _C.call("adam_step_dev", _C.ptr(self.table), len(self.params), _C.LL(self.max_numel), _C.ptr(hyper_dev))

    # ---- gradients in torch layout (tests / autograd bridge) ---------------------------------
    def grad_as_torch(self, i):
        p = self.params[i]
        g = self.grad_views[i]
        if i in self.conv_of:
            l = self.conv_of[i]
            if l.kind == "conv":
                parts = [l.grad[:, :l.O, o:o + c] for o, c in zip(l.k_off, l.in_split)]
                return torch.cat(parts, 2).reshape(l.kh, l.kw, l.O, l.I).permute(2, 3, 0, 1).contiguous()
            if l.kind == "head":      # [1][tap][ci] -> [1][I][kh][kw]
                parts = [l.grad[0, :l.taps, o:o + c] for o, c in zip(l.k_off, l.in_split)]
                return torch.cat(parts, 1).reshape(l.kh, l.kw, l.I).permute(2, 0, 1).unsqueeze(0).contiguous()
            if l.kind == "cols":      # [1][o][tap*I + c] -> [O][I][kh][kw]
                return l.grad[0, :l.O, :l.taps * l.I].reshape(l.O, l.kh, l.kw, l.I).permute(0, 3, 1, 2).contiguous()
            parts = [l.grad[:, o:o + c, :l.O] for o, c in zip(l.k_off, l.in_split)]
            return torch.cat(parts, 1).reshape(l.kh, l.kw, l.I, l.O).permute(2, 3, 0, 1).contiguous()
        if i in self.bias_of:
            return g[:self.bias_of[i].O].clone()
        return g.view_as(p).clone()

    def grads_by_name(self):
        return {n: self.grad_as_torch(i) for i, n in enumerate(self.names)}

    def set_grad_from_torch(self, i, grad):
        """Inverse of grad_as_torch: write a torch-layout gradient (what autograd accumulated in p.grad on the
        bridge path) into the arena slot the fused Adam kernel reads; padded rows / columns stay zero."""
        g = self.grad_views[i]
        grad = grad.detach().to(device=g.device, dtype=torch.float32)
        if i in self.conv_of:
            l = self.conv_of[i]
            l.grad.zero_()
            if l.kind == "conv":
                t = grad.permute(2, 3, 0, 1).reshape(l.taps, l.O, l.I)
                c0 = 0
                for o, c in zip(l.k_off, l.in_split):
                    l.grad[:, :l.O, o:o + c] = t[:, :, c0:c0 + c]
                    c0 += c
            elif l.kind == "head":
                t = grad[0].permute(1, 2, 0).reshape(l.taps, l.I)
                c0 = 0
                for o, c in zip(l.k_off, l.in_split):
                    l.grad[0, :l.taps, o:o + c] = t[:, c0:c0 + c]
                    c0 += c
            elif l.kind == "cols":
                l.grad[0, :l.O, :l.taps * l.I] = grad.permute(0, 2, 3, 1).reshape(l.O, l.taps * l.I)
            else:
                t = grad.permute(2, 3, 0, 1).reshape(l.taps, l.I, l.O)
                c0 = 0
                for o, c in zip(l.k_off, l.in_split):
                    l.grad[:, o:o + c, :l.O] = t[:, c0:c0 + c, :]
                    c0 += c
        elif i in self.bias_of:
            g.zero_()
            g[:self.bias_of[i].O] = grad
        else:
            g.view_as(self.params[i]).copy_(grad)
