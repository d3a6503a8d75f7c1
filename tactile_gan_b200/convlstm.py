"""ConvLSTM / ConvBLSTM on the device, forward and backward (reference generators/BCDUNet.py:6-103).

The reference constructs these skip modules inside BCDUNet (BCDUNet.py:145-152) but its forward never calls them,
so they are outside the training step; this engine makes the classes usable -- and trainable -- as what their names
say. Per time step: ONE two-source implicit GEMM (cat[X, H_prev] is never materialised: X and H are the two K-loop
sources of the tcgen05 kernel, 4*C output channels + bias) and ONE gate kernel (tg_convlstm_gates). Under autograd
the sequence is one node (_SeqFn / _CellFn): the forward keeps z, c and the bf16 operands of every step, the backward
walks the steps in reverse -- tg_convlstm_gates_bwd (gates recomputed, peephole gradients accumulated), the gate conv's
bias / weight gradients (two-source wgrad GEMM) and its input gradients w.r.t. X (-> dX frame) and H_prev (-> the
recurrent dh of the previous step). There is no eager fallback: CPU tensors raise.
"""
from ctypes import c_void_p

import torch

from . import _C
from ._C import ACT_RELU, ACT_TANH, LL, ptr
from .engine import GraphEngine, bf16
from .layers import pad64


def _off(t, elems):
    """Pointer `elems` fp32 elements into tensor t."""
    return c_void_p(t.data_ptr() + 4 * int(elems))


class ConvLSTMCellEngine(GraphEngine):
    """One ConvLSTMCell for a fixed (N, H, W): static buffers + the gate conv's launch plan."""

    def __init__(self, cell, n, h, w):
        super().__init__(cell, n, h, w, backward=False)
        conv = cell.conv
        self.C = conv.out_channels // 4
        self.cin = conv.in_channels - self.C
        if tuple(cell.W_ci.shape) != (self.C, h, w):
            raise ValueError(f"ConvLSTMCell was built for frame {tuple(cell.W_ci.shape[1:])}, got {(h, w)}")
        if self.C % 8:
            raise ValueError("the ConvLSTM gate kernel needs out_channels divisible by 8")
        if conv.kernel_size[0] != conv.kernel_size[1] or conv.padding[0] != conv.padding[1]:
            raise ValueError("square kernels / paddings only")
        dev = self.device
        self.act = ACT_TANH if cell.activation is torch.tanh else ACT_RELU
        self.x = bf16(n, h, w, pad64(self.cin), device=dev)
        self.hb = bf16(n, h, w, pad64(self.C), device=dev)
        self.z = bf16(n, h, w, pad64(4 * self.C), device=dev)
        self.layer = self.conv_layer("conv", conv, [self.cin, self.C])
        self.finish()
        self.plan = self.layer.fwd_plan([self.x, self.hb], self.z)
        self.state = torch.zeros(n, self.C, h, w, device=dev)     # cell state of a running sequence
        self.hw = h * w
        self._train = {}                                           # T -> saved tensors + plans of a training sequence

    def load_h(self, h_prev):
        """H_prev: fp32 NCHW (or None = zeros) -> the conv's bf16 NHWC operand."""
        if h_prev is None:
            self.hb.zero_()
        else:
            _C.call("pack_nchw_tiled", ptr(h_prev), LL(self.C * self.hw), ptr(self.hb), self.n, self.C, self.hw,
                    self.hb.shape[3])

    def step(self, x_ptr, x_sn, c_prev_ptr, c_prev_sn, c_out_ptr, c_out_sn, h_out_ptr, h_out_sn):
        """One time step. x: fp32 NCHW time slice (pointer + image stride); self.hb holds H_prev and receives H."""
        cell = self.module
        _C.call("pack_nchw_tiled", x_ptr, LL(x_sn), ptr(self.x), self.n, self.cin, self.hw, self.x.shape[3])
        self.plan.run()
        _C.call("convlstm_gates", ptr(self.z), self.z.shape[3], ptr(cell.W_ci.detach()), ptr(cell.W_cf.detach()),
                ptr(cell.W_co.detach()), c_prev_ptr, LL(c_prev_sn), c_out_ptr, LL(c_out_sn), h_out_ptr, LL(h_out_sn),
                ptr(self.hb), self.hb.shape[3], self.n, self.hw, self.C, self.act)

    def run_sequence(self, X, out, c_off=0, reverse=False):
        """X (B,T,Cin,H,W) fp32 -> out[:, t, c_off:c_off+C] = H_t (out is (B,T,Ctot,H,W) fp32, contiguous).
        reverse: walk the frames last to first, each H written at its own frame's slot (the backward cell of ConvBLSTM,
        BCDUNet.py:97-99)."""
        b, t, cin, h, w = X.shape
        assert (b, cin, h, w) == (self.n, self.cin, self.h, self.w) and X.is_contiguous() and out.is_contiguous()
        ctot = out.shape[2]
        self.store.refresh()
        self.load_h(None)
        cn = self.C * self.hw
        for k in range(t):
            f = t - 1 - k if reverse else k
            self.step(_off(X, f * cin * self.hw), t * cin * self.hw,
                      ptr(self.state) if k else c_void_p(0), cn, ptr(self.state), cn,
                      _off(out, (f * ctot + c_off) * self.hw), t * ctot * self.hw)
        return out


    # ------------------------------------------------------------------------------------------- training
    def _train_state(self, T):
        """Buffers and launch plans of a T-step sequence that keeps what the backward needs."""
        if T in self._train:
            return self._train[T]
        dev, n, h, w = self.device, self.n, self.h, self.w
        st = {}
        st["xs"] = [bf16(*self.x.shape, device=dev) for _ in range(T)]
        st["hs"] = [bf16(*self.hb.shape, device=dev) for _ in range(T + 1)]       # hs[k] = H entering step k
        st["zs"] = [bf16(*self.z.shape, device=dev) for _ in range(T)]
        st["cs"] = torch.zeros(T, n, self.C, h, w, device=dev)                     # cell state after step k
        st["fwd"] = [self.layer.fwd_plan([st["xs"][k], st["hs"][k]], st["zs"][k]) for k in range(T)]
        st["dz"] = bf16(*self.z.shape, device=dev)
        st["dxb"] = bf16(*self.x.shape, device=dev)
        st["dhb"] = bf16(*self.hb.shape, device=dev)
        st["dc"] = torch.zeros(n, self.C, h, w, device=dev)
        st["wgrad"] = [self.layer.wgrad_plans([st["xs"][k], st["hs"][k]], st["dz"]) for k in range(T)]
        st["dgrad_x"] = self.layer.dgrad_plans(st["dz"], st["dxb"], 0)
        st["dgrad_h"] = self.layer.dgrad_plans(st["dz"], st["dhb"], 1)
        self._train[T] = st
        return st

    def forward_train(self, X, out, c_off=0, reverse=False, h0=None, c0=None):
        """Like run_sequence, keeping z / c / the bf16 operands of every step. X (B,T,Cin,H,W) fp32 contiguous;
        h0 / c0: initial state (fp32 NCHW) or None = zeros. Returns the per-step state (for backward_train)."""
        b, t, cin, h, w = X.shape
        assert (b, cin, h, w) == (self.n, self.cin, self.h, self.w) and X.is_contiguous() and out.is_contiguous()
        st = self._train_state(t)
        cell = self.module
        ctot = out.shape[2]
        self.store.refresh()
        if h0 is None:
            st["hs"][0].zero_()
        else:
            _C.call("pack_nchw_tiled", ptr(h0), LL(self.C * self.hw), ptr(st["hs"][0]), self.n, self.C, self.hw,
                    self.hb.shape[3])
        cn = self.C * self.hw
        for k in range(t):
            f = t - 1 - k if reverse else k
            _C.call("pack_nchw_tiled", _off(X, f * cin * self.hw), LL(t * cin * self.hw), ptr(st["xs"][k]), self.n,
                    self.cin, self.hw, self.x.shape[3])
            st["fwd"][k].run()
            c_prev = ptr(st["cs"][k - 1]) if k else ptr(c0)
            _C.call("convlstm_gates", ptr(st["zs"][k]), self.z.shape[3], ptr(cell.W_ci.detach()),
                    ptr(cell.W_cf.detach()), ptr(cell.W_co.detach()), c_prev, LL(cn), ptr(st["cs"][k]), LL(cn),
                    _off(out, (f * ctot + c_off) * self.hw), LL(t * ctot * self.hw), ptr(st["hs"][k + 1]),
                    self.hb.shape[3], self.n, self.hw, self.C, self.act)
        st["c0"], st["T"] = c0, t
        return st

    def backward_train(self, d_out, dX, c_off=0, reverse=False, d_c_last=None, want_h0=False, accumulate_dx=False):
        """Back-propagation through time of the sequence forward_train just ran. d_out (B,T,Ctot,H,W) fp32 contiguous:
        gradient of the H frames (this cell's channels at c_off); d_c_last: gradient of the final cell state or None;
        dX (B,T,Cin,H,W) receives (or accumulates) the input gradient. Parameter gradients land in the store's arena
        (zeroed here). Returns (dH0 as bf16 NHWC or None, dC0 fp32)."""
        st = self._train[d_out.shape[1]]
        t = st["T"]
        cell = self.module
        gs = self.store
        gs.zero_grad()
        ctot = d_out.shape[2]
        cn = self.C * self.hw
        zc, hc = self.z.shape[3], self.hb.shape[3]
        dc = st["dc"]
        if d_c_last is not None:
            dc.copy_(d_c_last)
        for k in range(t - 1, -1, -1):
            f = t - 1 - k if reverse else k
            last = k == t - 1
            c_prev = ptr(st["cs"][k - 1]) if k else ptr(st["c0"])
            _C.call("convlstm_gates_bwd", ptr(st["zs"][k]), zc, ptr(cell.W_ci.detach()), ptr(cell.W_cf.detach()),
                    ptr(cell.W_co.detach()), c_prev, LL(cn), ptr(st["cs"][k]), LL(cn),
                    _off(d_out, (f * ctot + c_off) * self.hw), LL(t * ctot * self.hw),
                    None if last else ptr(st["dhb"]), hc,
                    ptr(dc) if (not last or d_c_last is not None) else None, ptr(dc), ptr(st["dz"]),
                    ptr(gs.grad_of(cell.W_ci)), ptr(gs.grad_of(cell.W_cf)), ptr(gs.grad_of(cell.W_co)),
                    self.n, self.hw, self.C, self.act)
            _C.call("bias_grad", ptr(st["dz"]), ptr(self.layer.bias_grad), LL(self.n * self.hw), zc, 4 * self.C)
            for p in st["wgrad"][k]:
                p.run()
            for p in st["dgrad_x"]:
                p.run()
            _C.call("unpack_nhwc_tiled", ptr(st["dxb"]), self.x.shape[3], _off(dX, f * self.cin * self.hw),
                    LL(t * self.cin * self.hw), self.n, self.cin, self.hw, int(accumulate_dx))
            if k or want_h0:
                for p in st["dgrad_h"]:
                    p.run()
        return (st["dhb"] if want_h0 else None), dc

    def param_grads(self):
        return [self.store.grad_as_torch(i) for i in range(len(self.store.params))]


class _SeqFn(torch.autograd.Function):
    """One ConvLSTM over a (B,T,Cin,H,W) sequence, zero initial state (BCDUNet.py:61-84), as a single autograd node."""

    @staticmethod
    def forward(ctx, X, eng, reverse, *params):
        b, t = X.shape[:2]
        out = torch.empty(b, t, eng.C, eng.h, eng.w, device=X.device)
        eng.forward_train(X.detach().contiguous().float(), out, 0, reverse)
        ctx.eng, ctx.reverse, ctx.xshape = eng, reverse, tuple(X.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        eng = ctx.eng
        dX = torch.empty(ctx.xshape, device=g.device)
        eng.backward_train(g.contiguous().float(), dX, 0, ctx.reverse)
        return (dX, None, None, *eng.param_grads())


class _CellFn(torch.autograd.Function):
    """ConvLSTMCell.forward(X, H_prev, C_prev) -> (H, C) (BCDUNet.py:32-47) as a single autograd node."""

    @staticmethod
    def forward(ctx, X, H_prev, C_prev, eng, *params):
        n = X.shape[0]
        out = torch.empty(n, 1, eng.C, eng.h, eng.w, device=X.device)
        c0 = C_prev.detach().contiguous().float()
        st = eng.forward_train(X.detach().contiguous().float().unsqueeze(1), out, 0, False,
                               h0=H_prev.detach().contiguous().float(), c0=c0)
        ctx.eng = eng
        return out[:, 0], st["cs"][0].clone()

    @staticmethod
    def backward(ctx, dH, dC):
        eng = ctx.eng
        n = eng.n
        d_out = (dH if dH is not None else torch.zeros(n, eng.C, eng.h, eng.w, device=eng.device))
        d_out = d_out.contiguous().float().unsqueeze(1).contiguous()
        dX = torch.empty(n, 1, eng.cin, eng.h, eng.w, device=eng.device)
        dhb, dc = eng.backward_train(d_out, dX, 0, False, d_c_last=None if dC is None else dC.contiguous().float(),
                                     want_h0=True)
        dH0 = torch.empty(n, eng.C, eng.h, eng.w, device=eng.device)
        _C.call("unpack_nhwc_tiled", ptr(dhb), dhb.shape[3], ptr(dH0), LL(eng.C * eng.hw), n, eng.C, eng.hw, 0)
        return (dX[:, 0], dH0, dc.clone(), None, *eng.param_grads())


def _wants_grad(x, module, *more):
    return torch.is_grad_enabled() and (x.requires_grad or any(t.requires_grad for t in more) or
                                        any(p.requires_grad for p in module.parameters()))


def _check(x, module=None):
    if not x.is_cuda:
        raise _C.TgError("tactile_gan_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")


def _engine(cell, n, h, w):
    cache = cell.__dict__.setdefault("_tg_engines", {})
    if (n, h, w) not in cache:
        cache[(n, h, w)] = ConvLSTMCellEngine(cell, n, h, w)
    return cache[(n, h, w)]


def cell_forward(cell, X, H_prev, C_prev):
    """ConvLSTMCell.forward(X, H_prev, C_prev) -> (H, C), all fp32 NCHW (BCDUNet.py:32-47)."""
    _check(X, cell)
    n, _, h, w = X.shape
    eng = _engine(cell, n, h, w)
    if _wants_grad(X, cell, H_prev, C_prev):
        return _CellFn.apply(X, H_prev, C_prev, eng, *cell.parameters())
    X = X.detach().contiguous().float()
    H_prev = H_prev.detach().contiguous().float()
    C_prev = C_prev.detach().contiguous().float()
    eng.store.refresh()
    eng.load_h(H_prev)
    H = torch.empty(n, eng.C, h, w, device=X.device)
    C = torch.empty_like(H)
    cn = eng.C * eng.hw
    eng.step(ptr(X), eng.cin * eng.hw, ptr(C_prev), cn, ptr(C), cn, ptr(H), cn)
    return H, C


def lstm_forward(lstm, X):
    """ConvLSTM.forward (BCDUNet.py:61-84): zero initial state, unrolled over dim 1."""
    _check(X, lstm)
    b, t, _, h, w = X.shape
    eng = _engine(lstm.convLSTMcell, b, h, w)
    if _wants_grad(X, lstm):
        out = _SeqFn.apply(X, eng, False, *lstm.convLSTMcell.parameters())
        return out if lstm.return_sequence else out[:, -1]
    X = X.detach().contiguous().float()
    out = torch.empty(b, t, lstm.out_channels, h, w, device=X.device)
    eng.run_sequence(X, out)
    return out if lstm.return_sequence else out[:, -1]


def blstm_forward(blstm, x):
    """ConvBLSTM.forward (BCDUNet.py:96-103): forward cell on the frames, backward cell on the reversed frames
    (un-reversed again), concatenated on the channel axis -- both cells write straight into their channel halves."""
    _check(x, blstm)
    b, t, _, h, w = x.shape
    ef = _engine(blstm.forward_cell.convLSTMcell, b, h, w)
    eb = _engine(blstm.backward_cell.convLSTMcell, b, h, w)
    if _wants_grad(x, blstm):
        # two autograd nodes (one per direction; the backward cell walks the frames last to first and writes each H at
        # its own frame, BCDUNet.py:97-99); torch.cat joins the channel halves and routes the gradient slices back
        out = torch.cat((_SeqFn.apply(x, ef, False, *blstm.forward_cell.convLSTMcell.parameters()),
                         _SeqFn.apply(x, eb, True, *blstm.backward_cell.convLSTMcell.parameters())), dim=2)
        return out if blstm.return_sequence else out[:, -1]
    x = x.detach().contiguous().float()
    out = torch.empty(b, t, ef.C + eb.C, h, w, device=x.device)
    ef.run_sequence(x, out, c_off=0)
    eb.run_sequence(x, out, c_off=ef.C, reverse=True)
    return out if blstm.return_sequence else out[:, -1]
