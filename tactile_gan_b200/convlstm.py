"""ConvLSTM / ConvBLSTM forward on the device (reference generators/BCDUNet.py:6-103).

The reference constructs these skip modules inside BCDUNet (BCDUNet.py:145-152) but its forward never calls them,
so they are outside the training step; this engine makes the classes usable as what their names say, forward only
(inference): per time step ONE two-source implicit GEMM (cat[X, H_prev] is never materialised: X and H are the two
K-loop sources of the tcgen05 kernel, 4*C output channels + bias) and ONE gate kernel (tg_convlstm_gates).
There is no eager fallback: CPU tensors raise.
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


def _check(x, module=None):
    if not x.is_cuda:
        raise _C.TgError("tactile_gan_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
    if torch.is_grad_enabled() and (x.requires_grad or (module is not None and
                                                        any(p.requires_grad for p in module.parameters()))):
        # no training path reaches these modules (BCDUNet.forward never calls them), so only the forward exists:
        # refuse to hand back tensors that silently carry no graph
        raise NotImplementedError("the ConvLSTM engine is forward-only: call it under torch.no_grad() "
                                  "(or with requires_grad_(False) parameters)")


def _engine(cell, n, h, w):
    cache = cell.__dict__.setdefault("_tg_engines", {})
    if (n, h, w) not in cache:
        cache[(n, h, w)] = ConvLSTMCellEngine(cell, n, h, w)
    return cache[(n, h, w)]


def cell_forward(cell, X, H_prev, C_prev):
    """ConvLSTMCell.forward(X, H_prev, C_prev) -> (H, C), all fp32 NCHW (BCDUNet.py:32-47). Forward only."""
    _check(X, cell)
    n, _, h, w = X.shape
    eng = _engine(cell, n, h, w)
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
    x = x.detach().contiguous().float()
    out = torch.empty(b, t, ef.C + eb.C, h, w, device=x.device)
    ef.run_sequence(x, out, c_off=0)
    eb.run_sequence(x, out, c_off=ef.C, reverse=True)
    return out if blstm.return_sequence else out[:, -1]
