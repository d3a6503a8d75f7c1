"""Static execution graphs for the generators: every buffer and kernel plan is created once for a
fixed (N, H, W); forward / backward are then straight sequences of kernel launches on the current
CUDA stream (capturable in a CUDA graph). No autograd, no torch math ops on the hot path.

Graph vocabulary
  Act        one NHWC bf16 activation tensor + who consumes it (for the fused multi-consumer dgrad)
  ConvUnit   conv (implicit GEMM) -> [InstanceNorm(+affine)] -> activation, with optional pooled and
             2x-upsampled copies of the output written by the same normalise pass
  HeadUnit   FeatureMapBlock: 1x1 conv + bias (+tanh) -> fp32 NCHW
"""
import os

import torch

from . import _C
from ._C import ACT_LRELU, ACT_NONE, ACT_RELU, F, LL, ptr
from .layers import ConvLayer, ParamStore, multi_dgrad_plan, pad64

EPS_IN = 1e-5


def bf16(*shape, device):
    return torch.zeros(*shape, dtype=torch.bfloat16, device=device)


class Act:
    def __init__(self, name, buf, c_real):
        self.name, self.buf, self.c = name, buf, c_real
        self.consumers = []   # (ConvUnit, segment index)
        self.grad_plans = None
        self.grad_buf = None

    @property
    def shape(self):
        return self.buf.shape

    def build_grad(self, scratch, scratch2=None, pooled=False):
        """Plan d(self) = sum of the input-gradients of all consuming convs. All stride-1 Conv2d consumers
        share ONE launch (the K loop of the implicit GEMM walks the consumers); a strided Conv2d or a
        ConvTranspose2d consumer gets its own launch(es) into a second buffer that is then added.
        pooled (nearest-upsampled copies): the gradient is produced directly at the source's resolution -- the
        conv epilogue sums each 2x2 block -- instead of at this 4x larger tensor's."""
        if not self.consumers:
            return
        n, h, w, c = self.buf.shape
        units = [(u.layer, u.dz, seg) for u, seg in self.consumers]
        same = [x for x in units if x[0].kind == "conv" and x[0].stride == 1]
        other = [x for x in units if not (x[0].kind == "conv" and x[0].stride == 1)]
        self.grad_plans, self.grad_adds = [], []
        self.grad_pooled = bool(pooled and same and not other and h % 2 == 0 and w % 2 == 0)
        if self.grad_pooled:
            self.grad_buf = scratch[: n * (h // 2) * (w // 2) * c].view(n, h // 2, w // 2, c)
            self.grad_plans.append(multi_dgrad_plan(same, self.grad_buf, pool_out=True))
            return
        self.grad_buf = scratch[: n * h * w * c].view(n, h, w, c)
        if same:
            self.grad_plans.append(multi_dgrad_plan(same, self.grad_buf))
        for i, (l, dz, seg) in enumerate(other):
            if not same and i == 0:
                self.grad_plans += l.dgrad_plans(dz, self.grad_buf, seg)
            else:
                assert scratch2 is not None and len(other) - (0 if same else 1) <= 1, "one extra gradient route"
                buf2 = scratch2[: n * h * w * c].view(n, h, w, c)
                self.grad_adds.append((l.dgrad_plans(dz, buf2, seg), buf2))

    def run_grad(self):
        for p in self.grad_plans:
            p.run()
        for plans, buf2 in self.grad_adds:
            for p in plans:
                p.run()
            _C.call("add", ptr(self.grad_buf), ptr(buf2), ptr(self.grad_buf), LL(buf2.numel()))
        return self.grad_buf


class ConvUnit:
    def __init__(self, eng, name, layer, srcs, norm, gamma=None, beta=None, act=ACT_RELU, slope=0.2, pool=0,
                 up=False):
        dev = eng.device
        self.eng, self.name, self.layer, self.srcs = eng, name, layer, srcs
        self.norm, self.gamma, self.beta, self.act, self.slope = norm, gamma, beta, act, slope
        n, h, w, _ = srcs[0].shape
        self.ho, self.wo = layer.out_hw(h, w)
        self.n, self.c, self.c_valid = n, layer.o_pad, layer.O
        shp = (n, self.ho, self.wo, self.c)
        self.y = Act(name + ".y", bf16(*shp, device=dev), layer.O)
        self.dz = bf16(*shp, device=dev) if eng.with_backward else None   # gradient w.r.t. raw: training engines only
        self.pool_mode, self.pool, self.up = pool, None, None
        if pool:
            self.pool = Act(name + ".pool", bf16(n, self.ho // 2, self.wo // 2, self.c, device=dev), layer.O)
        if up:
            self.up = Act(name + ".up", bf16(n, self.ho * 2, self.wo * 2, self.c, device=dev), layer.O)
        for i, t in enumerate(srcs):
            t.consumers.append((self, i))
        self.raw = bf16(*shp, device=dev) if norm else None
        if norm:
            # statistics come from the conv epilogue when a 128-pixel tile stays inside one image; tiny maps
            # (UNet's 2x2 .. 8x8 levels) use a direct reduction instead
            ph = layer.stride if layer.kind == "convT" else 1
            gh, gw = self.ho // ph, self.wo // ph
            self.epi_stats = gh * gw >= 128
            self.tpi = _C.conv_query_tiles(n, gh, gw, True)[3] * ph * ph if self.epi_stats else 1
            self.partial = torch.zeros(n, self.tpi, self.c, 2, device=dev) if self.epi_stats else None
            self.mr = torch.zeros(n, self.c, 2, device=dev)
            self.red = None          # [n][c][2] backward sums: a slice of the engine's red arena (GraphEngine.finish)
        eng.units.append(self)

    def build(self, backward):
        srcs = [t.buf for t in self.srcs]
        _C.SPLITK_SLOT = getattr(self, "ws_slot", 0)     # units launched on a side stream: that stream's workspace
        try:
            if self.norm:
                self.fwd_plans = self.layer.fwd_plans(srcs, self.raw, stats_partial=self.partial)
            else:
                self.fwd_plans = self.layer.fwd_plans(srcs, self.y.buf, act=self.act, slope=self.slope)
        finally:
            _C.SPLITK_SLOT = 0
        if backward:
            eng = self.eng
            self.wgrad_plans = self.layer.wgrad_plans(srcs, self.dz)
            self.y.build_grad(eng.g_scratch, eng.g2_scratch)
            if self.pool:
                self.pool.build_grad(eng.gp_scratch)
            if self.up:
                self.up.build_grad(eng.gu_scratch, pooled=True)

    def _aff(self):
        g = self.gamma.detach() if self.gamma is not None else None
        b = self.beta.detach() if self.beta is not None else None
        return ptr(g), ptr(b)

    # ---- forward
    def forward(self):
        for p in self.fwd_plans:
            p.run()
        if self.norm:
            n, ho, wo, c = self.n, self.ho, self.wo, self.c
            g, b = self._aff()
            if self.epi_stats:
                _C.call("in_finalize", ptr(self.partial), ptr(self.mr), n, self.tpi, c, ho * wo, F(EPS_IN))
            else:
                _C.call("in_stats_direct", ptr(self.raw), ptr(self.mr), n, ho * wo, c, F(EPS_IN))
            # algorithmic traffic: raw in, y out (+ the pooled quarter-size / upsampled 4x-size copies)
            nb = 2.0 * n * ho * wo * c * (2 + (0.25 if self.pool else 0) + (4 if self.up else 0))
            _C.call("in_act_fwd", ptr(self.raw), ptr(self.mr), g, b, ptr(self.y.buf),
                    ptr(self.pool.buf if self.pool else None), self.pool_mode,
                    ptr(self.up.buf if self.up else None), n, ho, wo, c, self.c_valid, self.act, F(self.slope),
                    nbytes=nb)
        elif self.pool:
            _C.call("pool_fwd", ptr(self.y.buf), ptr(self.pool.buf), self.n, self.ho, self.wo, self.c, self.pool_mode)

    # ---- backward. The gradient of y = same-res consumers (+ explicit extra) + pooled copy + upsampled copy.
    def backward(self, g_extra=None, wgrad=True, keep_dn=False):
        eng = self.eng
        dn = self.dn_keep if (keep_dn and self.norm) else eng.dn_scratch
        n, ho, wo, c = self.n, self.ho, self.wo, self.c
        g_same = self.y.run_grad() if self.y.consumers else None
        if g_extra is not None:
            assert g_same is None, "explicit output gradient only for units without conv consumers"
            g_same = g_extra
        g_pool = self.pool.run_grad() if (self.pool and self.pool.consumers) else None
        g_up = self.up.run_grad() if (self.up and self.up.consumers) else None
        up_pooled = int(bool(g_up is not None and self.up.grad_pooled))
        # opt-in (TG_WGRAD_DEFER=1 / TG_SLIM=1; a no-op otherwise): the previous unit's weight gradient goes
        # to the side stream AFTER this unit's input-gradient GEMMs were queued, so that it is the GEMM running while the
        # passes below do -- and those take the slim form that fits on an SM beside its CTAs
        slim = eng._flush_wgrad(2.0 * n * ho * wo * c * 5 if self.norm else 0.0)
        g, b = self._aff()
        if self.norm:
            if not eng.red_clean:
                self.red.zero_()               # stand-alone call; the engines clear the whole arena once per pass
            elems = 2.0 * n * ho * wo * c      # bytes of one bf16 tensor of this unit
            routes = ((1 if g_same is not None else 0) + (0.25 if g_pool is not None else 0) +
                      ((1 if up_pooled else 4) if g_up is not None else 0))
            # max-pool routing re-reads the 2x2 window of y per pixel: too dear to do twice, keep dn for that route
            stored = g_pool is not None and self.pool_mode == 2
            keep = (keep_dn and self.norm) or stored
            routes += 4 if stored else 0
            prev_slim = _C.lib().tg_in_stream_slim(1) if slim else None
            # pass 1: statistics of dn = g * act'(n) (dn itself is stored only for the GP double backward);
            # pass 2: the same loads again, dn recomputed, dz written -- 5 tensor-sizes instead of 6
            _C.call("in_bwd_reduce", ptr(self.raw), ptr(self.y.buf), ptr(self.mr), g, b, ptr(g_same), ptr(g_pool),
                    self.pool_mode, ptr(g_up), up_pooled, ptr(dn) if keep else None, ptr(self.red), n, ho, wo, c,
                    self.c_valid, self.act, F(self.slope), nbytes=elems * (1 + routes + (1 if keep else 0)))
            aff = wgrad and self.gamma is not None
            dg = ptr(eng.store.grad_of(self.gamma)) if aff else None
            db = ptr(eng.store.grad_of(self.beta)) if aff else None
            if stored:
                _C.call("in_bwd_apply", ptr(dn), ptr(self.raw), ptr(self.mr), g, ptr(self.red), ptr(self.dz), n,
                        ho * wo, c, self.c_valid, dg, db, nbytes=3 * elems)
            else:
                _C.call("in_bwd_apply_re", ptr(self.raw), ptr(self.y.buf), ptr(self.mr), g, b, ptr(g_same),
                        ptr(g_pool), self.pool_mode, ptr(g_up), up_pooled, ptr(self.red), ptr(self.dz), n, ho, wo, c,
                        self.c_valid, self.act, F(self.slope), dg, db, nbytes=elems * (2 + routes))
            if slim:
                _C.lib().tg_in_stream_slim(prev_slim)
        else:
            _C.call("in_bwd_reduce", None, ptr(self.y.buf), None, None, None, ptr(g_same), ptr(g_pool),
                    self.pool_mode, ptr(g_up), up_pooled, ptr(self.dz), None, n, ho, wo, c, self.c_valid, self.act,
                    F(self.slope))
        if wgrad:
            ws = getattr(eng, "wgrad_stream", None)
            if ws is None:
                self._weight_grads()
            else:
                # the weight gradient only needs dz (final now) and the unit's inputs: it leaves the dependency
                # chain of backward and runs on a side stream, beside the bandwidth-bound passes of the next unit
                ev = torch.cuda.Event()
                ev.record()
                if eng.defer_wgrad:
                    eng._pending_wgrad = [self, ev, None, False]
                else:
                    with torch.cuda.stream(ws):
                        ws.wait_event(ev)
                        self._weight_grads()

    def _weight_grads(self):
        if self.layer.bias is not None:
            _C.call("bias_grad", ptr(self.dz), ptr(self.layer.bias_grad), LL(self.n * self.ho * self.wo), self.c,
                    self.c_valid)
        for p in self.wgrad_plans:
            p.run()


class HeadUnit:
    """1x1 conv + bias (+tanh) on a 64-channel (padded) input -> fp32 NCHW output."""

    def __init__(self, eng, name, weight, bias, src, use_tanh):
        self.eng, self.src, self.weight, self.bias, self.use_tanh = eng, src, weight, bias, use_tanh
        n, h, w, c = src.shape
        self.co, self.ci = weight.shape[0], weight.shape[1]
        if c != 64 or self.co > 4:
            raise ValueError(f"the FeatureMapBlock kernel serves num_filter <= 64 and output_dim <= 4 "
                             f"(got {self.ci} -> {self.co})")
        self.n, self.hw = n, h * w
        self.out = torch.zeros(n, self.co, h, w, device=eng.device)
        self.dx = bf16(n, h, w, c, device=eng.device)
        self.dwpad = torch.zeros(32, self.co, 64, device=eng.device)   # 32 replicas: see tg_fmap_bwd; kept zero between passes

    def forward(self):
        _C.call("fmap_fwd", ptr(self.src.buf), ptr(self.weight.detach()), ptr(self.bias.detach()), ptr(self.out),
                self.n, self.hw, 64, self.co, int(self.use_tanh), self.ci)
        return self.out

    def backward(self, g1, g2=None, wgrad=True):
        st = self.eng.store
        _C.call("fmap_bwd", ptr(self.src.buf), ptr(self.weight.detach()), ptr(self.out), ptr(g1), ptr(g2),
                ptr(self.dx), ptr(self.dwpad), self.dwpad.shape[0], ptr(st.grad_of(self.bias)), self.n, self.hw, 64,
                self.co, int(self.use_tanh), self.ci)
        # replicas -> the weight's gradient slot (and cleared for the next pass), one launch
        _C.call("fmap_wgrad_fold", ptr(self.dwpad), self.dwpad.shape[0], self.co, self.ci,
                ptr(st.grad_of(self.weight)) if wgrad else None)
        return self.dx


class GraphEngine:
    """Common machinery: parameter store, units in execution order, scratch buffers."""

    def __init__(self, module, n, h, w, backward=True):
        self.module = module
        self.device = next(module.parameters()).device
        if self.device.type != "cuda":
            raise _C.TgError("the tactile-gan hot path only exists on CUDA (sm_100a); there is no CPU fallback")
        self.n, self.h, self.w = n, h, w
        self.with_backward = backward
        self.store = getattr(module, "_tg_store", None)
        self.own_store = self.store is None
        if self.own_store:
            self.store = ParamStore(module, self.device)
        self.units = []
        self.layers = {}
        self.red_arena = None
        self.red_clean = False      # True while a pass that cleared the whole red arena is running
        # side-stream weight gradients (TrainStep sets wgrad_stream) can be queued one unit late, behind the next unit's
        # input-gradient GEMMs, with the InstanceNorm passes beside them in the slim form (tg_in_stream_slim): measured
        # on the headline step at 821 vs 826 / 810 img/s and 825.0 vs 825.5 (ABAB) -- inside box noise -- so both are opt-in
        self._pending_wgrad = None
        self.defer_wgrad = os.environ.get("TG_WGRAD_DEFER", "0") != "0"
        self.slim_tail = os.environ.get("TG_SLIM", "0") != "0"
        self.slim_min_ratio = float(os.environ.get("TG_SLIM_MIN_RATIO", "1.0"))

    def clear_red(self):
        """One memset for the backward sums of every unit; ConvUnit.backward then skips its own clear."""
        self.red_arena.zero_()
        self.red_clean = True

    def conv_layer(self, name, conv, in_split, kind="conv"):
        """One ConvLayer per nn.Conv2d, shared by every engine built on the same module."""
        cache = self.store.layer_cache
        if isinstance(conv, torch.nn.ConvTranspose2d):
            assert kind == "conv"
            kind = "convT"
        if name not in cache:
            l = ConvLayer(name, conv.weight, conv.bias, kind, conv.stride[0], conv.padding[0], in_split,
                          self.device)
            self.store.register_conv(l)
            cache[name] = l
        return cache[name]

    PDL_MAX_PIXELS = 8 * 256 * 256     # inference batches up to this many pixels are launch / prologue bound
    MS_MAX_PIXELS = 4 * 256 * 256      # ... and up to this many leave SMs idle: UNet++ walks its chains on three streams

    def _pdl(self):
        """Context: programmatic dependent launch for the launches of a small inference pass (batch 1 at 256^2: 819 ->
        922 img/s eager, 922 -> 960 through the graph; the training step measured -1 % and keeps it off)."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            on = (not self.with_backward) and self.n * self.h * self.w <= self.PDL_MAX_PIXELS
            prev = _C.lib().tg_pdl_policy(2) if on else None
            try:
                yield
            finally:
                if on:
                    _C.lib().tg_pdl_policy(prev)
        return ctx()

    def forward_graphed(self, x):
        """Inference forward (test.py:202-203) replayed from a CUDA graph: the ~370 launches of a UNet++ forward are
        captured once (static input / output buffers), which is what bounds small batches -- 2.3 ms of launches for
        0.6 ms of kernels at batch 1. The weight re-pack check stays outside the graph."""
        assert x.shape == (self.n, self.cin, self.h, self.w) and x.dtype == torch.float32
        self.store.refresh()
        if getattr(self, "_graph", None) is not None and self._graph_gen != self.store.generation:
            self._graph = None       # parameters were re-allocated (.to(), .float(), p.data = ...): the graph's pointers are stale
        if getattr(self, "_graph", None) is None:
            if _C.TIMING["on"] or _C.TIMING["tail"]:
                return self._forward_launches(x.contiguous())     # per-launch events cannot be captured
            self._static_in = torch.empty_like(x, memory_format=torch.contiguous_format)
            self._static_in.copy_(x)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                           # one eager pass before capture
                self._forward_launches(self._static_in)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph), self._pdl():
                self._graph_out = self._forward_launches(self._static_in)
            self._graph = graph
            self._graph_gen = self.store.generation
        self._static_in.copy_(x)
        self._graph.replay()
        return self._graph_out

    def _flush_wgrad(self, tail_bytes=0.0):
        """Queue the deferred weight gradient (ConvUnit.backward) on the side stream. Returns True when the caller's
        InstanceNorm passes (`tail_bytes` of traffic) should take the slim form, i.e. when that GEMM is long enough
        to cover them (TG_SLIM_MIN_RATIO, default 1.0, of their time at 4.5 TB/s)."""
        pend, self._pending_wgrad = self._pending_wgrad, None
        if pend is None:
            return False
        unit, ev, after_unit, done = pend
        ws = self.wgrad_stream
        with torch.cuda.stream(ws):
            ws.wait_event(ev)
            unit._weight_grads()
            if done and after_unit is not None:
                after_unit(unit)
        if not self.slim_tail or tail_bytes <= 0:
            return False
        w_est = sum(p.flops for p in unit.wgrad_plans) / 1.0e15
        return w_est >= self.slim_min_ratio * tail_bytes / 4.5e12

    def _unit_done(self, unit, after_unit):
        """after_unit(unit) runs in the stream that finalises the unit's parameter gradients."""
        pend = self._pending_wgrad
        if pend is not None and pend[0] is unit:
            pend[2], pend[3] = after_unit, True      # runs behind the deferred weight gradient (_flush_wgrad)
            return
        if after_unit is None:
            return
        ws = getattr(self, "wgrad_stream", None)
        if ws is None:
            after_unit(unit)
        else:
            with torch.cuda.stream(ws):
                after_unit(unit)

    def _join_wgrad(self):
        ws = getattr(self, "wgrad_stream", None)
        self._flush_wgrad()
        if ws is not None:
            torch.cuda.current_stream().wait_stream(ws)

    def completion_order(self):
        """Parameter indices in the order backward() finalises their gradients (head first, then the units in
        reverse execution order)."""
        idx = self.store.index
        out = []
        head = getattr(self, "head", None)
        if head is not None:
            out += [idx[id(head.weight)], idx[id(head.bias)]]
        for u in reversed([u for u in self.units if isinstance(u, ConvUnit)]):
            for p in (u.layer.weight, u.layer.bias, u.gamma, u.beta):
                if p is not None and id(p) in idx:
                    out.append(idx[id(p)])
        return out

    def finish(self):
        if self.own_store:
            self.store.finalize(self.completion_order())
            self.module._tg_store = self.store
        cu = [u for u in self.units if isinstance(u, ConvUnit)]
        # the backward sums of every InstanceNorm unit live in ONE fp32 arena, cleared by a single memset per pass
        normed = [u for u in cu if u.norm]
        self.red_arena = torch.zeros(sum(u.n * u.c * 2 for u in normed) or 1, device=self.device)
        off = 0
        for u in normed:
            u.red = self.red_arena[off:off + u.n * u.c * 2].view(u.n, u.c, 2)
            off += u.n * u.c * 2
        if self.with_backward:
            mx = max(u.dz.numel() for u in cu)
            mx_up = max([u.up.buf.numel() for u in cu if u.up] + [8])
            dev = self.device
            self.dn_scratch = bf16(mx, device=dev)
            self.g_scratch = bf16(mx, device=dev)
            self.g2_scratch = bf16(mx, device=dev)
            self.gp_scratch = bf16(mx, device=dev)
            self.gu_scratch = bf16(mx_up, device=dev)
        for u in cu:
            u.build(self.with_backward)


def _block_params(block):
    """(conv0, norm0, conv1, norm1) of a reference-style Sequential [conv, IN, act, conv, IN, act]."""
    return block[0], block[1], block[3], block[4]


def _affine(norm):
    return (norm.weight, norm.bias) if getattr(norm, "weight", None) is not None else (None, None)


def unetpp_chain_schedule(depth=5, n_streams=3):
    """Launch plan of the nested UNet++ grid on `n_streams` streams (pure host logic, tested on the CPU).
    X(i,j) reads its row X(i, k<j) and X(i+1, j-1) (X(i,0) reads X(i-1,0)): the anti-diagonals X(d,0) -> X(d-1,1) ->
    ... -> X(0,d) are serial chains, chain d one step behind chain d-1. Chain d runs on stream d % n_streams; nodes are
    queued by longest-path level, so every stream sees its nodes in dependency order, and a node only waits (events)
    for sources produced on OTHER streams. Returns (order, stream_of, waits, deps, level)."""
    nodes = [(i, 0) for i in range(depth)] + [(i, j) for j in range(1, depth) for i in range(depth - j)]
    deps = {(i, j): [(i, k) for k in range(j)] + ([(i + 1, j - 1)] if j else ([(i - 1, 0)] if i else []))
            for (i, j) in nodes}
    level = {}
    for nd in nodes:                     # `nodes` lists every source before its consumers
        level[nd] = 1 + max([level[d] for d in deps[nd]], default=-1)
    order = sorted(nodes, key=lambda ij: (level[ij], ij[0]))
    stream_of = {nd: (nd[0] + nd[1]) % n_streams for nd in nodes}
    waits = {nd: [d for d in deps[nd] if stream_of[d] != stream_of[nd]] for nd in nodes}
    return order, stream_of, waits, deps, level


class UNetPPEngine(GraphEngine):
    """UNet++ (reference generators/UNet_plusplus.py:37-86): nested grid X(i,j) of ConvBlocks.
    torch.cat and nn.Upsample never materialise a concat: each operand is a separate K-loop source;
    the 2x nearest-upsampled and 2x2 average-pooled copies are written by the producer's normalise pass."""

    def __init__(self, module, n, h, w, backward=True):
        super().__init__(module, n, h, w, backward)
        dev = self.device
        nf = module.conv0_0.layer[0].out_channels
        cin = module.conv0_0.layer[0].in_channels
        if h % 16 or w % 16:
            raise ValueError("UNet++ needs H, W divisible by 16")
        self.x_in = Act("input", bf16(n, h, w, pad64(cin), device=dev), cin)
        self.cin = cin
        ch = [nf, nf * 2, nf * 4, nf * 8, nf * 16]
        X = {}
        order = [(i, 0) for i in range(5)] + [(i, j) for j in range(1, 5) for i in range(0, 5 - j)]
        self.order = order
        for (i, j) in order:
            blk = getattr(module, f"conv{i}_{j}").layer
            c0, n0, c1, n1 = _block_params(blk)
            if j == 0:
                srcs = [self.x_in] if i == 0 else [X[i - 1, 0][1].pool]
            else:
                srcs = [X[i, k][1].y for k in range(j)] + [X[i + 1, j - 1][1].up]
            split = [t.c for t in srcs]
            l0 = self.conv_layer(f"conv{i}_{j}.0", c0, split)
            l1 = self.conv_layer(f"conv{i}_{j}.3", c1, [ch[i]])
            g0, b0 = _affine(n0)
            g1, b1 = _affine(n1)
            relu = getattr(module, "_tg_debug_act", ACT_RELU)   # bring-up only: linearised network
            u0 = ConvUnit(self, f"x{i}{j}a", l0, srcs, True, g0, b0, relu)
            pool = 1 if (j == 0 and i < 4) else 0            # AvgPool2d feeds X(i+1,0)
            up = i >= 1                                       # Upsample feeds X(i-1,j+1)
            u1 = ConvUnit(self, f"x{i}{j}b", l1, [u0.y], True, g1, b1, relu, pool=pool, up=up)
            X[i, j] = (u0, u1)
        self.X = X
        # Small inference batches leave most SMs idle in the deep levels (one to eight tiles per launch at batch 1):
        # the nested grid is walked as its five anti-diagonal chains X(d,0) -> X(d-1,1) -> ... -> X(0,d) on three
        # streams (chain d on stream d % 3), each node waiting for the events of its same-row sources; the critical
        # path is 9 of the 15 nodes. TG_INFER_STREAMS=0 keeps the single launch sequence.
        self.multi_stream = ((not backward) and n * h * w <= self.MS_MAX_PIXELS and
                             os.environ.get("TG_INFER_STREAMS", "1") != "0")
        if self.multi_stream:
            self.level_order, self.stream_of, self.waits, _, _ = unetpp_chain_schedule(5, 3)
            self._side = None
            for nd in order:
                for u in X[nd]:
                    u.ws_slot = self.stream_of[nd]
        self.head = HeadUnit(self, "downfeature", module.downfeature.conv.weight, module.downfeature.conv.bias,
                             X[0, 4][1].y, module.downfeature.activation)
        self.finish()

    def forward(self, x):
        """x: fp32 NCHW on the engine's device -> fp32 NCHW (tensor owned by the engine)."""
        assert x.shape == (self.n, self.cin, self.h, self.w) and x.dtype == torch.float32 and x.is_contiguous()
        self.store.refresh()
        with self._pdl():
            return self._forward_launches(x)

    def _forward_launches(self, x):
        _C.call("pack_nchw", ptr(x), None, None, None, ptr(self.x_in.buf), self.n, self.h * self.w, self.cin,
                self.x_in.buf.shape[3], 0)
        if not self.multi_stream:
            for (i, j) in self.order:
                u0, u1 = self.X[i, j]
                u0.forward()
                u1.forward()
            return self.head.forward()
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(2)]
        streams = [cur] + self._side
        start = torch.cuda.Event()
        start.record(cur)
        for s in self._side:
            s.wait_event(start)
        done = {}
        for nd in self.level_order:
            st = streams[self.stream_of[nd]]
            with torch.cuda.stream(st):
                for d in self.waits[nd]:           # sources produced on another stream (unetpp_chain_schedule)
                    st.wait_event(done[d])
                u0, u1 = self.X[nd]
                u0.forward()
                u1.forward()
                done[nd] = torch.cuda.Event()
                done[nd].record(st)
        for s in self._side:
            cur.wait_stream(s)
        return self.head.forward()

    def backward(self, g1, g2=None, after_unit=None):
        """g1 (+ g2): gradients w.r.t. the fp32 NCHW output. Accumulates into the store's grad arena.
        after_unit(unit) is called once a unit's parameter gradients are final (data-parallel overlap)."""
        self.clear_red()
        dx = self.head.backward(g1, g2)
        for (i, j) in reversed(self.order):
            u0, u1 = self.X[i, j]
            u1.backward(g_extra=dx if (i, j) == (0, 4) else None)
            self._unit_done(u1, after_unit)
            u0.backward()
            self._unit_done(u0, after_unit)
        self.red_clean = False
        self._join_wgrad()


class SequentialGenEngine(GraphEngine):
    """Generators whose units run in construction order (UNet, BCDUNet)."""

    def forward(self, x):
        assert x.shape == (self.n, self.cin, self.h, self.w) and x.dtype == torch.float32 and x.is_contiguous()
        self.store.refresh()
        with self._pdl():
            return self._forward_launches(x)

    def _forward_launches(self, x):
        _C.call("pack_nchw", ptr(x), None, None, None, ptr(self.x_in.buf), self.n, self.h * self.w, self.cin,
                self.x_in.buf.shape[3], 0)
        for u in self.units:
            u.forward()
        return self.head.forward()

    def backward(self, g1, g2=None, after_unit=None):
        self.clear_red()
        dx = self.head.backward(g1, g2)
        for u in reversed(self.units):
            u.backward(g_extra=dx if u is self.last else None)
            self._unit_done(u, after_unit)
        self.red_clean = False
        self._join_wgrad()

    def _double(self, name, block, srcs, pool=0):
        """[conv | convT] -> IN -> ReLU -> conv3x3 -> IN -> ReLU (reference ConvDown / DeconvUp / conv_block)."""
        c0, n0, c1, n1 = _block_params(block)
        relu = getattr(self.module, "_tg_debug_act", ACT_RELU)   # bring-up only: linearised network
        l0 = self.conv_layer(name + ".0", c0, [t.c for t in srcs])
        g0, b0 = _affine(n0)
        u0 = ConvUnit(self, name + "a", l0, srcs, True, g0, b0, relu)
        l1 = self.conv_layer(name + ".3", c1, [l0.O])
        g1, b1 = _affine(n1)
        return ConvUnit(self, name + "b", l1, [u0.y], True, g1, b1, relu, pool=pool)


class UNetEngine(SequentialGenEngine):
    """UNet (reference generators/UNet.py:55-99): 7 x ConvDown (conv k4 s2 p1 first), 7 x DeconvUp
    (ConvTranspose k4 s2 p1 first, run as 4 sub-pixel phase convs), skip concats as K-loop sources."""

    def __init__(self, module, n, h, w, backward=True):
        super().__init__(module, n, h, w, backward)
        if h % 128 or w % 128 or h < 256 or w < 256:
            raise ValueError("UNet needs H, W >= 256 and divisible by 128 (7 halvings, InstanceNorm on >1 pixel)")
        cin = module.conv1.layer[0].in_channels
        self.cin = cin
        self.x_in = Act("input", bf16(n, h, w, pad64(cin), device=self.device), cin)
        c = [None]
        t = self.x_in
        for i in range(1, 8):
            u = self._double(f"conv{i}", getattr(module, f"conv{i}").layer, [t])
            c.append(u.y)
            t = u.y
        d = c[7]
        for i in range(2, 9):
            srcs = [d] if i == 2 else [d, c[9 - i]]
            d = self._double(f"deconv{i}", getattr(module, f"deconv{i}").layer, srcs).y
        self.last = self.units[-1]
        self.head = HeadUnit(self, "downfeature", module.downfeature.conv.weight, module.downfeature.conv.bias, d,
                             module.downfeature.activation)
        self.finish()


class BCDUNetEngine(SequentialGenEngine):
    """BCDUNet (reference generators/BCDUNet.py:106-181): 4 levels of [conv3x3+bias -> IN (no affine) ->
    ReLU] x 2 with MaxPool2d between them, ConvTranspose k2 s2 (+bias) up, skip concats. The ConvLSTM modules
    are parameters only (never called by the reference forward)."""

    def __init__(self, module, n, h, w, backward=True):
        super().__init__(module, n, h, w, backward)
        if h % 8 or w % 8:
            raise ValueError("BCDUNet needs H, W divisible by 8")
        cin = module.conv1[0].in_channels
        self.cin = cin
        self.x_in = Act("input", bf16(n, h, w, pad64(cin), device=self.device), cin)
        u1 = self._double("conv1", module.conv1, [self.x_in], pool=2)
        u2 = self._double("conv2", module.conv2, [u1.pool], pool=2)
        u3 = self._double("conv3", module.conv3, [u2.pool], pool=2)
        u4 = self._double("conv4", module.conv4, [u3.pool])
        t = u4.y
        for k, skip in ((3, u3), (2, u2), (1, u1)):
            up = getattr(module, f"upconv{k}")
            lu = self.conv_layer(f"upconv{k}", up, [t.c])
            uu = ConvUnit(self, f"upconv{k}", lu, [t], False, act=ACT_NONE)
            t = self._double(f"conv{k}m", getattr(module, f"conv{k}m"), [skip.y, uu.y]).y
        self.last = self.units[-1]
        self.head = HeadUnit(self, "conv0", module.conv0.weight, module.conv0.bias, t, module.activation)
        self.finish()


class VGGFeatEngine(GraphEngine):
    """The four VGG16 slices of the reference's VGGPerceptualLoss (util.py:100-144; --version 1): input transform
    (channel repeat, ImageNet mean / std, bilinear 224x224), 10 x [conv3x3 + bias -> ReLU] with MaxPool2d(2) in
    front of slices 1-3. Weights are frozen (util.py:108-110): backward is the activation-gradient chain only,
    seeded at the four slice outputs and ending in the gradient w.r.t. the input image."""
    CONVS = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21))

    def __init__(self, blocks, n, h, w, src_channels=3, resize=True, backward=True):
        super().__init__(blocks, n, h, w, backward)
        dev = self.device
        self.cs, self.resize = src_channels, resize
        self.oh, self.ow = (224, 224) if resize else (h, w)
        if self.oh % 8 or self.ow % 8:
            raise ValueError("VGG slices need H, W divisible by 8")
        self.cin = 3
        self.x_in = Act("vgg_in", bf16(n, self.oh, self.ow, 64, device=dev), 3)
        t = self.x_in
        self.taps = []
        for b, convs in enumerate(self.CONVS):
            for j, i in enumerate(convs):
                conv = blocks[b][j * 2 + (1 if b > 0 else 0)]
                assert isinstance(conv, torch.nn.Conv2d), "unexpected VGG slice layout"
                layer = self.conv_layer(f"{b}.{i}", conv, [conv.in_channels])
                last = j == len(convs) - 1
                u = ConvUnit(self, f"vgg{b}_{i}", layer, [t], False, act=ACT_RELU, pool=2 if (last and b < 3) else 0)
                t = u.y
            self.taps.append(u)
            if b < 3:
                t = u.pool
        self.finish()
        if backward:
            self.g_feat = [bf16(*u.y.buf.shape, device=dev) for u in self.taps]
            self.dx_in = bf16(*self.x_in.buf.shape, device=dev)
            self.x_in.build_grad(self.dx_in.view(-1))

    def forward(self, x):
        """x: fp32 NCHW (n, 1|3, h, w) -> list of the four slice outputs (Act)."""
        assert x.shape == (self.n, self.cs, self.h, self.w) and x.dtype == torch.float32 and x.is_contiguous()
        self.store.refresh()
        _C.call("vgg_prep_fwd", ptr(x), ptr(self.x_in.buf), self.n, self.cs, self.h, self.w, self.oh, self.ow, 64,
                int(self.resize))
        for u in self.units:
            u.forward()
        return [u.y for u in self.taps]

    def loss_and_seed(self, other, weights, scale, loss_slot, feature_layers=(0, 1, 2, 3)):
        """*loss_slot += scale * sum_i weights[i] * L1mean(self_i, other_i); seeds d/d(self_i) for backward()."""
        self.active = []
        for i, (u, o) in enumerate(zip(self.taps, other.taps)):
            wi = float(weights[i]) * scale if i in feature_layers else 0.0
            self.active.append(wi != 0.0)
            if wi == 0.0:
                continue
            numel = u.y.buf.numel()
            _C.call("feat_loss", ptr(o.y.buf), ptr(u.y.buf), LL(numel), F(wi), 0, ptr(loss_slot))
            if self.with_backward:
                _C.call("feat_loss_grad", ptr(u.y.buf), ptr(o.y.buf), LL(numel), F(wi), ptr(self.g_feat[i]))

    def backward(self, grad_image, scale=1.0):
        """grad_image (fp32 NCHW, n x cs x h x w) += scale * d loss / d x. Units behind the last active tap are skipped."""
        tap_of = {id(u): i for i, u in enumerate(self.taps)}
        last = max([i for i, a in enumerate(self.active) if a], default=-1)
        if last < 0:
            return grad_image
        started = False
        self.clear_red()
        for u in reversed(self.units):
            i = tap_of.get(id(u))
            if not started:
                if i != last:
                    continue
                started = True
            # slice outputs feed the next slice through their pooled copy only, so the loss gradient is the sole
            # same-resolution route; an unweighted slice output just passes the pooled gradient on
            g_extra = self.g_feat[i] if (i is not None and self.active[i]) else None
            u.backward(g_extra=g_extra, wgrad=False)
        self.red_clean = False
        self.x_in.run_grad()
        _C.call("vgg_prep_bwd", ptr(self.dx_in), ptr(grad_image), self.n, self.cs, self.h, self.w, self.oh, self.ow, 64,
                int(self.resize), F(scale))
        return grad_image


def build_generator_engine(kind, module, n, h, w, backward=True):
    kind = kind.lower()
    if kind == "unet++":
        return UNetPPEngine(module, n, h, w, backward)
    if kind == "unet":
        return UNetEngine(module, n, h, w, backward)
    if kind == "bcdunet":
        return BCDUNetEngine(module, n, h, w, backward)
    raise NameError(f"{kind} has no engine")


# ======================================================================================= discriminator
class PatchDInstance(GraphEngine):
    """PatchDiscriminator (reference discriminators/PatchDiscriminator.py:5-43) for a fixed batch:
    cat(A,B) -> conv3x3 s2 +bias -> LReLU -> [conv3x3 -> IN(affine) -> LReLU] x3 (strides 2,1,1)
    -> conv3x3 +bias -> (Sigmoid). `second_order` adds the buffers / plans of the gradient-penalty
    double backward (reference util.py:72-97)."""

    def __init__(self, module, n, h, w, backward=True, second_order=False):
        super().__init__(module, n, h, w, backward)
        dev = self.device
        m = module.model
        convs = [m[0], m[2], m[5], m[8], m[11]]
        norms = [None, m[3], m[6], m[9], None]
        self.c_in = convs[0].in_channels
        self.has_sigmoid = len(m) > 12
        # The first conv (thin input) runs on im2col rows written by pack_input, the single-output head as a
        # 1x1 GEMM over its taps (layers.ConvLayer kinds "cols" / "head"): both are dense 1-tap problems.
        c0 = convs[0]
        self.k0, self.s0 = c0.kernel_size[0], c0.stride[0]
        self.cols = self.k0 * self.k0 * self.c_in <= 64 and c0.padding[0] == 0
        if not self.cols:
            raise _C.TgError("PatchDiscriminator engine expects kernel^2 * in_channels <= 64 for the first conv")
        self.h1, self.w1 = (h - self.k0) // self.s0 + 1, (w - self.k0) // self.s0 + 1
        self.x0 = Act("x0cols", bf16(n, self.h1, self.w1, 64, device=dev), self.k0 * self.k0 * self.c_in)
        self.u = []
        src = self.x0
        for k, (conv, norm) in enumerate(zip(convs, norms)):
            kind = "cols" if k == 0 else ("head" if (k == 4 and conv.out_channels == 1) else "conv")
            layer = self.conv_layer(f"model.{[0, 2, 5, 8, 11][k]}", conv, [conv.in_channels], kind=kind)
            if norm is not None:
                g, b = _affine(norm)
                unit = ConvUnit(self, f"d{k + 1}", layer, [src], True, g, b, ACT_LRELU, 0.2)
            elif k == 0:
                unit = ConvUnit(self, "d1", layer, [src], False, act=ACT_LRELU, slope=0.2)
            else:
                unit = ConvUnit(self, "d5", layer, [src], False, act=_C.ACT_SIGMOID if self.has_sigmoid else ACT_NONE)
            self.u.append(unit)
            src = unit.y
        self.second_order = second_order
        if second_order:
            for unit in self.u[1:4]:
                unit.dn_keep = bf16(*unit.dz.shape, device=dev)
        self.finish()
        self.pred = self.u[4].y.buf                      # [n, h5, w5, 64], channel 0 meaningful
        self.hw5 = self.u[4].ho * self.u[4].wo
        if backward:
            self.dx0 = bf16(*self.x0.buf.shape, device=dev)      # gradient w.r.t. the im2col rows
            self.x0.build_grad(self.dx0.view(-1))
        if second_order:
            self._build_second_order()

    # ---- forward / first-order backward
    def pack_input(self, img_a, img_b, wa=None, b2=None, wb=None, n0=0, n=None):
        """x0[n0:n0+n, ..., :ca] = A ; x0[n0:n0+n, ..., ca:ca+cb] = wa*B (+ wb*B2)  (fp32 NCHW sources;
        wa / wb are per-sample weights, e.g. the gradient-penalty interpolation of util.py:79-83)."""
        n = self.n if n is None else n
        dst = self.x0.buf[n0:n0 + n]
        ca, cb = (img_a.shape[1] if img_a is not None else self.c_in - img_b.shape[1]), img_b.shape[1]
        assert ca + cb == self.c_in
        _C.call("im2col_pack", ptr(img_a), ptr(img_b), ptr(b2), ptr(wa), ptr(wb), ptr(dst), n, ca, cb, self.h,
                self.w, self.k0, self.s0)

    def forward(self):
        self.store.refresh()
        for unit in self.u:
            unit.forward()
        return self.pred

    def backward(self, wgrad=True, input_grad=False):
        """Expects d(loss)/d(z5) in self.u[4].dz (written by the loss kernels)."""
        u5 = self.u[4]
        if wgrad:
            _C.call("bias_grad", ptr(u5.dz), ptr(u5.layer.bias_grad), LL(u5.n * u5.ho * u5.wo), u5.c, u5.c_valid)
            for p in u5.wgrad_plans:
                p.run()
        self.clear_red()
        for unit in reversed(self.u[:4]):
            unit.backward(wgrad=wgrad)
        self.red_clean = False
        if input_grad:
            self.x0.run_grad()
        return self.dx0

    def input_grad_image(self, c_off, cj, out, scale=1.0):
        """d / d(input image channels [c_off, c_off+cj)) as fp32 NCHW, folded back from the im2col-row gradient
        (valid after backward(input_grad=True) / gp_first_backward)."""
        _C.call("col2im_grad", ptr(self.dx0), ptr(out), self.n, self.c_in, c_off, cj, self.h, self.w, self.k0,
                self.s0, F(scale))
        return out

    def features(self):
        return [unit.y for unit in self.u[:4]]

    # ---- gradient penalty: second-order sweep
    def _build_second_order(self):
        dev = self.device
        u1, u2, u3, u4, u5 = self.u
        L = [x.layer for x in self.u]
        so = self.so = {}
        so["seed"] = bf16(*self.x0.buf.shape, device=dev)      # im2col rows of the second-order seed image
        so["g_img"] = None
        # every accumulator of the second-order sweep in one arena: cleared once, at the start of gp_penalty
        sizes = [self.n] + [x.n * x.c * 4 for x in self.u[1:4]] + [x.n * x.c * 2 for x in self.u[1:4]]
        so["zero_arena"] = torch.zeros(sum(sizes), device=dev)
        cuts = so["zero_arena"].split(sizes)
        so["nsq"] = cuts[0]
        so["coef"] = torch.zeros(self.n, device=dev)
        so["U"] = [bf16(*x.dz.shape, device=dev) for x in self.u]      # adj(dz_k) (U[4] = adj(dz5))
        so["V"] = [bf16(*x.dz.shape, device=dev) for x in self.u[:4]]  # adj(da_k)
        so["INJ"] = [None] + [bf16(*x.dz.shape, device=dev) for x in self.u[1:4]]
        so["red2"] = [None] + [cuts[1 + k].view(x.n, x.c, 4) for k, x in enumerate(self.u[1:4])]
        so["T5"] = bf16(*u5.dz.shape, device=dev)
        so["E"] = [bf16(*x.dz.shape, device=dev) for x in self.u[:4]]  # adj(a_k) in the downward sweep
        so["Z"] = [bf16(*x.dz.shape, device=dev) for x in self.u[:4]]  # adj(z_k) total
        so["red_dn"] = [None] + [cuts[4 + k].view(x.n, x.c, 2) for k, x in enumerate(self.u[1:4])]
        # upward: U_k = conv_k(no bias)(V_{k-1}), V_0 = seed ; wgrad(P = V_{k-1}, Q = dz_k)
        ins = [so["seed"]] + so["V"]
        so["up_conv"] = [L[k].fwd_plans([ins[k]], so["U"][k], use_bias=False) for k in range(5)]
        so["up_wgrad"] = [L[k].wgrad_plans([ins[k]], self.u[k].dz) for k in range(5)]
        # downward: wgrad(P = a_{k-1}, Q = Z_k / T5), E_{k-1} = dgrad(Z_k)
        acts_in = [self.x0.buf] + [x.y.buf for x in self.u[:4]]
        outs = so["Z"] + [so["T5"]]
        so["dn_wgrad"] = [L[k].wgrad_plans([acts_in[k]], outs[k]) for k in range(5)]
        so["dn_dgrad"] = [None] + [L[k].dgrad_plans(outs[k], so["E"][k - 1]) for k in range(1, 5)]

    def gp_first_backward(self):
        """d(sum pred)/d(x0): seeds dz5 = sigmoid'(z5) and runs the activation-gradient chain only."""
        u5 = self.u[4]
        _C.call("gp_first_seed", ptr(self.pred), int(self.has_sigmoid), 0, self.n, self.hw5, u5.c, ptr(u5.dz))
        self.clear_red()
        for unit in reversed(self.u[:4]):
            unit.backward(wgrad=False, keep_dn=True)
        self.red_clean = False
        self.x0.run_grad()
        return self.dx0

    def gp_penalty(self, c_off, cj, lambda_gp, constant, loss_slot):
        """penalty value -> *loss_slot; builds the second-order seed coef_n * (g_n + 1e-16)."""
        so = self.so
        n = self.n
        if so["g_img"] is None or so["g_img"].shape[1] != cj:
            so["g_img"] = torch.zeros(n, cj, self.h, self.w, device=self.device)
        g_img = self.input_grad_image(c_off, cj, so["g_img"])
        so["zero_arena"].zero_()        # nsq and the red2 / red_dn sums gp_second_backward accumulates into
        _C.call("gp_normsq_img", ptr(g_img), n, LL(cj * self.h * self.w), ptr(so["nsq"]))
        _C.call("gp_finish", ptr(so["nsq"]), n, F(lambda_gp), F(constant), ptr(loss_slot), ptr(so["coef"]))
        # seed image coef_n * g_n (the + 1e-16 of util.py:92 is far below bf16 resolution), as im2col rows with the
        # conditioning-image channels zero
        assert c_off + cj == self.c_in
        _C.call("im2col_pack", None, ptr(g_img), None, ptr(so["coef"]), None, ptr(so["seed"]), n, c_off, cj, self.h,
                self.w, self.k0, self.s0)

    def gp_second_backward(self):
        """Backward of the penalty through the first backward pass: accumulates weight gradients."""
        so, st = self.so, self.store
        u = self.u
        # ---------------- upward sweep through the (linearised) backward graph
        for k in range(5):
            for pl in so["up_conv"][k]:
                pl.run()                     # U_k = adj(dz_k)
            for pl in so["up_wgrad"][k]:
                pl.run()                     # from dX_{k-1} = W_k^T dz_k
            if k == 0:
                _C.call("act_bwd", ptr(so["U"][0]), ptr(u[0].y.buf), ptr(so["V"][0]), LL(so["U"][0].numel()),
                        ACT_LRELU, F(0.2))
            elif k < 4:
                x = u[k]
                g, b = x._aff()
                _C.call("in_bwd2", ptr(so["U"][k]), ptr(x.raw), ptr(x.dn_keep), ptr(x.mr), g, b, ptr(x.red),
                        ptr(so["red2"][k]), ptr(so["V"][k]), ptr(so["INJ"][k]), x.n, x.ho * x.wo, x.c, x.c_valid,
                        ACT_LRELU, F(0.2))
                if x.gamma is not None:
                    _C.call("gamma_grad2", ptr(so["red2"][k]), ptr(st.grad_of(x.gamma)), x.n, x.c, x.c_valid)
        u5 = u[4]
        _C.call("gp_top", ptr(so["U"][4]), ptr(self.pred), int(self.has_sigmoid), LL(u5.n * self.hw5), u5.c,
                ptr(so["T5"]))
        # ---------------- downward sweep: ordinary backward with the injected adj(z_k)
        _C.call("bias_grad", ptr(so["T5"]), ptr(u5.layer.bias_grad), LL(u5.n * self.hw5), u5.c, u5.c_valid)
        for pl in so["dn_wgrad"][4]:
            pl.run()
        for k in range(4, 0, -1):
            for p in so["dn_dgrad"][k]:
                p.run()                       # E_{k-1} = adj(a_{k-1})
            x = u[k - 1]
            e, z = so["E"][k - 1], so["Z"][k - 1]
            if k - 1 >= 1:
                g, b = x._aff()
                red = so["red_dn"][k - 1]
                _C.call("in_bwd_reduce", ptr(x.raw), ptr(x.y.buf), ptr(x.mr), g, b, ptr(e), None, 0, None, 0,
                        None, ptr(red), x.n, x.ho, x.wo, x.c, x.c_valid, ACT_LRELU, F(0.2))
                aff = x.gamma is not None
                _C.call("in_bwd_apply_re", ptr(x.raw), ptr(x.y.buf), ptr(x.mr), g, b, ptr(e), None, 0, None, 0,
                        ptr(red), ptr(z), x.n, x.ho, x.wo, x.c, x.c_valid, ACT_LRELU, F(0.2),
                        ptr(st.grad_of(x.gamma)) if aff else None, ptr(st.grad_of(x.beta)) if aff else None)
                _C.call("add", ptr(z), ptr(so["INJ"][k - 1]), ptr(z), LL(z.numel()))
            else:
                _C.call("act_bwd", ptr(e), ptr(x.y.buf), ptr(z), LL(z.numel()), ACT_LRELU, F(0.2))
                _C.call("bias_grad", ptr(z), ptr(x.layer.bias_grad), LL(x.n * x.ho * x.wo), x.c, x.c_valid)
            for pl in so["dn_wgrad"][k - 1]:
                pl.run()
