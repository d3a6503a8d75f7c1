"""The fused G+D training iteration (reference train.py:99-168) as one static launch sequence.

Everything between "inputs are in HBM" and "weights are updated" is a kernel of
libtactile_gan_b200.so issued on the current CUDA stream: no autograd graph, no host syncs, five loss
scalars accumulated in one device buffer. Differences from the reference that do not change results:
  * the gradient-penalty interpolate is built from fake_B without a graph to G, so the reference's
    wasted generator backward inside the D step (train.py:126,134 -> util.py:83) is not paid;
  * fake/real discriminator passes of the D step run as one 2B batch (InstanceNorm is per sample);
  * the version-2 perceptual term (pan_loss of detached D features, train.py:155-162) is evaluated
    for logging only -- it has no gradient in the reference either. Version 1 (VGG16 slices, frozen
    weights) does back-propagate into fake_B: engine.VGGFeatEngine.
Data parallel: one process per GPU; the flat fp32 gradient arenas of D and G are all-reduced (sum) with
NCCL and scaled by 1/world inside the fused Adam kernel.
"""
import os

import torch

from . import _C
from ._C import F, LL, ptr
from .engine import PatchDInstance, build_generator_engine

SLOT = {"loss_D": 0, "gp": 1, "G_GAN": 2, "L1": 3, "per": 4}


class _EventWork:
    """work.wait() of a kernel launched on the communication stream: the current stream waits for its event."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class TrainStep:
    def __init__(self, netG, netD, batch, height, width, loss="ls", version=2, lambda_a=1.0, lambda_gp=0.01,
                 lambda_per=1.0, w_per=(0, .1, .3, .6), lr=1e-3, beta1=0.9, label_smoothing=True, gen_kind=None,
                 process_group=None, vgg_blocks=None):
        self.netG, self.netD = netG, netD
        self.B, self.H, self.W = batch, height, width
        self.loss, self.version = loss, version
        self.mode = _C.GAN_MODES[loss]
        self.lambda_a, self.lambda_gp, self.lambda_per = float(lambda_a), float(lambda_gp), float(lambda_per)
        self.w_per = [float(v) for v in w_per]
        self.lr, self.beta1 = float(lr), float(beta1)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (
            torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        dev = next(netG.parameters()).device
        self.device = dev
        kind = gen_kind or netG.engine_kind
        self.G = netG._engine(batch, height, width, True) if hasattr(netG, "_engine") else \
            build_generator_engine(kind, netG, batch, height, width, True)
        self.DA = PatchDInstance(netD, 2 * batch, height, width, backward=True)
        self.S1 = PatchDInstance(netD, batch, height, width, backward=True, second_order=True)
        # version 2 needs a fifth D forward for the real-pair features (train.py:156); version 1 replaces it with
        # two passes through frozen VGG16 slices (train.py:48-49,151-153)
        self.S2 = PatchDInstance(netD, batch, height, width, backward=False) if version == 2 else None
        self.VR = self.VF = None
        if version != 2 and self.lambda_per != 0:
            if vgg_blocks is None:
                raise ValueError("version 1 needs the VGG16 slices: TrainStep(..., vgg_blocks=VGGPerceptualLoss().blocks)")
            from .engine import VGGFeatEngine
            cb = netD.model[0].in_channels - self.G.cin
            self.VR = VGGFeatEngine(vgg_blocks, batch, height, width, src_channels=cb, resize=True, backward=False)
            self.VF = VGGFeatEngine(vgg_blocks, batch, height, width, src_channels=cb, resize=True, backward=True)
        self.c_in = netD.model[0].in_channels
        self.c_a = self.G.cin
        self.c_b = self.c_in - self.c_a
        u5 = self.S1.u[4]
        self.h5, self.w5, self.c5 = u5.ho, u5.wo, u5.c
        self.losses = torch.zeros(8, device=dev)
        self.alpha = torch.zeros(batch, device=dev)
        self.one_minus_alpha = torch.zeros(batch, device=dev)
        self.gan_grad = torch.zeros(batch, self.c_b, height, width, device=dev)
        self.l1_grad = torch.zeros(batch, self.c_b, height, width, device=dev)
        if os.environ.get("TG_WGRAD_STREAM", "1") != "0":
            self.G.wgrad_stream = torch.cuda.Stream(device=dev)
        self.real_label = None
        self.label_smoothing = label_smoothing
        self.fake_B = None
        # Programmatic dependent launch for the step's launches when kernels are short enough that launch latency and
        # prologues show (measured, profiles/r02_pdl_ab.txt): UNet batch 4 +12.6 %, UNet++ batch 4 +8 %, UNet batch 32
        # +3 %; UNet++ batch 32 -1 % -> off there. TG_PDL in the environment overrides the choice.
        small = batch * height * width <= 8 * 256 * 256 or str(kind).lower() == "unet"
        self.pdl = small if os.environ.get("TG_PDL") is None else None
        # After two eager iterations the whole step is captured in a CUDA graph per `regularize` value and replayed.
        # Small batches are bound by the HOST issuing ~800 launches per step (batch 4 = the reference CLI's default:
        # 6.4 ms/step of which ~4 ms is launch cost); at batch 32 the host keeps ahead, but the ~450 graph nodes start
        # ~1.5 us sooner each than stream launches do: 39.6 -> 38.9 ms/step, +1.7 % (ABAB on one box,
        # profiles/r02_step_graph_ab.txt) -- so every single-process step is replayed from the graph. The only
        # per-step scalars -- learning rate and Adam's bias corrections -- live in device memory (tg_adam_step_dev) and
        # are rewritten before every replay; the GP alpha is drawn outside the graph into a static buffer. Single
        # process only (NCCL stays eager). TG_STEP_GRAPH=0|1 overrides.
        env = os.environ.get("TG_STEP_GRAPH")
        self.use_graph = True if env is None else env == "1"
        self.use_graph = self.use_graph and self.world == 1
        self._graphs, self._eager_done, self._hyper = {}, {}, None
        # TG_COMM_PROFILE=1 (bench.py): CUDA-event pairs around the points where the compute stream waits for a
        # gradient collective -- the time between them is communication the step could not hide
        self.comm_profile = [] if (self.world > 1 and os.environ.get("TG_COMM_PROFILE")) else None
        # TG_P2P=1: D's 6 MB gradient arena and the generator's small last buckets are summed by a one-shot kernel over
        # NVLink peer memory (p2p.PeerReducer) instead of NCCL. Off by default: measured on 8 B200s NCCL 2.28 reduces
        # the same 6.35 MB inside the NVSwitch (NVLS) in 65 us, the one-shot kernel reads 7 peers' copies in 106 us and
        # competes for SMs with the work the step overlaps it with (profiles/r02_p2p_vs_nccl.txt).
        self.peer = None
        if self.world > 1 and os.environ.get("TG_P2P", "0") == "1":
            from .p2p import PeerReducer, PeerUnavailable
            try:
                self.peer = PeerReducer(max(self.DA.store.grad_arena.numel(), self.PEER_MAX_BYTES // 4), dev,
                                        process_group)
            except PeerUnavailable as e:
                if torch.distributed.get_rank(process_group) == 0:
                    print(f"tactile_gan_b200: peer-memory all-reduce unavailable ({e}); using NCCL for every bucket")

    # ------------------------------------------------------------------ labels / alpha (host RNG parity)
    def ensure_label(self, generator=None):
        """GANLoss.get_target_tensor (generators/generators.py:52-63): one CPU-RNG draw, cached."""
        if self.real_label is None and self.label_smoothing and self.loss in ("ls", "ce"):
            shape = (self.B, 1, self.h5, self.w5)
            lab = torch.clamp(torch.normal(1.0, 0.02, size=shape, generator=generator), 0, 1)
            self.real_label = lab.to(self.device).contiguous()
        return self.real_label

    def set_label(self, label):
        if label is not None and tuple(label.shape) != (self.B, 1, self.h5, self.w5):
            raise ValueError(f"real-label tensor must be {(self.B, 1, self.h5, self.w5)}, got {tuple(label.shape)}")
        self.real_label = None if label is None else label.to(self.device).float().contiguous()
        self.reset_graphs()              # a captured step holds the old tensor's address

    def draw_alpha(self, alpha=None):
        """util.py:79-81: alpha ~ U[0,1) on the CUDA generator; version 2 maps it to [0.5, 1)."""
        if alpha is None:
            alpha = torch.rand(self.B, 1, device=self.device)
        elif not (alpha.is_cuda and alpha.dtype == torch.float32 and alpha.is_contiguous()):
            alpha = alpha.to(self.device).float().contiguous()
        self._alpha_src = alpha       # keep alive: the launch is asynchronous
        _C.call("gp_alpha", ptr(alpha), self.version, ptr(self.alpha), ptr(self.one_minus_alpha), self.B)

    def _gan_loss(self, inst, n0, n1, target_is_real, for_disc, scale, slot, write_dz=True):
        u5 = inst.u[4]
        lab = self.real_label if (target_is_real and self.loss in ("ls", "ce")) else None
        const = 1.0 if target_is_real else 0.0
        _C.call("gan_loss", ptr(inst.pred), ptr(lab), F(const), self.mode, int(target_is_real), int(for_disc),
                int(inst.has_sigmoid), F(scale), n0, n1, self.h5 * self.w5, u5.c, ptr(self.losses[slot:slot + 1]),
                ptr(u5.dz) if write_dz else None)

    def _allreduce_async(self, store):
        """Sum the whole gradient arena over the ranks on the communication stream; returns the work handle (None on
        one GPU). The compute stream keeps running until _allreduce_wait."""
        if self.world == 1:
            return None
        if not hasattr(self, "_comm_stream"):
            self._comm_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            return self._reduce_on_comm_stream(store.grad_arena)

    def _reduce_on_comm_stream(self, buf):
        """Called with the communication stream current. Returns something with .wait() (the compute stream waits)."""
        if self.peer is not None and buf.numel() * 4 <= self.PEER_MAX_BYTES and buf.numel() % 4 == 0:
            self.peer.allreduce_(buf)
            done = torch.cuda.Event()
            done.record()
            return _EventWork(done)
        return torch.distributed.all_reduce(buf, group=self.pg, async_op=True)

    def _allreduce_wait(self, work, tag):
        if work is None:
            return
        prof = self.comm_profile
        if prof is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        work.wait()
        if prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            prof.append((tag, e0, e1))

    # ---- generator gradients: bucketed allreduce overlapped with the rest of backward -----------------
    BUCKET_BYTES = 32 << 20
    TAIL_BUCKET_BYTES = 1 << 20      # the last-completing parameters travel alone: the only exposed collective
    PEER_MAX_BYTES = 8 << 20         # buffers up to this size take the one-shot peer-memory all-reduce

    def _plan_g_buckets(self):
        """The gradient arena is laid out in backward-completion order (ParamStore.finalize), so bucket k is a
        contiguous range that is final once the unit owning its last parameter has run."""
        from .layers import plan_buckets
        store = self.G.store
        buckets = plan_buckets(store.arena_layout, self.BUCKET_BYTES // 4, self.TAIL_BUCKET_BYTES // 4)
        owner = {}
        for u in self.G.units:
            for p in (getattr(u, "gamma", None), getattr(u, "beta", None), getattr(getattr(u, "layer", None), "weight", None),
                      getattr(getattr(u, "layer", None), "bias", None)):
                if p is not None and id(p) in store.index:
                    owner[store.index[id(p)]] = u
        self._g_buckets = [(a, b, owner.get(last)) for a, b, last in buckets]
        if not hasattr(self, "_comm_stream"):
            self._comm_stream = torch.cuda.Stream(device=self.device)

    def _g_backward(self):
        G, gs = self.G, self.G.store
        if self.world == 1:
            G.backward(self.gan_grad, self.l1_grad)
            return
        if not hasattr(self, "_g_buckets"):
            self._plan_g_buckets()
        pending = list(self._g_buckets)
        works = []

        def launch(a, b):
            ev = torch.cuda.Event()
            ev.record()          # on the stream that finalised the bucket (the generator's weight-gradient stream)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                works.append(self._reduce_on_comm_stream(gs.grad_arena[a:b]))

        def after_unit(u):
            while pending and pending[0][2] is u:
                a, b, _ = pending.pop(0)
                launch(a, b)

        G.backward(self.gan_grad, self.l1_grad, after_unit=after_unit)
        for a, b, _ in pending:          # buckets whose last parameter has no unit (head-only / dead parameters)
            launch(a, b)
        prof = self.comm_profile
        if prof is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()                  # backward's own kernels are all queued in front of this marker
        for w in works:
            w.wait()                     # the compute stream waits for the collectives; the host does not block
        if prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            prof.append(("G", e0, e1))

    # ------------------------------------------------------------------ the iteration
    def step(self, real_A, real_B, regularize=True, alpha=None, real_B_ready=None):
        """real_A (B,in,H,W) in [-1,1], real_B (B,out,H,W) in [0,1]: fp32, contiguous, on the device.
        Returns the device tensor of loss slots (see SLOT); reading it is the caller's only sync."""
        if self.use_graph and not (_C.TIMING["on"] or _C.TIMING["tail"]):
            return self._step_graphed(real_A, real_B, regularize, alpha, real_B_ready)
        return self._step_eager(real_A, real_B, regularize, alpha, real_B_ready)

    def _step_eager(self, real_A, real_B, regularize, alpha, real_B_ready):
        if self.pdl:
            prev = _C.lib().tg_pdl_policy(1)
            try:
                return self._step(real_A, real_B, regularize, alpha, real_B_ready)
            finally:
                _C.lib().tg_pdl_policy(prev)
        return self._step(real_A, real_B, regularize, alpha, real_B_ready)

    # ---- whole-step CUDA graph (small batches) ----------------------------------------------------------------
    def reset_graphs(self):
        """Drop the captured steps (after anything baked into them changed: label tensor, loss weights, ...)."""
        self._graphs.clear()
        self._eager_done.clear()

    def _write_hyper(self):
        """Advance both optimisers' step counts and put this iteration's Adam scalars where the captured launches read
        them (pinned host -> device on the current stream, ordered before the replay)."""
        from .layers import ParamStore
        gs, ds = self.G.store, self.DA.store
        ds.step_count += 1
        gs.step_count += 1
        import ctypes
        vals = ParamStore.adam_hyper(self.lr, self.beta1, ds.step_count, grad_scale=1.0 / self.world) + [0.0] + \
            ParamStore.adam_hyper(self.lr, self.beta1, gs.step_count, grad_scale=1.0 / self.world) + [0.0]
        # sixteen scalars as kernel arguments: no pinned staging buffer the host could overwrite before the copy ran
        _C.call("write_floats", ptr(self._hyper), (ctypes.c_float * 16)(*vals), 16)

    def _step_graphed(self, real_A, real_B, regularize, alpha, real_B_ready):
        reg = bool(regularize) and self.lambda_gp != 0
        if self._eager_done.get(reg, 0) < 2:          # lazy buffers, labels, cuBLAS-free but still: warm every path
            self._eager_done[reg] = self._eager_done.get(reg, 0) + 1
            return self._step_eager(real_A, real_B, regularize, alpha, real_B_ready)
        G, DA = self.G, self.DA
        cur = torch.cuda.current_stream()
        if self._hyper is None:
            self._hyper = torch.zeros(16, device=self.device)     # Adam scalars of D [0:7] and G [8:15]
            self._gA, self._gB = torch.empty_like(real_A), torch.empty_like(real_B)
        G.store.refresh()                              # weight re-packs after load_state_dict etc. stay outside the graph
        DA.store.refresh()
        self.ensure_label()
        self._gA.copy_(real_A, non_blocking=True)
        if real_B_ready is not None:
            cur.wait_event(real_B_ready)
        self._gB.copy_(real_B, non_blocking=True)
        if reg:
            self.draw_alpha(alpha)                     # static alpha / 1 - alpha buffers, written outside the graph
        entry = self._graphs.get(reg)
        if entry is None:
            entry = self._capture(reg)
            if entry is None:                          # capture failed: stay eager for good
                self.use_graph = False
                return self._step_eager(real_A, real_B, regularize, alpha, real_B_ready)
            self._graphs[reg] = entry
        graph, launches = entry
        self._write_hyper()
        graph.replay()
        _C.COUNTERS["launches"] += launches
        return self.losses

    def _capture(self, reg):
        launches0 = _C.COUNTERS["launches"]
        self._in_graph = True
        # programmatic graph edges wherever the eager step would launch programmatic dependents
        pdl_on = self.pdl if self.pdl is not None else _C.lib().tg_pdl_policy(-1) != 0
        prev_pdl = _C.lib().tg_pdl_policy(2 if pdl_on else 0)
        graph = torch.cuda.CUDAGraph()
        try:
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._step(self._gA, self._gB, reg, None, None)
        except Exception as e:  # pragma: no cover - depends on the driver / torch build
            print(f"tactile_gan_b200: CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); "
                  f"staying eager")
            return None
        finally:
            self._in_graph = False
            _C.lib().tg_pdl_policy(prev_pdl)
        return graph, _C.COUNTERS["launches"] - launches0

    def _step(self, real_A, real_B, regularize, alpha, real_B_ready):
        B, HW = self.B, self.H * self.W
        G, DA, S1, S2 = self.G, self.DA, self.S1, self.S2
        gs, ds = G.store, DA.store
        self.ensure_label()
        self.losses.zero_()
        regularize = bool(regularize) and self.lambda_gp != 0
        in_graph = getattr(self, "_in_graph", False)
        if regularize and not in_graph:
            self.draw_alpha(alpha)
        # ---- generator forward (train.py:104)
        fake = G.forward(real_A)
        self.fake_B = fake
        # ---- D step (train.py:107-135)
        if real_B_ready is not None:      # real_B is still arriving on a copy stream (step_from_host)
            torch.cuda.current_stream().wait_event(real_B_ready)
        ds.zero_grad()
        DA.pack_input(real_A, fake, n0=0, n=B)
        DA.pack_input(real_A, real_B, n0=B, n=B)
        DA.forward()
        DA.u[4].dz.zero_()
        self._gan_loss(DA, 0, B, False, True, 0.5, SLOT["loss_D"])
        self._gan_loss(DA, B, 2 * B, True, True, 0.5, SLOT["loss_D"])
        DA.backward(wgrad=True)
        if regularize:
            S1.pack_input(real_A, real_B, wa=self.alpha, b2=fake, wb=self.one_minus_alpha)
            S1.forward()
            S1.gp_first_backward()
            S1.gp_penalty(self.c_a, self.c_b, self.lambda_gp, 1.0, self.losses[1:2])
            S1.gp_second_backward()
        d_work = self._allreduce_async(ds)
        # ---- G step (train.py:138-168). What does not depend on the updated discriminator is queued while D's
        # gradients are in flight: the generator's gradient clear, the im2col rows of the (real_A, fake_B) pair (and of
        # the real pair for the feature term) and the L1 term
        gs.zero_grad()
        S1.pack_input(real_A, fake)
        _C.call("l1_loss", ptr(fake), ptr(real_B), LL(fake.numel()), F(self.lambda_a),
                ptr(self.losses[3:4]), ptr(self.l1_grad))
        want_feat = self.lambda_per != 0 and self.version == 2
        if want_feat:
            S2.pack_input(real_A, real_B)
        self._allreduce_wait(d_work, "D")
        if in_graph:
            ds.adam_step_dev(self._hyper[0:8])
        else:
            ds.adam_step(self.lr, self.beta1, grad_scale=1.0 / self.world)
        S1.forward()
        S1.u[4].dz.zero_()
        self._gan_loss(S1, 0, B, True, False, 1.0, SLOT["G_GAN"])
        S1.backward(wgrad=False, input_grad=True)
        S1.input_grad_image(self.c_a, self.c_b, self.gan_grad)
        if want_feat:
            S2.forward()
            wsum = sum(self.w_per)
            for fr, ff, w in zip(S2.features(), S1.features(), self.w_per):
                if w == 0:
                    continue
                cpad = fr.buf.shape[3]
                _C.call("feat_loss", ptr(fr.buf), ptr(ff.buf), LL(fr.buf.numel()),
                        F(self.lambda_per * w / wsum * cpad / fr.c), 0, ptr(self.losses[4:5]))
        if self.VF is not None:
            # version 1: per = lambda_per * sum_i w_i * L1mean(VGG_i(real_B), VGG_i(fake_B)) (util.py:119-144);
            # its gradient w.r.t. fake_B is added to the L1 gradient buffer
            self.VR.forward(real_B)
            self.VF.forward(fake)
            self.VF.loss_and_seed(self.VR, self.w_per, self.lambda_per, self.losses[4:5])
            self.VF.backward(self.l1_grad)
        self._g_backward()
        if in_graph:
            gs.adam_step_dev(self._hyper[8:16])
        else:
            gs.adam_step(self.lr, self.beta1, grad_scale=1.0 / self.world)
        return self.losses

    def step_from_host(self, host_A, host_B, regularize=True, prefetch=None, lag=False, read=True, alpha=None):
        """End-to-end iteration as a training loop issues it (reference train.py:101-168): pinned host
        batch -> H2D copy -> fused step -> D2H read of the five loss scalars (the only sync).
        prefetch: the NEXT iteration's pinned (host_A, host_B). Its H2D copy is issued now, on the copy stream, into
        the second pair of device buffers and runs under this step's kernels -- the one-batch lookahead of a
        DataLoader; the next call recognises the batch (same tensors) and only waits for the copy's event.
        lag: read the loss scalars back with a one-iteration lag -- this call enqueues a pinned D2H copy of its own
        losses and returns the PREVIOUS iteration's (None on the first call), so the host can queue the next iteration
        while this one runs instead of draining the device every step (train.py itself syncs once per epoch). A call
        with lag=False returns its own losses synchronously, as before. read=False: no D2H at all, the device tensor of
        loss slots is returned (train.py accumulates it on the device and syncs once per epoch)."""
        if not hasattr(self, "_dev"):
            self._dev = [(torch.empty(host_A.shape, device=self.device), torch.empty(host_B.shape, device=self.device))
                         for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._slot_free = [None, None]       # event: the last step that read the slot has finished
            self._inflight = None                # (host_A, host_B, slot, ready_A, ready_B)
            self._slot = 0
        cur = torch.cuda.current_stream()
        cs = self._copy_stream
        fl = self._inflight
        self._inflight = None
        if fl is not None and fl[0] is host_A and fl[1] is host_B:
            slot, ready_a, ready_b = fl[2], fl[3], fl[4]
            cur.wait_event(ready_a)
        else:
            # no lookahead: the generator forward only needs real_A, the target's copy runs beside it
            slot = self._slot
            if fl is not None:                   # a lookahead copy of some other batch may still be writing this slot
                cur.wait_event(fl[4])
            if self._slot_free[slot] is not None:
                cs.wait_event(self._slot_free[slot])
            self._dev[slot][0].copy_(host_A, non_blocking=True)
            ready_b = torch.cuda.Event()
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                self._dev[slot][1].copy_(host_B, non_blocking=True)
                ready_b.record()
        if prefetch is not None:
            nslot = slot ^ 1
            ra, rb = torch.cuda.Event(), torch.cuda.Event()
            if self._slot_free[nslot] is not None:
                cs.wait_event(self._slot_free[nslot])
            with torch.cuda.stream(cs):
                self._dev[nslot][0].copy_(prefetch[0], non_blocking=True)
                ra.record()
                self._dev[nslot][1].copy_(prefetch[1], non_blocking=True)
                rb.record()
            self._inflight = (prefetch[0], prefetch[1], nslot, ra, rb)
        self.step(self._dev[slot][0], self._dev[slot][1], regularize=regularize, alpha=alpha, real_B_ready=ready_b)
        done = torch.cuda.Event()
        done.record()
        self._slot_free[slot] = done
        self._slot = slot ^ 1
        if not read:
            return self.losses
        if not lag:
            self._loss_pending = None
            return self.loss_dict()
        if not hasattr(self, "_loss_host"):
            self._loss_host = [torch.empty(self.losses.shape, dtype=self.losses.dtype).pin_memory() for _ in range(2)]
            self._loss_k = 0
        prev = getattr(self, "_loss_pending", None)
        self._loss_k ^= 1
        buf = self._loss_host[self._loss_k]
        buf.copy_(self.losses, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._loss_pending = (buf, ev)
        if prev is None:
            return None
        prev[1].synchronize()
        v = prev[0].tolist()
        return {"loss_D": v[0], "gp": v[1], "G_GAN": v[2], "L1": v[3], "per": v[4]}

    def loss_dict(self):
        """Host copy of the loss slots with the reference's logging conventions (train.py:121-163):
        loss_D excludes the penalty, L1 is logged unscaled."""
        v = self.losses.tolist()
        return {"loss_D": v[0], "gp": v[1], "G_GAN": v[2], "L1": v[3], "per": v[4]}
