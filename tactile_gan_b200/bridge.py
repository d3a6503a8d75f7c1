"""nn.Module <-> engine bridge: the drop-in modules keep the reference's call conventions
(`netG(x)`, `loss.backward()`, `optimizer.step()`) while every device computation runs in the
sm_100a engines. There is no eager fallback: a CPU tensor or a missing library raises."""
import torch
import torch.nn as nn

from . import _C

INFER_CHUNK_PIXELS = 64 * 256 * 256


class _EngineFn(torch.autograd.Function):
    """Whole-network autograd node: forward = engine forward; backward = engine backward, returning the
    parameter gradients in torch layout."""

    @staticmethod
    def forward(ctx, x, eng, *params):
        out = eng.forward(x.detach().contiguous().float())
        ctx.eng = eng
        ctx.nparams = len(params)
        return out.clone()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.eng
        eng.store.zero_grad()
        eng.backward(g.contiguous().float())
        grads = [eng.store.grad_as_torch(i) for i in range(ctx.nparams)]
        return (None, None, *grads)


class EngineModule(nn.Module):
    """Base of the generators: caches one engine per input shape."""
    engine_kind = None

    def _engine(self, n, h, w, backward):
        from .engine import build_generator_engine
        cache = self.__dict__.setdefault("_tg_engines", {})
        key = (n, h, w, backward)
        if key not in cache:
            if (n, h, w, True) in cache:      # a training engine also serves inference
                return cache[(n, h, w, True)]
            cache[key] = build_generator_engine(self.engine_kind, self, n, h, w, backward)
        return cache[key]

    @staticmethod
    def infer_chunk(h, w):
        """Images per inference engine: activations are ~0.4 GB per 256x256 image (UNet++), so a batch-512 sweep
        (BASELINE.json configs[4]) runs as chunks of 64 images' worth of pixels; throughput is flat beyond that."""
        return max(1, INFER_CHUNK_PIXELS // (h * w))

    def forward(self, x):
        if not x.is_cuda:
            raise _C.TgError("tactile_gan_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
        n, _, h, w = x.shape
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if need_grad:
            return _EngineFn.apply(x, self._engine(n, h, w, True), *params)
        chunk = self.infer_chunk(h, w)
        xs = x.detach().contiguous().float()
        if n <= chunk:
            return self._engine(n, h, w, False).forward_graphed(xs).clone()
        out = None
        for i in range(0, n, chunk):
            m = min(chunk, n - i)
            y = self._engine(m, h, w, False).forward_graphed(xs[i:i + m])
            if out is None:
                out = torch.empty(n, *y.shape[1:], device=y.device)
            out[i:i + m].copy_(y)
        return out


# ------------------------------------------------------------------------------- discriminator bridge
class _DiscFn(torch.autograd.Function):
    """Whole-discriminator autograd node (first order). Gradients flow to the parameters and to img_B
    (the generator output); the gradient-penalty double backward has its own fused entry point
    (tactile_gan_b200.util.gradient_penalty)."""

    @staticmethod
    def forward(ctx, img_a, img_b, inst, *params):
        inst.pack_input(img_a.detach().contiguous().float(), img_b.detach().contiguous().float())
        pred = inst.forward()
        ctx.inst, ctx.nparams = inst, len(params)
        ctx.need_b = img_b.requires_grad
        ctx.ca, ctx.cb = img_a.shape[1], img_b.shape[1]
        u5 = inst.u[4]
        out = torch.empty(inst.n, 1, u5.ho, u5.wo, device=pred.device)
        _C.call("unpack_nhwc", _C.ptr(pred), _C.ptr(out), inst.n, u5.ho * u5.wo, u5.c, 0, 1, _C.F(1.0))
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        inst = ctx.inst
        (out,) = ctx.saved_tensors
        u5 = inst.u[4]
        gz = g.contiguous().float()
        if inst.has_sigmoid:
            gz = gz * out * (1 - out)
        u5.dz.zero_()
        _C.call("pack_nchw", _C.ptr(gz), None, None, None, _C.ptr(u5.dz), inst.n, u5.ho * u5.wo, 1, u5.c, 0)
        want_w = any(ctx.needs_input_grad[3:])
        if want_w:
            inst.store.zero_grad()
        inst.backward(wgrad=want_w, input_grad=ctx.need_b)
        gb = None
        if ctx.need_b:
            gb = torch.empty(inst.n, ctx.cb, inst.h, inst.w, device=g.device)
            inst.input_grad_image(ctx.ca, ctx.cb, gb)
        grads = [inst.store.grad_as_torch(i) if want_w else None for i in range(ctx.nparams)]
        return (None, gb, None, *grads)


def disc_instance(module, n, h, w, backward=True, second_order=False, slot=0):
    from .engine import PatchDInstance
    cache = module.__dict__.setdefault("_tg_engines", {})
    key = (n, h, w, backward, second_order, slot)
    if key not in cache:
        cache[key] = PatchDInstance(module, n, h, w, backward, second_order)
    return cache[key]


def disc_forward(module, img_a, img_b):
    if not img_a.is_cuda:
        raise _C.TgError("tactile_gan_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
    n, _, h, w = img_a.shape
    params = list(module.parameters())
    need = torch.is_grad_enabled() and (img_b.requires_grad or any(p.requires_grad for p in params))
    slot = module.__dict__.get("_tg_slot", 0)
    module.__dict__["_tg_slot"] = (slot + 1) % 4   # consecutive forwards keep separate activations
    inst = disc_instance(module, n, h, w, True, False, slot)
    if need:
        pred = _DiscFn.apply(img_a, img_b, inst, *params)
    else:
        inst.pack_input(img_a.detach().contiguous().float(), img_b.detach().contiguous().float())
        p = inst.forward()
        u5 = inst.u[4]
        pred = torch.empty(n, 1, u5.ho, u5.wo, device=p.device)
        _C.call("unpack_nhwc", _C.ptr(p), _C.ptr(pred), n, u5.ho * u5.wo, u5.c, 0, 1, _C.F(1.0))
    feats = []
    if module.return_filters:
        for a in inst.features():
            nn_, hh, ww, cc = a.buf.shape
            f = torch.empty(nn_, a.c, hh, ww, device=a.buf.device)
            _C.call("unpack_nhwc", _C.ptr(a.buf), _C.ptr(f), nn_, hh * ww, cc, 0, a.c, _C.F(1.0))
            feats.append(f)
    return pred, feats
