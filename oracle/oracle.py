"""ORACLE -- test infrastructure only (never imported by the product path).

A CPU, fp32, plain-PyTorch *functional* restatement of the tactile-gan G+D training step. It works on
flat ``state_dict``s (name -> tensor) whose keys are exactly the reference's parameter names, so the
reference modules and this file can be driven with the same weights.

What follows which reference lines (all under /root/reference):
  unetpp_forward   generators/UNet_plusplus.py:18-34 (ConvBlock), :65-86 (forward), :5-16 (head)
  unet_forward     generators/UNet.py:17-51 (ConvDown/DeconvUp), :80-99 (forward)
  bcdunet_forward  generators/BCDUNet.py:120-143 (blocks), :154-181 (forward; ConvLSTM never called)
  patchd_forward   discriminators/PatchDiscriminator.py:12-37 (+ the 4 LeakyReLU feature taps :39-43)
  gan_loss         generators/generators.py:80-105
  pan_loss         util.py:41-70 (mode='normal' only, as called from train.py:160)
  gradient_penalty util.py:72-97
  vgg_perceptual   util.py:100-144 (VGGPerceptualLoss: 4 slices of torchvision VGG16.features[:23]; version 1 only)
  eval_pair_fuzzy  test.py:113-124 (eval_pair, fuzzy=True: the accuracy / Dice / Jaccard test_model reports)
  augment_pair     datasets/PairedDataset.py:30-44,80-92 (flip + affine + ToTensor/Normalize). PARITY UNPINNED: the
                   arithmetic belongs to albumentations (unpinned, not installed); this restates the transform
                   tactile_gan_b200/augment.py specifies (inverse map in 16.16 fixed point), not albumentations' own.
  convlstm_cell / convlstm / convblstm   generators/BCDUNet.py:32-47 (cell), :61-84 (unrolled sequence, zero initial
                   state), :96-103 (bidirectional: second cell on the reversed frames, channel concat). Constructed by
                   BCDUNet but never called by its forward; pinned by tests/golden/convlstm.pt (reference classes run here)
  adam_update      torch.optim.Adam as configured at train.py:56-57 (betas=(beta1,0.99), eps=1e-8)
  train_step       train.py:99-168

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference's own modules executed here -- see oracle/make_golden.py, which
imports /root/reference, and tests/test_oracle_golden.py which replays the committed fixtures.
Arithmetic the reference delegates to third parties: PyTorch (torch 2.11.0 here; conv / instance_norm /
autograd / Adam) and, for the version-1 perceptual term only, torchvision VGG16 weights (not
obtainable offline: version 1 is compared with shared random-init weights).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

EPS_IN = 1e-5

# ---- optional bf16 storage emulation ---------------------------------------------------------------
# The CUDA path stores activations / packed weights as bf16 (fp32 accumulate). With QUANT["on"] the oracle
# rounds at exactly those storage points (straight-through in backward), which isolates logic errors
# from the legitimate bf16-vs-fp32 drift (ReLU masks flip when activations move by ~1%).
QUANT = {"on": False}


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


ACT = {"relu": F.relu}   # tests may linearise the generators (ACT["relu"] = identity) to remove mask flips


def max_pool_2x2(t):
    return F.max_pool2d(t, 2, 2)


# tests may pin the arg-max routes (and ACT["relu"] the masks) to the pattern the CUDA forward took: tests/parity_util.py
POOL = {"max": max_pool_2x2}


def _q(x):
    return _RoundBF16.apply(x) if QUANT["on"] else x



# --------------------------------------------------------------------------- generators
def _in(x, sd, key):
    """InstanceNorm2d, biased variance, eps 1e-5, optional affine (key.weight / key.bias)."""
    return F.instance_norm(x, weight=sd.get(key + ".weight"), bias=sd.get(key + ".bias"), eps=EPS_IN)


def _double_conv(x, sd, p, first=None):
    """[conv -> IN -> ReLU] x 2 with parameter names p.0 / p.1 / p.3 / p.4.
    `first` overrides the first op (strided conv / transposed conv for UNet)."""
    if first is None:
        x = F.conv2d(x, _q(sd[p + ".0.weight"]), sd.get(p + ".0.bias"), stride=1, padding=1)
    else:
        x = first(x)
    x = _q(ACT["relu"](_in(_q(x), sd, p + ".1")))
    x = F.conv2d(x, _q(sd[p + ".3.weight"]), sd.get(p + ".3.bias"), stride=1, padding=1)
    return _q(ACT["relu"](_in(_q(x), sd, p + ".4")))


def _head(x, sd, key, activation):
    x = F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"])
    return torch.tanh(x) if activation else x


def unetpp_forward(sd, x, activation=True):
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    down = lambda t: _q(F.avg_pool2d(t, 2, 2))
    x = _q(x)
    X = {}
    blk = lambda i, j, inp: _double_conv(inp, sd, f"conv{i}_{j}.layer")
    X[0, 0] = blk(0, 0, x)
    for i in range(1, 5):
        X[i, 0] = blk(i, 0, down(X[i - 1, 0]))
    for j in range(1, 5):
        for i in range(0, 5 - j):
            X[i, j] = blk(i, j, torch.cat([X[i, k] for k in range(j)] + [up(X[i + 1, j - 1])], 1))
    return _head(X[0, 4], sd, "downfeature.conv", activation)


def unet_forward(sd, x, activation=True):
    c = [None]
    t = _q(x)
    for i in range(1, 8):
        w = sd[f"conv{i}.layer.0.weight"]
        t = _double_conv(t, sd, f"conv{i}.layer", first=lambda z, w=w: F.conv2d(z, _q(w), None, stride=2, padding=1))
        c.append(t)
    d = c[7]
    for i in range(2, 9):
        w = sd[f"deconv{i}.layer.0.weight"]
        inp = d if i == 2 else torch.cat([d, c[9 - i]], 1)  # d3 <- (d2, c6) ... d8 <- (d7, c1)
        d = _double_conv(inp, sd, f"deconv{i}.layer",
                         first=lambda z, w=w: F.conv_transpose2d(z, _q(w), None, stride=2, padding=1))
    return _head(d, sd, "downfeature.conv", activation)


def bcdunet_forward(sd, x, activation=True):
    blk = lambda name, inp: _double_conv(inp, sd, name)
    pool = lambda t: POOL["max"](t)
    c1 = blk("conv1", _q(x))
    c2 = blk("conv2", pool(c1))
    c3 = blk("conv3", pool(c2))
    c4 = blk("conv4", pool(c3))
    u3 = _q(F.conv_transpose2d(c4, _q(sd["upconv3.weight"]), sd["upconv3.bias"], stride=2))
    m3 = blk("conv3m", torch.cat([c3, u3], 1))
    u2 = _q(F.conv_transpose2d(m3, _q(sd["upconv2.weight"]), sd["upconv2.bias"], stride=2))
    m2 = blk("conv2m", torch.cat([c2, u2], 1))
    u1 = _q(F.conv_transpose2d(m2, _q(sd["upconv1.weight"]), sd["upconv1.bias"], stride=2))
    m1 = blk("conv1m", torch.cat([c1, u1], 1))
    return _head(m1, sd, "conv0", activation)


GEN_FORWARD = {"unet++": unetpp_forward, "unet": unet_forward, "bcdunet": bcdunet_forward}


def gen_forward(name, sd, x, activation=True):
    return GEN_FORWARD[name.lower()](sd, x, activation)


# --------------------------------------------------------------------------- discriminator
def patchd_forward(sd, img_a, img_b, activation=True):
    """Returns (prediction, [4 LeakyReLU feature maps])."""
    feats = []
    x = _q(torch.cat([img_a, img_b], 1))
    x = _q(F.leaky_relu(F.conv2d(x, _q(sd["model.0.weight"]), sd["model.0.bias"], stride=2), 0.2))
    feats.append(x)
    for conv, norm, stride in ((2, 3, 2), (5, 6, 1), (8, 9, 1)):
        x = _q(F.conv2d(x, _q(sd[f"model.{conv}.weight"]), None, stride=stride))
        x = _q(F.leaky_relu(_in(x, sd, f"model.{norm}"), 0.2))
        feats.append(x)
    x = F.conv2d(x, _q(sd["model.11.weight"]), sd["model.11.bias"])
    if activation:
        x = torch.sigmoid(x)
    return _q(x), feats


# --------------------------------------------------------------------------- losses
def make_real_label(shape, smoothing=True, real_label=1.0, generator=None):
    """GANLoss caches this on first use (generators.py:54-61): one CPU-RNG draw per run."""
    if smoothing:
        return torch.clamp(torch.normal(real_label, 0.02, size=shape, generator=generator), 0, 1)
    return torch.full((1,), real_label)


def gan_loss(pred, target_is_real, mode, real_label, for_discriminator=True, fake_label=0.0):
    if mode in ("ls", "ce"):
        tgt = real_label.to(pred).expand_as(pred) if target_is_real else torch.full_like(pred, fake_label)
        if mode == "ls":
            return F.mse_loss(pred, tgt)
        return F.binary_cross_entropy_with_logits(pred, tgt)
    if mode == "hinge":
        if for_discriminator:
            m = (pred - 1) if target_is_real else (-pred - 1)
            return -torch.mean(torch.minimum(m, torch.zeros_like(m)))
        return -torch.mean(pred)
    if mode == "w":
        return -pred.mean() if target_is_real else pred.mean()
    raise ValueError(f"Unexpected gan mode {mode}")


def pan_loss(real_feats, fake_feats, weights):
    w = [float(v) / float(sum(weights)) for v in weights]
    total = 0.0
    for i in range(4):
        total = total + F.l1_loss(real_feats[i], fake_feats[i]) * w[i]
    return total


def gradient_penalty(sd_d, real_a, real_b, fake_b, alpha, activation, lambda_gp, version=2, constant=1.0):
    """alpha: (B,1) uniform draw (the reference draws it on the CUDA generator, util.py:79)."""
    a = (alpha + 1) / 2 if version == 2 else alpha
    a = a.view(-1, 1, 1, 1)
    inter = (a * real_b + (1 - a) * fake_b).detach().requires_grad_(True)
    pred, _ = patchd_forward(sd_d, real_a, inter, activation)
    (g,) = torch.autograd.grad(pred, inter, torch.ones_like(pred), create_graph=True, retain_graph=True)
    g = g.reshape(g.shape[0], -1)
    return (((g + 1e-16).norm(2, dim=1) - constant) ** 2).mean() * lambda_gp


# --------------------------------------------------------------------------- version-1 perceptual term
# torchvision VGG16.features[:23] in the reference's four slices (util.py:104-107). nn.Sequential slicing keeps the
# original child names, so the ModuleList's state_dict keys are blocks.<slice>.<features index>.{weight,bias}.
VGG_SLICES = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21))      # conv indices; slices 1..3 start with MaxPool2d(2)
VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)


def vgg_features(sd, x):
    """x: normalised (and resized) (B,3,h,w) -> the four slice outputs (util.py:133-135)."""
    feats = []
    for b, convs in enumerate(VGG_SLICES):
        if b > 0:
            x = F.max_pool2d(x, 2)
        for i in convs:
            x = _q(F.relu(F.conv2d(x, _q(sd[f"blocks.{b}.{i}.weight"]), sd[f"blocks.{b}.{i}.bias"], padding=1)))
        feats.append(x)
    return feats


def vgg_transform(x, resize=True):
    """util.py:120-129: repeat to 3 channels, ImageNet mean / std, bilinear resize to 224 (align_corners=False)."""
    if x.shape[1] != 3:
        x = x.repeat(1, 3, 1, 1)
    mean = torch.tensor(VGG_MEAN, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor(VGG_STD, dtype=x.dtype).view(1, 3, 1, 1)
    x = (x - mean) / std
    if resize:
        x = F.interpolate(x, mode="bilinear", size=(224, 224), align_corners=False)
    return x


def vgg_perceptual(sd, inp, target, weights=(0.25, 0.25, 0.25, 0.25), feature_layers=(0, 1, 2, 3), resize=True):
    """VGGPerceptualLoss.forward with style_layers=[] (util.py:119-144): sum_i weights[i] * L1mean(x_i, y_i);
    unlike pan_loss the weights are NOT normalised."""
    fx = vgg_features(sd, _q(vgg_transform(inp, resize)))
    fy = vgg_features(sd, _q(vgg_transform(target, resize)))
    loss = 0.0
    for i in range(4):
        if i in feature_layers:
            loss = loss + F.l1_loss(fx[i], fy[i]) * weights[i]
    return loss


# --------------------------------------------------------------------------- input pipeline (PairedDataset.py:30-92)
def augment_pair(img_u8, mask_u8, params):
    """numpy restatement of tg_augment_pair: img (N,H,W,ca) / mask (N,H,W,cb) uint8, params (N,8) int64
    {flip, a00..a12 in 16.16 fixed point (inverse map)} -> (fp32 NCHW in [-1,1], fp32 NCHW in [0,1])."""
    import numpy as np
    img, mask, q = img_u8.numpy(), mask_u8.numpy(), params.numpy().astype(np.int64)
    n, h, w, ca = img.shape
    cb = mask.shape[3]
    ys, xs = np.meshgrid(np.arange(h, dtype=np.int64), np.arange(w, dtype=np.int64), indexing="ij")
    out_a = np.zeros((n, ca, h, w), np.float32)
    out_b = np.zeros((n, cb, h, w), np.float32)
    for i in range(n):
        sx = q[i, 1] * xs + q[i, 2] * ys + q[i, 3]
        sy = q[i, 4] * xs + q[i, 5] * ys + q[i, 6]
        if q[i, 0]:
            sx = ((w - 1) << 16) - sx
        x0, y0 = sx >> 16, sy >> 16
        fx = ((sx & 0xFFFF).astype(np.float32) * np.float32(1 / 65536))[..., None]
        fy = ((sy & 0xFFFF).astype(np.float32) * np.float32(1 / 65536))[..., None]

        def tap(yy, xx, src):
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.float32)
            return np.where(ok[..., None], v, np.float32(0))
        v00, v01 = tap(y0, x0, img[i]), tap(y0, x0 + 1, img[i])
        v10, v11 = tap(y0 + 1, x0, img[i]), tap(y0 + 1, x0 + 1, img[i])
        top, bot = v00 + fx * (v01 - v00), v10 + fx * (v11 - v10)
        val = (top + fy * (bot - top)) / np.float32(255)
        out_a[i] = ((val - np.float32(0.5)) / np.float32(0.5)).transpose(2, 0, 1)
        xn, yn = (sx + 32768) >> 16, (sy + 32768) >> 16
        out_b[i] = (tap(yn, xn, mask[i]) / np.float32(255)).transpose(2, 0, 1)
    return torch.from_numpy(out_a), torch.from_numpy(out_b)


# --------------------------------------------------------------------------- evaluation (test.py:113-124)
def eval_pair_fuzzy(real, out):
    """eval_pair(real, out, fuzzy=True) for one (C,H,W) pair, in float64 numpy like the reference."""
    o = out.detach().cpu().numpy().astype("float64")
    r = real.detach().cpu().numpy().astype("float64")
    inter = (o * r).sum()
    denom = (o ** 2 + r ** 2).sum()
    import numpy as np
    return {"accuracy": float(np.minimum(o, r).sum() / r.sum()), "dice": float(2 * inter / denom),
            "jaccard": float(inter / (denom - inter))}


# --------------------------------------------------------------------------- optimiser
def convlstm_cell(sd, x, h_prev, c_prev, activation="tanh", prefix=""):
    """BCDUNet.py:32-47. sd keys: conv.weight [4C, Cin+C, k, k], conv.bias, W_ci / W_cf / W_co [C, H, W]."""
    act = torch.tanh if activation == "tanh" else torch.relu
    w = sd[prefix + "conv.weight"]
    z = F.conv2d(torch.cat([_q(x), _q(h_prev)], 1), _q(w), sd[prefix + "conv.bias"], padding=w.shape[2] // 2)
    zi, zf, zg, zo = torch.chunk(_q(z), 4, dim=1)
    i = torch.sigmoid(zi + sd[prefix + "W_ci"] * c_prev)
    f = torch.sigmoid(zf + sd[prefix + "W_cf"] * c_prev)
    c = f * c_prev + i * act(zg)
    o = torch.sigmoid(zo + sd[prefix + "W_co"] * c)
    return o * act(c), c


def convlstm(sd, x, activation="tanh", prefix="convLSTMcell.", return_sequence=True):
    """BCDUNet.py:61-84: x (B,T,Cin,H,W) -> (B,T,C,H,W) (or the last frame)."""
    b, t, _, hh, ww = x.shape
    c_out = sd[prefix + "W_ci"].shape[0]
    h = x.new_zeros(b, c_out, hh, ww)
    c = x.new_zeros(b, c_out, hh, ww)
    outs = []
    for k in range(t):
        h, c = convlstm_cell(sd, x[:, k], h, c, activation, prefix)
        outs.append(h)
    out = torch.stack(outs, 1)
    return out if return_sequence else out[:, -1]


def convblstm(sd, x, activation="tanh", return_sequence=True):
    """BCDUNet.py:96-103."""
    fwd = convlstm(sd, x, activation, "forward_cell.convLSTMcell.")
    bwd = convlstm(sd, x.flip(1), activation, "backward_cell.convLSTMcell.").flip(1)
    out = torch.cat((fwd, bwd), dim=2)
    return out if return_sequence else out[:, -1]


def adam_update(params, grads, state, lr, beta1, beta2=0.99, eps=1e-8):
    """In-place Adam on dicts keyed by parameter name; state[name] = dict(step, exp_avg, exp_avg_sq)."""
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        st = state.setdefault(k, dict(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)))
        st["step"] += 1
        st["exp_avg"].mul_(beta1).add_(g, alpha=1 - beta1)
        st["exp_avg_sq"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** st["step"]
        bc2 = 1 - beta2 ** st["step"]
        denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(eps)
        p.data.addcdiv_(st["exp_avg"], denom, value=-lr / bc1)


# --------------------------------------------------------------------------- the step
class StepConfig:
    def __init__(self, gen="UNet++", loss="ls", version=2, lambda_a=1.0, lambda_gp=0.01, lambda_per=1.0,
                 w_per=(0, .1, .3, .6), lr=1e-3, beta1=0.9, regularize=True):
        self.gen, self.loss, self.version = gen, loss, version
        self.lambda_a, self.lambda_gp, self.lambda_per = lambda_a, lambda_gp, lambda_per
        self.w_per, self.lr, self.beta1, self.regularize = tuple(w_per), lr, beta1, regularize
        self.activation = loss == "ls"  # train.py:33
        self.vgg_sd = None              # version 1: state_dict of the four VGG16 slices (blocks.<b>.<i>.weight/bias)
        # tests only: discriminator weights to run the G step with INSTEAD of the oracle's own post-Adam weights. The
        # first Adam step moves every weight by ~lr*sign(grad); gradients near zero flip sign under bf16 noise, so the
        # two sides' updated discriminators differ by 2*lr on those weights. Handing the oracle the other side's
        # updated D isolates the G-step comparison from that (the D step is compared on its own gradients).
        self.sd_d_after = None


def _leaf(sd):
    return OrderedDict((k, v.detach().clone().requires_grad_(v.is_floating_point())) for k, v in sd.items())


def train_step(sd_g, sd_d, opt_g, opt_d, real_a, real_b, real_label, alpha, cfg):
    """One G+D iteration (train.py:99-168). Mutates sd_g / sd_d / opt_g / opt_d in place and returns
    dict(loss_D (before GP), gp, G_GAN, L1, per, fake_B, grads_D, grads_G)."""
    act = cfg.activation
    pg = _leaf(sd_g)
    pd = _leaf(sd_d)
    fake_b = gen_forward(cfg.gen, pg, real_a, act)

    # ---- D step (train.py:107-135)
    pred_fake, _ = patchd_forward(pd, real_a, fake_b.detach(), act)
    pred_real, _ = patchd_forward(pd, real_a, real_b, act)
    loss_d = (gan_loss(pred_fake, False, cfg.loss, real_label) + gan_loss(pred_real, True, cfg.loss, real_label)) / 2
    out = {"loss_D": float(loss_d.detach())}
    total_d = loss_d
    if cfg.regularize and cfg.lambda_gp != 0:
        gp = gradient_penalty(pd, real_a, real_b, fake_b.detach(), alpha, act, cfg.lambda_gp, cfg.version)
        total_d = total_d + gp
        out["gp"] = float(gp.detach())
    else:
        out["gp"] = 0.0
    names_d = [k for k, v in pd.items() if v.requires_grad]
    gd = torch.autograd.grad(total_d, [pd[k] for k in names_d], allow_unused=True)
    grads_d = {k: g for k, g in zip(names_d, gd) if g is not None}
    adam_update(sd_d, grads_d, opt_d, cfg.lr, cfg.beta1)

    # ---- G step (train.py:138-168), D already updated
    pd2 = OrderedDict((k, v.detach()) for k, v in (cfg.sd_d_after if cfg.sd_d_after is not None else sd_d).items())
    pred_fake, feats_fake = patchd_forward(pd2, real_a, fake_b, act)
    g_gan = gan_loss(pred_fake, True, cfg.loss, real_label, for_discriminator=False)
    l1 = F.l1_loss(real_b, fake_b)
    total_g = g_gan + l1 * cfg.lambda_a
    out["G_GAN"], out["L1"] = float(g_gan.detach()), float(l1.detach())
    if cfg.lambda_per != 0 and cfg.version != 2:
        # train.py:151-153,160: perceptual_loss(real_B, fake_B, weights=w_per) with frozen VGG16 weights
        per = vgg_perceptual(cfg.vgg_sd, real_b, fake_b, cfg.w_per) * cfg.lambda_per
        total_g = total_g + per
        out["per"] = float(per.detach())
    elif cfg.lambda_per != 0:
        _, feats_real = patchd_forward(pd2, real_a, real_b, act)
        # the reference stores detached clones of both feature lists: the term carries no gradient
        per = pan_loss([f.detach() for f in feats_real], [f.detach() for f in feats_fake], cfg.w_per) * cfg.lambda_per
        total_g = total_g + per
        out["per"] = float(per.detach())
    else:
        out["per"] = 0.0
    names_g = [k for k, v in pg.items() if v.requires_grad]
    gg = torch.autograd.grad(total_g, [pg[k] for k in names_g], allow_unused=True)
    grads_g = {k: g for k, g in zip(names_g, gg) if g is not None}
    adam_update(sd_g, grads_g, opt_g, cfg.lr, cfg.beta1)
    out.update(fake_B=fake_b.detach(), grads_D=grads_d, grads_G=grads_g)
    return out


def synthetic_batch(generator, batch, size, in_nc=3, out_nc=3):
    """Synthetic pair in the ranges the dataset produces: source in [-1,1] (Normalize(.5,.5),
    datasets/PairedDataset.py:52-58), target in [0,1] (ToTensor, :86)."""
    real_a = torch.rand(batch, in_nc, size, size, generator=generator) * 2 - 1
    real_b = torch.rand(batch, out_nc, size, size, generator=generator)
    return real_a, real_b


# --------------------------------------------------------------------------- init (util.py:23-34)
def init_state_dict(shapes, generator, gain=0.02):
    """N(0, gain) for conv weights, zeros for conv biases, (1, 0) for InstanceNorm affine.
    `shapes`: OrderedDict name -> shape with the reference's parameter names."""
    sd = OrderedDict()
    for k, shp in shapes.items():
        if len(shp) == 4:
            sd[k] = torch.randn(shp, generator=generator) * gain
        elif k.endswith(".weight"):
            sd[k] = torch.ones(shp)
        else:
            sd[k] = torch.zeros(shp)
    return sd
