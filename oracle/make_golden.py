"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported from
/root/reference) through a device-agnostic restatement of the train.py:99-168 loop body that uses the
reference's own classes: create_gen / create_disc / GANLoss / pan_loss / gradient_penalty / init_weights
and torch.optim.Adam. Run in the build container (the reference is not on the GPU box):

    python oracle/make_golden.py

train.py itself cannot be imported (needs albumentations, hard-codes cuda:0, reads a global `opt`), and
GANLoss hard-codes device='cuda' for its fake/zero label tensors (generators/generators.py:59,67,75), so
those two cached tensors are pre-seeded on the CPU -- the cache-on-None pattern makes that a no-code-
change workaround. The GP `alpha` is drawn through a patched torch.rand so the fixture records it.
"""
import os
import sys
from collections import OrderedDict

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import synthetic_batch  # noqa: E402

REF = os.environ.get("TACTILE_GAN_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def reference_modules():
    sys.path.insert(0, REF)
    import util as ref_util  # noqa: seeds torch/np/random with 21 at import (util.py:8-11)
    from generators.generators import create_gen, GANLoss
    from discriminators.discriminators import create_disc
    sys.path.pop(0)
    return ref_util, create_gen, GANLoss, create_disc


def run_case(gen_name, nf, size, batch, loss, steps, seed, lambda_gp=0.01, compact=False):
    ref_util, create_gen, GANLoss, create_disc = reference_modules()
    torch.manual_seed(seed)
    activation = loss == "ls"  # train.py:33
    netG = create_gen(gen_name, 3, 3, nf, activation)
    ref_util.init_weights(netG)
    netD = create_disc("patch", 3, 3, nf, return_filter=True, activation=activation)
    ref_util.init_weights(netD)
    gan = GANLoss(gan_mode=loss, label_smoothing=True, tensor=torch.FloatTensor)
    gan.fake_label_tensor = torch.tensor([0.0])
    gan.zero_tensor = torch.tensor([0.0])
    optG = torch.optim.Adam(netG.parameters(), lr=1e-3, betas=(0.9, 0.99))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-3, betas=(0.9, 0.99))
    w_per = [0, .1, .3, .6]
    fixture = OrderedDict(
        meta=dict(gen=gen_name, nf=nf, size=size, batch=batch, loss=loss, steps=steps, seed=seed,
                  lambda_a=1.0, lambda_gp=lambda_gp, lambda_per=1.0, w_per=w_per, lr=1e-3, beta1=0.9,
                  torch=torch.__version__),
        init_G=OrderedDict((k, v.clone()) for k, v in netG.state_dict().items() if not k.startswith("clstm")),
        init_D=OrderedDict((k, v.clone()) for k, v in netD.state_dict().items()),
        steps=[])
    g = torch.Generator().manual_seed(seed + 1000)
    for step in range(steps):
        real_A, real_B = synthetic_batch(g, batch, size)
        alphas = []
        orig_rand = torch.rand

        def rec_rand(*a, **k):
            k.pop("device", None)
            t = orig_rand(*a, generator=g)
            alphas.append(t.clone())
            return t

        # ---- train.py:104-168 with the reference's own objects
        fake_B = netG(real_A)
        ref_util.set_requires_grad(nets=netD, requires_grad=True)
        optD.zero_grad()
        pred_fake = netD(real_A, fake_B.detach())
        pred_real = netD(real_A, real_B)
        loss_D_fake = gan(pred_fake, False, for_discriminator=True).mean()
        loss_D_real = gan(pred_real, True, for_discriminator=True).mean()
        loss_D = (loss_D_fake + loss_D_real) / 2
        rec = dict(loss_D=loss_D.item())  # inputs are regenerated from the seed (oracle.synthetic_batch)
        optD.zero_grad()
        torch.rand = rec_rand
        try:
            gp = ref_util.gradient_penalty(netD, real_A, real_B, fake_B, "cpu", 2, lambda_gp=lambda_gp)
        finally:
            torch.rand = orig_rand
        loss_D = loss_D + gp
        rec["gp"] = float(gp)
        rec["alpha"] = alphas[0] if alphas else None
        loss_D.backward(retain_graph=True)
        rec["grad_D"] = OrderedDict((k, p.grad.clone()) for k, p in netD.named_parameters() if p.grad is not None)
        optD.step()
        ref_util.set_requires_grad(nets=netD, requires_grad=False)
        optG.zero_grad()
        pred_fake = netD(real_A, fake_B)
        loss_G_GAN = gan(pred_fake, True, for_discriminator=False).mean()
        loss_G_L1 = nn.L1Loss()(real_B, fake_B)
        loss_G = loss_G_GAN + loss_G_L1 * 1.0
        features_fake = netD.get_intermediate_output()
        _ = netD(real_A, real_B)
        features_real = netD.get_intermediate_output()
        per = ref_util.pan_loss(features_real, features_fake, weights=w_per) * 1.0
        loss_G = loss_G + per
        loss_G.backward()
        rec.update(G_GAN=loss_G_GAN.item(), L1=loss_G_L1.item(), per=float(per))
        rec["fake_B_sub"] = fake_B.detach()[:, :, ::4, ::4].clone()
        rec["fake_B_mean_std"] = (float(fake_B.mean()), float(fake_B.std()))
        if step == 0:
            rec["real_label"] = None if gan.real_label_tensor is None else gan.real_label_tensor.clone()
            gkeep = (lambda v: dict(norm=float(v.norm()), head=v.flatten()[:8].clone())) if compact else (lambda v: v.clone())
            rec["grad_G"] = OrderedDict((k, gkeep(p.grad)) for k, p in netG.named_parameters() if p.grad is not None)
            if size <= 64:
                rec["features_fake"] = [f.clone() for f in features_fake]
            rec["features_fake_norms"] = [float(f.norm()) for f in features_fake]
        else:
            rec["grad_G_norms"] = OrderedDict((k, float(p.grad.norm())) for k, p in netG.named_parameters()
                                              if p.grad is not None)
            rec["grad_D"] = OrderedDict((k, float(v.norm())) for k, v in rec["grad_D"].items())
        optG.step()
        fixture["steps"].append(rec)
    big = compact or sum(v.numel() for v in netG.state_dict().values()) > 200_000
    keep = (lambda v: dict(norm=float(v.norm()), head=v.flatten()[:8].clone())) if big else (lambda v: v.clone())
    fixture["final_G"] = OrderedDict((k, keep(v)) for k, v in netG.state_dict().items() if not k.startswith("clstm"))
    fixture["final_D"] = OrderedDict((k, v.clone()) for k, v in netD.state_dict().items())
    fixture["optD_state"] = optD.state_dict()
    fixture["optD_param_names"] = [k for k, _ in netD.named_parameters()]
    return fixture


def vgg_blocks(seed):
    """The reference's four VGG16 slices (util.py:104-107) with seeded random-init weights: the pretrained
    ImageNet weights cannot be downloaded here (no network)."""
    import torchvision
    torch.manual_seed(seed)
    feats = torchvision.models.vgg16(weights=None).features
    return nn.ModuleList([feats[:4], feats[4:9], feats[9:16], feats[16:23]]).eval()


def vgg_case(seed=31, batch=2, size=64, channels=3):
    """VGGPerceptualLoss.forward (util.py:119-144) of the UNMODIFIED reference class, called unbound on a stand-in
    `self`: the class's __init__ hard-codes .cuda() and pretrained=True (download), so it cannot be constructed
    here, but forward only touches self.blocks / mean / std / transform / resize."""
    import types
    ref_util, _, _, _ = reference_modules()
    blocks = vgg_blocks(seed)
    stub = types.SimpleNamespace(blocks=blocks, mean=torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1),
                                 std=torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1),
                                 transform=torch.nn.functional.interpolate, resize=True)
    g = torch.Generator().manual_seed(seed + 1000)
    real_b = torch.rand(batch, channels, size, size, generator=g)
    fake_b = torch.rand(batch, channels, size, size, generator=g).requires_grad_(True)
    w_per = [0, .1, .3, .6]
    loss = ref_util.VGGPerceptualLoss.forward(stub, real_b, fake_b, weights=w_per)
    (grad,) = torch.autograd.grad(loss, fake_b)
    return dict(meta=dict(seed=seed, batch=batch, size=size, channels=channels, w_per=w_per,
                          torch=torch.__version__), loss=float(loss), grad_norm=float(grad.norm()),
                grad_sub=grad[:, :, ::4, ::4].clone(),
                weight_probe={k: float(v.flatten()[0]) for k, v in list(blocks.state_dict().items())[:4]})


def eval_case(seed=51):
    """eval_pair of the reference's test.py (lines 113-146). test.py cannot be imported (matplotlib / seaborn are
    not installed), so the function's own source is cut out of the file with ast and executed unmodified."""
    import ast
    import numpy as np
    src = open(os.path.join(REF, "test.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "eval_pair")
    ns = {"np": np}
    exec(compile(ast.Module([fn], []), os.path.join(REF, "test.py"), "exec"), ns)
    g = torch.Generator().manual_seed(seed)
    real = (torch.rand(3, 3, 64, 64, generator=g) > 0.7).float() * torch.rand(3, 3, 64, 64, generator=g)
    out = (real + 0.2 * torch.randn(3, 3, 64, 64, generator=g)).clamp(0, 1)
    return dict(meta=dict(seed=seed), results=[ns["eval_pair"](real[i], out[i]) for i in range(3)])


def convlstm_case(seed=61):
    """The reference's own ConvLSTMCell / ConvLSTM / ConvBLSTM classes (generators/BCDUNet.py:6-103), run on the CPU
    (their `self.device` falls back to cpu when CUDA is absent). Conv weights N(0, 0.1) so the gates leave the linear
    region; peepholes keep the reference's Xavier init."""
    sys.path.insert(0, REF)
    from generators import BCDUNet as ref
    out = dict(meta=dict(seed=seed, torch=torch.__version__), cases=OrderedDict())
    specs = [("cell_tanh", "cell", 24, 16, "tanh", 1), ("lstm_tanh", "lstm", 16, 8, "tanh", 3),
             ("lstm_relu", "lstm", 8, 8, "relu", 2), ("blstm_tanh", "blstm", 16, 16, "tanh", 3)]
    for name, kind, cin, cout, act, t in specs:
        torch.manual_seed(seed)
        frame, b = (16, 24), 2
        if kind == "cell":
            m = ref.ConvLSTMCell(cin, cout, (3, 3), (1, 1), act, frame)
        elif kind == "lstm":
            m = ref.ConvLSTM(cin, cout, (3, 3), (1, 1), act, frame, return_sequence=True)
        else:
            m = ref.ConvBLSTM(cin, cout, (3, 3), (1, 1), act, frame, return_sequence=True)
        with torch.no_grad():
            for k, p in m.named_parameters():
                if k.endswith("conv.weight"):
                    p.normal_(0, 0.1)
                elif k.endswith("conv.bias"):
                    p.normal_(0, 0.2)
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            if kind == "cell":
                x = torch.randn(b, cin, *frame, generator=g)
                h0 = torch.randn(b, cout, *frame, generator=g) * 0.5
                c0 = torch.randn(b, cout, *frame, generator=g)
                h, c = m(x, h0, c0)
                io = dict(x=x, h0=h0, c0=c0, h=h, c=c)
            else:
                x = torch.randn(b, t, cin, *frame, generator=g)
                io = dict(x=x, out=m(x))
        out["cases"][name] = dict(kind=kind, cin=cin, cout=cout, act=act, frame=frame,
                                  sd=OrderedDict((k, v.clone()) for k, v in m.state_dict().items()), **io)
    return out


def convlstm_grad_case(seed=71):
    """Gradients of the reference's own ConvLSTMCell / ConvLSTM / ConvBLSTM (generators/BCDUNet.py:6-103) under torch
    autograd on the CPU: loss = sum(out * G) with a seeded G; d loss / d input and d loss / d every parameter. The
    trainable device modules (tactile_gan_b200/convlstm.py: _SeqFn / _CellFn) are compared against these."""
    sys.path.insert(0, REF)
    from generators import BCDUNet as ref
    out = dict(meta=dict(seed=seed, torch=torch.__version__), cases=OrderedDict())
    specs = [("cell_tanh", "cell", 24, 16, "tanh", 1), ("lstm_tanh", "lstm", 16, 8, "tanh", 3),
             ("lstm_relu", "lstm", 8, 8, "relu", 2), ("blstm_tanh", "blstm", 16, 16, "tanh", 3),
             ("lstm_last", "lstm_last", 8, 16, "tanh", 4)]
    for name, kind, cin, cout, act, t in specs:
        torch.manual_seed(seed)
        frame, b = (16, 24), 2
        if kind == "cell":
            m = ref.ConvLSTMCell(cin, cout, (3, 3), (1, 1), act, frame)
        elif kind in ("lstm", "lstm_last"):
            m = ref.ConvLSTM(cin, cout, (3, 3), (1, 1), act, frame, return_sequence=kind == "lstm")
        else:
            m = ref.ConvBLSTM(cin, cout, (3, 3), (1, 1), act, frame, return_sequence=True)
        with torch.no_grad():
            for k, p in m.named_parameters():
                if k.endswith("conv.weight"):
                    p.normal_(0, 0.1)
                elif k.endswith("conv.bias"):
                    p.normal_(0, 0.2)
        g = torch.Generator().manual_seed(seed + 1)
        if kind == "cell":
            x = torch.randn(b, cin, *frame, generator=g).requires_grad_(True)
            h0 = (torch.randn(b, cout, *frame, generator=g) * 0.5).requires_grad_(True)
            c0 = torch.randn(b, cout, *frame, generator=g).requires_grad_(True)
            h, c = m(x, h0, c0)
            gh, gc = torch.randn(h.shape, generator=g), torch.randn(c.shape, generator=g)
            loss = (h * gh).sum() + (c * gc).sum()
            loss.backward()
            io = dict(x=x.detach(), h0=h0.detach(), c0=c0.detach(), gh=gh, gc=gc, dx=x.grad, dh0=h0.grad, dc0=c0.grad)
        else:
            x = torch.randn(b, t, cin, *frame, generator=g).requires_grad_(True)
            y = m(x)
            gy = torch.randn(y.shape, generator=g)
            (y * gy).sum().backward()
            io = dict(x=x.detach(), gy=gy, dx=x.grad)
        out["cases"][name] = dict(kind=kind, cin=cin, cout=cout, act=act, frame=frame,
                                  sd=OrderedDict((k, v.detach().clone()) for k, v in m.state_dict().items()),
                                  grads=OrderedDict((k, p.grad.clone()) for k, p in m.named_parameters()), **io)
    return out


def state_dict_keys():
    """Key/shape inventory of all reference networks at nf=64 (the checkpoint-layout contract)."""
    _, create_gen, _, create_disc = reference_modules()
    inv = OrderedDict()
    for name in ("UNet++", "UNet", "BCDUNet"):
        net = create_gen(name, 3, 3, 64, True)
        inv[name] = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    d = create_disc("patch", 3, 3, 64, return_filter=True, activation=True)
    inv["patch"] = OrderedDict((k, tuple(v.shape)) for k, v in d.state_dict().items())
    return inv


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    cases = [
        ("unetpp_ls", dict(gen_name="UNet++", nf=4, size=32, batch=2, loss="ls", steps=2, seed=21)),
        ("unet_ls", dict(gen_name="UNet", nf=4, size=256, batch=1, loss="ls", steps=1, seed=22)),
        ("bcdunet_ls", dict(gen_name="BCDUNet", nf=4, size=32, batch=2, loss="ls", steps=1, seed=23)),
        ("unetpp_hinge", dict(gen_name="UNet++", nf=4, size=32, batch=2, loss="hinge", steps=1, compact=True, seed=24)),
        ("unetpp_ce", dict(gen_name="UNet++", nf=4, size=32, batch=2, loss="ce", steps=1, compact=True, seed=25)),
        ("unetpp_w", dict(gen_name="UNet++", nf=4, size=32, batch=2, loss="w", steps=1, compact=True, seed=26)),
    ]
    for name, kw in cases:
        if len(sys.argv) > 1 and sys.argv[1] == "vgg":
            break                       # `make_golden.py vgg`: only (re)write the VGG fixtures
        fx = run_case(**kw)
        path = os.path.join(OUT, f"{name}.pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB", {k: round(v, 6) for k, v in fx["steps"][0].items()
                                                            if isinstance(v, float)})
    for name, kw in (("vgg_v1_rgb", dict(channels=3)), ("vgg_v1_gray", dict(seed=32, channels=1))):
        fx = vgg_case(**kw)
        torch.save(fx, os.path.join(OUT, f"{name}.pt"))
        print(name, fx["loss"], fx["grad_norm"])
    torch.save(eval_case(), os.path.join(OUT, "eval_pair_fuzzy.pt"))
    torch.save(convlstm_case(), os.path.join(OUT, "convlstm.pt"))
    torch.save(convlstm_grad_case(), os.path.join(OUT, "convlstm_grad.pt"))
    torch.save(state_dict_keys(), os.path.join(OUT, "state_dict_keys.pt"))
    print("wrote", OUT)
