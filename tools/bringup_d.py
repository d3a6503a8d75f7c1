"""GPU bring-up of the PatchDiscriminator engine (first-order backward and the gradient-penalty
double backward) against the CPU oracle. Usage: python tools/bringup_d.py [nf] [size] [n] [loss]"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402
from tactile_gan_b200 import _C  # noqa: E402
from tactile_gan_b200._C import F, ptr  # noqa: E402
from tactile_gan_b200.discriminators.discriminators import create_disc  # noqa: E402
from tactile_gan_b200.engine import PatchDInstance  # noqa: E402


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / (b.float().norm() + 1e-20)).item()


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    loss = sys.argv[4] if len(sys.argv) > 4 else "ls"
    act = loss == "ls"
    orc.QUANT["on"] = os.environ.get("QUANT", "1") == "1"
    torch.manual_seed(0)
    net = create_disc("patch", 3, 3, nf, True, act)
    for name, p in net.named_parameters():
        if p.dim() == 4:
            torch.nn.init.normal_(p, 0.0, 0.05)
        elif name.endswith("weight"):
            torch.nn.init.normal_(p, 1.0, 0.1)
        else:
            torch.nn.init.normal_(p, 0.0, 0.1)
    sd = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
    g = torch.Generator().manual_seed(1)
    a, b = orc.synthetic_batch(g, n, size)
    fake = torch.rand(n, 3, size, size, generator=g)
    alpha = torch.rand(n, 1, generator=g)
    # ---------------- oracle
    psd = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    names = list(psd.keys())
    pred_r, feats_r = orc.patchd_forward(psd, a, b, act)
    label = orc.make_real_label(pred_r.shape, True, generator=g)
    pred_f, _ = orc.patchd_forward(psd, a, fake, act)
    loss_d = (orc.gan_loss(pred_f, False, loss, label) + orc.gan_loss(pred_r, True, loss, label)) / 2
    gd = dict(zip(names, torch.autograd.grad(loss_d, [psd[k] for k in names], retain_graph=True)))
    gp = orc.gradient_penalty(psd, a, b, fake, alpha, act, 0.01)
    ggp = dict(zip(names, torch.autograd.grad(gp, [psd[k] for k in names], allow_unused=True)))
    # input gradient of the G-side GAN loss
    fake_g = fake.clone().requires_grad_(True)
    pred_g, _ = orc.patchd_forward(OrderedDict((k, v.detach()) for k, v in psd.items()), a, fake_g, act)
    lg = orc.gan_loss(pred_g, True, loss, label, for_discriminator=False)
    (gin,) = torch.autograd.grad(lg, fake_g)
    # ---------------- engine
    net = net.cuda()
    dev = "cuda"
    A = PatchDInstance(net, 2 * n, size, size, backward=True)
    S = PatchDInstance(net, n, size, size, backward=True, second_order=True)
    a2 = torch.cat([a, a]).cuda().contiguous()
    fb = torch.cat([fake, b]).cuda().contiguous()
    A.pack_input(a2, fb)
    pred = A.forward()
    u5 = A.u[4]
    hw5 = u5.ho * u5.wo
    got_pred = pred[..., 0].float().cpu()
    print(f"pred fake rel={rel(got_pred[:n], pred_f[:, 0]):.3e}  real rel={rel(got_pred[n:], pred_r[:, 0]):.3e}")
    for i, (f, fr) in enumerate(zip(A.features(), feats_r)):
        print(f"  feature{i} (real half) rel={rel(f.buf[n:, :, :, :f.c].permute(0, 3, 1, 2), fr):.3e}")
    losses = torch.zeros(8, device=dev)
    lab = label.cuda().contiguous()
    mode = _C.GAN_MODES[loss]
    A.store.zero_grad()
    u5.dz.zero_()
    _C.call("gan_loss", ptr(pred), None, F(0.0), mode, 0, 1, int(A.has_sigmoid), F(0.5), 0, n, hw5, u5.c,
            ptr(losses[0:1]), ptr(u5.dz))
    _C.call("gan_loss", ptr(pred), ptr(lab), F(1.0), mode, 1, 1, int(A.has_sigmoid), F(0.5), n, 2 * n, hw5, u5.c,
            ptr(losses[0:1]), ptr(u5.dz))
    A.backward(wgrad=True)
    torch.cuda.synchronize()
    print(f"loss_D got={losses[0].item():.6f} ref={loss_d.item():.6f}")
    got = A.store.grads_by_name()
    for k in names:
        print(f"  dLoss/d {k:18s} rel={rel(got[k], gd[k]):.3e} |ref|={gd[k].norm().item():.3e}")
    # ---------------- gradient penalty
    al = ((alpha + 1) / 2).view(-1).cuda().contiguous()
    one_minus = (1 - al).contiguous()
    S.store.zero_grad()
    S.pack_input(a.cuda().contiguous(), b.cuda().contiguous(), wa=al, b2=fake.cuda().contiguous(), wb=one_minus)
    S.forward()
    S.gp_first_backward()
    S.gp_penalty(3, 3, 0.01, 1.0, losses[1:2])
    S.gp_second_backward()
    torch.cuda.synchronize()
    print(f"gp got={losses[1].item():.6e} ref={gp.item():.6e}")
    got = S.store.grads_by_name()
    for k in names:
        if ggp[k] is None:
            print(f"  dGP/d {k:18s} ref None, got |g|={got[k].norm().item():.3e}")
        else:
            print(f"  dGP/d {k:18s} rel={rel(got[k], ggp[k]):.3e} |ref|={ggp[k].norm().item():.3e}")
    # ---------------- G-side input gradient
    S.pack_input(a.cuda().contiguous(), fake.cuda().contiguous())
    p2 = S.forward()
    S.u[4].dz.zero_()
    _C.call("gan_loss", ptr(p2), ptr(lab), F(1.0), mode, 1, 0, int(S.has_sigmoid), F(1.0), 0, n, hw5, u5.c,
            ptr(losses[2:3]), ptr(S.u[4].dz))
    dx0 = S.backward(wgrad=False, input_grad=True)
    torch.cuda.synchronize()
    print(f"G_GAN got={losses[2].item():.6f} ref={lg.item():.6f}")
    print(f"  d G_GAN / d fake rel={rel(dx0[..., 3:6].permute(0, 3, 1, 2), gin):.3e} |ref|={gin.norm().item():.3e}")
    print("error flag", _C.error_flag())


if __name__ == "__main__":
    main()
