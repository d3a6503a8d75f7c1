"""BASELINE ARM -- never imported by the product path (tactile_gan_b200/*). Times the UNMODIFIED reference modules
(baseline/_ref/: generators.create_gen, discriminators.create_disc, util.init_weights / gradient_penalty / pan_loss /
set_requires_grad, generators.GANLoss, torch.optim.Adam) driven through the loop body of the reference's
train.py:99-168, restated only as far as train.py cannot be imported (it needs albumentations, hard-codes cuda:0 and
reads a global `opt`). Runs on the host cores (bench.py --impl reference, cpu_baseline) and, as the bar that matters,
on the same B200 under PyTorch eager + cuDNN in fp32 / TF32 / bf16 autocast (bench.py's `cudnn_baseline` leg).

Kept from the reference: module construction and init (train.py:36-43), GANLoss with label smoothing (:45), the five
`.item()` reads per iteration (:121,129,148,151,163 -- they synchronise the device every step, as the reference does),
`retain_graph=True` on loss_D.backward (:134), the wasted generator backward inside the D step (util.py:83 builds the
interpolate from the attached fake_B), set_requires_grad toggles, both Adam steps.
"""
import contextlib
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "generators"))


def reference_modules():
    """Import the reference's own modules from baseline/_ref (their names -- util, generators, discriminators -- are
    top-level, so they are imported with baseline/_ref in front of sys.path and that entry is removed again)."""
    if not available():
        raise RuntimeError("baseline/_ref is empty: run `python baseline/install_ref.py` where /root/reference exists")
    sys.path.insert(0, REF)
    try:
        import util as ref_util  # noqa: seeds torch / numpy / random with 21 at import (util.py:8-11)
        from discriminators.discriminators import create_disc
        from generators.generators import GANLoss, create_gen
    finally:
        sys.path.remove(REF)
    assert os.path.abspath(ref_util.__file__).startswith(REF), ref_util.__file__
    return ref_util, create_gen, GANLoss, create_disc


class RefStep:
    def __init__(self, gen="UNet++", nf=64, device="cpu", loss="ls", version=2, lambda_a=1.0, lambda_gp=0.01,
                 lambda_per=1.0, w_per=(0, .1, .3, .6), lr=1e-3, beta1=0.9, label_smoothing=True, autocast=None,
                 seed=21):
        self.u, create_gen, GANLoss, create_disc = reference_modules()
        self.device = torch.device(device)
        torch.manual_seed(seed)
        act = False if loss in ("w", "hinge") else (loss != "ce")               # train.py:33
        self.netG = create_gen(gen, 3, 3, nf, act).to(self.device)
        self.u.init_weights(self.netG)
        self.netD = create_disc("patch", 3, 3, nf, return_filter=version == 2, activation=act).to(self.device)
        self.u.init_weights(self.netD)
        cuda = self.device.type == "cuda"
        self.gan = GANLoss(gan_mode=loss, label_smoothing=label_smoothing,
                           tensor=torch.cuda.FloatTensor if cuda else torch.FloatTensor)
        if not cuda:
            # GANLoss hard-codes device='cuda' for these two cached tensors (generators.py:59,67,75); its
            # cache-on-None pattern lets them be pre-seeded without touching the reference's code
            self.gan.fake_label_tensor = torch.tensor([0.0])
            self.gan.zero_tensor = torch.tensor([0.0])
        self.optG = torch.optim.Adam(self.netG.parameters(), lr=lr, betas=(beta1, 0.99))
        self.optD = torch.optim.Adam(self.netD.parameters(), lr=lr, betas=(beta1, 0.99))
        self.version, self.lambda_a, self.lambda_gp, self.lambda_per = version, lambda_a, lambda_gp, lambda_per
        self.w_per = list(w_per)
        self.autocast = autocast                         # None | torch.bfloat16 (cuda only)

    def _amp(self):
        if self.autocast is None:
            return contextlib.nullcontext()
        return torch.autocast(self.device.type, dtype=self.autocast)

    def step(self, real_A, real_B, regularize=True):
        u, netG, netD, gan = self.u, self.netG, self.netD, self.gan
        out = {}
        with self._amp():
            fake_B = netG(real_A)
            u.set_requires_grad(nets=netD, requires_grad=True)
            self.optD.zero_grad()
            pred_fake = netD(real_A, fake_B.detach())
            pred_real = netD(real_A, real_B)
            loss_D_fake = gan(pred_fake, False, for_discriminator=True).mean()
            loss_D_real = gan(pred_real, True, for_discriminator=True).mean()
            loss_D = (loss_D_fake + loss_D_real) / 2
            out["loss_D"] = loss_D.item()
            if regularize and self.lambda_gp != 0:
                self.optD.zero_grad()
                gp = u.gradient_penalty(netD, real_A, real_B, fake_B, self.device, self.version,
                                        lambda_gp=self.lambda_gp)
                loss_D = loss_D + gp
                out["gp"] = gp.item()
            else:
                out["gp"] = 0.0
        loss_D.backward(retain_graph=self.lambda_gp != 0)
        self.optD.step()
        with self._amp():
            u.set_requires_grad(nets=netD, requires_grad=False)
            self.optG.zero_grad()
            pred_fake = netD(real_A, fake_B)
            loss_G_GAN = gan(pred_fake, True, for_discriminator=False).mean()
            out["G_GAN"] = loss_G_GAN.item()
            loss_G_L1 = nn.L1Loss()(real_B, fake_B)
            out["L1"] = loss_G_L1.item()
            loss_G = loss_G_GAN + loss_G_L1 * self.lambda_a
            if self.lambda_per != 0 and self.version == 2:
                features_fake = netD.get_intermediate_output()
                _ = netD(real_A, real_B)
                features_real = netD.get_intermediate_output()
                per = u.pan_loss(features_real, features_fake, weights=self.w_per) * self.lambda_per
                loss_G = loss_G + per
                out["per"] = per.item()
            else:
                out["per"] = 0.0
        loss_G.backward()
        self.optG.step()
        return out
