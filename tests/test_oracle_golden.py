"""The oracle (oracle/oracle.py, a functional restatement) must reproduce what the UNMODIFIED reference
modules produced when oracle/make_golden.py ran them (fixtures under tests/golden/)."""
import os
from collections import OrderedDict

import pytest
import torch

import oracle as orc

CASES = ["unetpp_ls", "unet_ls", "bcdunet_ls", "unetpp_hinge", "unetpp_ce", "unetpp_w"]
TOL = dict(rtol=2e-4, atol=2e-6)  # same fp32 torch ops, different op order / fusion


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _replay(fx):
    m = fx["meta"]
    cfg = orc.StepConfig(gen=m["gen"], loss=m["loss"], version=2, lambda_a=m["lambda_a"], lambda_gp=m["lambda_gp"],
                         lambda_per=m["lambda_per"], w_per=m["w_per"], lr=m["lr"], beta1=m["beta1"])
    sd_g = OrderedDict((k, v.clone()) for k, v in fx["init_G"].items())
    sd_d = OrderedDict((k, v.clone()) for k, v in fx["init_D"].items())
    opt_g, opt_d = {}, {}
    g = torch.Generator().manual_seed(m["seed"] + 1000)
    outs = []
    label = None
    for rec in fx["steps"]:
        real_a, real_b = orc.synthetic_batch(g, m["batch"], m["size"])
        alpha = torch.rand(m["batch"], 1, generator=g)  # same draw order as the fixture generator
        assert torch.equal(alpha, rec["alpha"])
        if label is None:
            label = rec.get("real_label")
            if label is None:
                label = torch.ones(1)
        outs.append(orc.train_step(sd_g, sd_d, opt_g, opt_d, real_a, real_b, label, alpha, cfg))
    return outs, sd_g, sd_d


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(golden_dir, name):
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    fx = _load(golden_dir, name)
    outs, sd_g, sd_d = _replay(fx)
    for out, rec in zip(outs, fx["steps"]):
        for k in ("loss_D", "gp", "G_GAN", "L1", "per"):
            assert out[k] == pytest.approx(rec[k], rel=2e-4, abs=1e-6), k
        torch.testing.assert_close(out["fake_B"][:, :, ::4, ::4], rec["fake_B_sub"], **TOL)
    rec0, out0 = fx["steps"][0], outs[0]
    for k, ref in rec0["grad_D"].items():
        torch.testing.assert_close(out0["grads_D"][k], ref, rtol=1e-3, atol=1e-7, msg=lambda s: f"grad_D {k}: {s}")
    for k, ref in rec0["grad_G"].items():
        got = out0["grads_G"][k]
        if isinstance(ref, dict):
            assert float(got.norm()) == pytest.approx(ref["norm"], rel=1e-3, abs=1e-8), k
            torch.testing.assert_close(got.flatten()[:8], ref["head"], rtol=2e-3, atol=1e-7)
        else:
            torch.testing.assert_close(got, ref, rtol=2e-3, atol=1e-7, msg=lambda s: f"grad_G {k}: {s}")
    for k, ref in fx["final_D"].items():
        torch.testing.assert_close(sd_d[k], ref, rtol=1e-3, atol=2e-5, msg=lambda s: f"final_D {k}: {s}")
    for k, ref in fx["final_G"].items():
        if isinstance(ref, dict):
            assert float(sd_g[k].norm()) == pytest.approx(ref["norm"], rel=1e-3), k
        else:
            torch.testing.assert_close(sd_g[k], ref, rtol=1e-3, atol=2e-5, msg=lambda s: f"final_G {k}: {s}")


def test_state_dict_key_inventory(golden_dir):
    """Checkpoint-layout contract: 92 / 86 / 66 / 13 keys."""
    inv = _load(golden_dir, "state_dict_keys")
    assert len(inv["UNet++"]) == 92 and len(inv["UNet"]) == 86 and len(inv["BCDUNet"]) == 66 and len(inv["patch"]) == 13


def vgg_state_dict(seed):
    """Seeded random-init VGG16 slices, keyed like the reference's VGGPerceptualLoss.state_dict()
    (same recipe as oracle/make_golden.py:vgg_blocks; pretrained weights are not available offline)."""
    import torchvision
    torch.manual_seed(seed)
    f = torchvision.models.vgg16(weights=None).features
    blocks = torch.nn.ModuleList([f[:4], f[4:9], f[9:16], f[16:23]])
    return OrderedDict(("blocks." + k, v.detach().clone()) for k, v in blocks.state_dict().items())


@pytest.mark.parametrize("name", ["vgg_v1_rgb", "vgg_v1_gray"])
def test_oracle_vgg_perceptual_matches_reference_forward(golden_dir, name):
    """oracle.vgg_perceptual against the fixture produced by the reference's own VGGPerceptualLoss.forward."""
    fx = _load(golden_dir, name)
    m = fx["meta"]
    sd = vgg_state_dict(m["seed"])
    for k, v in fx["weight_probe"].items():
        assert float(sd["blocks." + k].flatten()[0]) == v, "seeded VGG init differs from the fixture's"
    g = torch.Generator().manual_seed(m["seed"] + 1000)
    real_b = torch.rand(m["batch"], m["channels"], m["size"], m["size"], generator=g)
    fake_b = torch.rand(m["batch"], m["channels"], m["size"], m["size"], generator=g).requires_grad_(True)
    loss = orc.vgg_perceptual(sd, real_b, fake_b, m["w_per"])
    (grad,) = torch.autograd.grad(loss, fake_b)
    assert float(loss) == pytest.approx(fx["loss"], rel=1e-5)
    assert float(grad.norm()) == pytest.approx(fx["grad_norm"], rel=1e-4)
    torch.testing.assert_close(grad[:, :, ::4, ::4], fx["grad_sub"], rtol=1e-3, atol=1e-9)


def eval_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    real = (torch.rand(3, 3, 64, 64, generator=g) > 0.7).float() * torch.rand(3, 3, 64, 64, generator=g)
    out = (real + 0.2 * torch.randn(3, 3, 64, 64, generator=g)).clamp(0, 1)
    return real, out


def test_oracle_eval_pair_matches_reference_function(golden_dir):
    """oracle.eval_pair_fuzzy against the fixture produced by executing the reference's own eval_pair source."""
    fx = _load(golden_dir, "eval_pair_fuzzy")
    real, out = eval_inputs(fx["meta"]["seed"])
    for i, ref in enumerate(fx["results"]):
        got = orc.eval_pair_fuzzy(real[i], out[i])
        for k in ("accuracy", "dice", "jaccard"):
            assert got[k] == pytest.approx(float(ref[k]), rel=1e-6), k


def test_oracle_convlstm_matches_reference_classes(golden_dir):
    """oracle.convlstm_cell / convlstm / convblstm against the reference's own ConvLSTMCell / ConvLSTM / ConvBLSTM
    (generators/BCDUNet.py:6-103) run by oracle/make_golden.py; same fp32 torch ops -> tight tolerance."""
    fx = _load(golden_dir, "convlstm")
    for name, c in fx["cases"].items():
        if c["kind"] == "cell":
            h, cc = orc.convlstm_cell(c["sd"], c["x"], c["h0"], c["c0"], c["act"])
            torch.testing.assert_close(h, c["h"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(cc, c["c"], rtol=1e-5, atol=1e-6)
        else:
            fn = orc.convlstm if c["kind"] == "lstm" else orc.convblstm
            torch.testing.assert_close(fn(c["sd"], c["x"], c["act"]), c["out"], rtol=1e-5, atol=1e-6)
            last = fn(c["sd"], c["x"], c["act"], return_sequence=False)
            assert torch.equal(last, fn(c["sd"], c["x"], c["act"])[:, -1])


def test_oracle_convlstm_gradients_match_reference_autograd(golden_dir):
    """The oracle's ConvLSTM restatement is differentiable torch code: its autograd gradients must equal the ones
    torch computed on the reference's own classes (tests/golden/convlstm_grad.pt) -- this pins the backward the
    device modules are tested against."""
    fx = _load(golden_dir, "convlstm_grad")
    for name, c in fx["cases"].items():
        sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
        x = c["x"].clone().requires_grad_(True)
        if c["kind"] == "cell":
            h0, c0 = c["h0"].clone().requires_grad_(True), c["c0"].clone().requires_grad_(True)
            h, cc = orc.convlstm_cell(sd, x, h0, c0, c["act"])
            ((h * c["gh"]).sum() + (cc * c["gc"]).sum()).backward()
            assert torch.allclose(h0.grad, c["dh0"], atol=1e-5) and torch.allclose(c0.grad, c["dc0"], atol=1e-5)
        else:
            if c["kind"] == "blstm":
                y = orc.convblstm(sd, x, c["act"])
            else:
                y = orc.convlstm(sd, x, c["act"], return_sequence=c["kind"] == "lstm")
            (y * c["gy"]).sum().backward()
        assert torch.allclose(x.grad, c["dx"], atol=1e-5), name
        for k, g in c["grads"].items():
            assert torch.allclose(sd[k].grad, g, atol=2e-4, rtol=1e-4), (name, k)
