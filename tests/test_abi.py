"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares; host-side
index math (tap tables, tile choice, module key layout) -- no kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    return ctypes.CDLL(ge.LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tactile_gan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.tg_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    from tactile_gan_b200 import _C
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", "/nonexistent/libtactile_gan_b200.so")
    with pytest.raises(_C.TgError):
        _C.lib()


def test_cpu_tensor_is_rejected():
    from tactile_gan_b200 import _C
    from tactile_gan_b200.generators.generators import create_gen
    net = create_gen("UNet++", 3, 3, 4, True)
    with pytest.raises(_C.TgError):
        net(torch.zeros(1, 3, 32, 32))


def test_tile_choice(lib):
    out = (ctypes.c_int * 4)()
    for n, h, w, stats in [(32, 256, 256, 1), (32, 127, 127, 0), (2, 59, 59, 1), (4, 16, 16, 1), (8, 4, 4, 0),
                           (64, 2, 2, 0)]:
        assert lib.tg_conv_query_tiles(n, h, w, stats, out) == 0
        th, tw, tn, tpi = out
        assert th * tw * tn == 128
        if stats:
            assert tn == 1
        generic = -(-h // th) * -(-w // tw)
        halo = -(-h // 16) * -(-w // 8)           # the halo-resident kernels tile 16x8
        assert tpi == (max(generic, halo) if stats else generic)


def test_policy_setters_return_the_previous_value(lib):
    """tg_in_stream_policy / tg_in_stream_slim / tg_in_stream_serpentine / tg_pdl_policy are host-side switches: they
    return the value they replace, ignore out-of-range arguments (query only) and need no GPU."""
    for fn, values, query in ((lib.tg_in_stream_policy, (0, 2, 1), -1), (lib.tg_in_stream_slim, (1, 0), -1),
                              (lib.tg_in_stream_serpentine, (5, 7, 0), 9), (lib.tg_pdl_policy, (1, 2, 0), -1)):
        first = fn(query)
        assert fn(query) == first                      # a query changes nothing
        prev = first
        for v in values:
            assert fn(v) == prev
            assert fn(query) == v
            prev = v
        fn(first)
        assert fn(query) == first


def test_phase_taps_cover_transposed_conv():
    """Every (output pixel, tap) pair of a stride-2 transposed conv appears in exactly one phase."""
    from tactile_gan_b200.layers import phase_taps
    for k, pad in [(3, 0), (4, 1), (2, 0)]:
        s = 2
        hin = 5
        hout = (hin - 1) * s - 2 * pad + k
        ref = {}
        for i in range(hin):
            for r in range(k):
                o = i * s - pad + r
                if 0 <= o < hout:
                    ref.setdefault(o, set()).add((i, r))
        got = {}
        for py in range(s):
            taps = phase_taps(k, 1, pad, s, py, 0, flipped=False) if False else phase_taps(k, k, pad, s, py, py, False)
            for a in range((hout - py + s - 1) // s):
                o = s * a + py
                for dy, dx, widx in taps:
                    r = widx // k
                    if widx % k != r:       # look at the diagonal taps only (1-D check on the y axis)
                        continue
                    i = a + dy
                    if 0 <= i < hin:
                        got.setdefault(o, set()).add((i, r))
        assert got == ref, (k, pad)


def test_state_dict_layout_matches_reference(golden_dir):
    from tactile_gan_b200.discriminators.discriminators import create_disc
    from tactile_gan_b200.generators.generators import create_gen
    inv = torch.load(os.path.join(golden_dir, "state_dict_keys.pt"))
    for name in ["UNet++"] + [n for n in ("UNet", "BCDUNet") if _has_gen(n)]:
        sd = create_gen(name, 3, 3, 64, True).state_dict()
        assert list(sd.keys()) == list(inv[name].keys()), name
        assert all(tuple(sd[k].shape) == inv[name][k] for k in sd), name
    sd = create_disc("patch", 3, 3, 64, True, True).state_dict()
    assert list(sd.keys()) == list(inv["patch"].keys())
    assert all(tuple(sd[k].shape) == inv["patch"][k] for k in sd)
    with pytest.raises(NameError):
        create_gen("resnet", 3, 3, 64)
    with pytest.raises(NameError):
        create_disc("pixel", 3, 3, 64, True)


def _has_gen(name):
    mod = {"UNet": "UNet", "BCDUNet": "BCDUNet"}[name]
    return os.path.exists(os.path.join(ROOT, "tactile_gan_b200", "generators", mod + ".py"))


def test_train_cli_surface():
    """The 27 reference flags with the reference's (code) defaults (train.py:231-257)."""
    from tactile_gan_b200.train import build_parser
    o = build_parser().parse_args([])
    exp = dict(data="./data", batch_size=4, input_dim=3, output_dim=3, initial_epoch=1, total_epochs=135,
               epoch_constant=25, lr=0.001, no_label_smoothing=False, beta1=0.9, threads=8, lambda_a=1,
               lambda_gp=0.01, lambda_per=1, w_per=[0, .1, .3, .6], gen="UNet++", nf=64, loss="ls", no_aug=False,
               target="rgb", version=1, folder_save="pix2obj", folder_load="pix2obj", checkpoint_interval=-1,
               continue_training=False, reg_every=1)
    for k, v in exp.items():
        assert getattr(o, k) == v, k


@pytest.mark.parametrize("mode", ["ls", "ce", "w", "hinge"])
def test_ganloss_class_matches_oracle(mode):
    """GANLoss (compat surface) against the oracle's restatement on CPU tensors, incl. error behaviour."""
    import oracle as orc
    from tactile_gan_b200.generators.generators import GANLoss
    g = torch.Generator().manual_seed(0)
    pred = torch.randn(2, 1, 7, 7, generator=g)
    gl = GANLoss(gan_mode=mode, label_smoothing=False)
    one = torch.ones(1)
    for real, disc in ((True, True), (False, True), (True, False)):
        assert float(gl(pred, real, for_discriminator=disc)) == pytest.approx(
            float(orc.gan_loss(pred, real, mode, one, for_discriminator=disc)), rel=1e-6, abs=1e-7)
    with pytest.raises(ValueError):
        GANLoss(gan_mode="nope")
    torch.manual_seed(3)
    sm = GANLoss(gan_mode="ls", label_smoothing=True)
    a = sm.get_target_tensor(pred, True)
    assert a.shape == pred.shape and float(a.max()) <= 1.0 and sm.get_target_tensor(pred, True) is not None
    assert torch.equal(sm.real_label_tensor, a)     # cached: drawn once


def test_convlstm_modules_keep_reference_parameter_names(golden_dir):
    """ConvLSTMCell / ConvLSTM / ConvBLSTM expose exactly the reference's state_dict keys and shapes (generators/
    BCDUNet.py:6-103; the fixture holds the reference modules' own state_dicts) and refuse CPU tensors."""
    import os
    import torch
    from tactile_gan_b200 import _C
    from tactile_gan_b200.generators.BCDUNet import ConvBLSTM, ConvLSTM, ConvLSTMCell
    fx = torch.load(os.path.join(golden_dir, "convlstm.pt"), weights_only=False)
    cls = {"cell": ConvLSTMCell, "lstm": ConvLSTM, "blstm": ConvBLSTM}
    for name, c in fx["cases"].items():
        m = cls[c["kind"]](c["cin"], c["cout"], (3, 3), (1, 1), c["act"], c["frame"])
        sd = m.state_dict()
        assert list(sd.keys()) == list(c["sd"].keys()), name
        assert all(sd[k].shape == c["sd"][k].shape for k in sd), name
        m.load_state_dict(c["sd"])
    with pytest.raises(_C.TgError):
        m(fx["cases"]["blstm_tanh"]["x"])


def test_inference_chunk_size():
    from tactile_gan_b200 import bridge
    assert bridge.EngineModule.infer_chunk(256, 256) == 64 and bridge.EngineModule.infer_chunk(512, 512) == 16
    assert bridge.EngineModule.infer_chunk(4096, 4096) == 1


def test_lr_schedule_milestones_match_the_reference_formula():
    """SURVEY 8 a13 / reference train.py:191-195: MultiStepLR, gamma 0.8, milestones
    int16(linspace(epoch_constant, total_epochs, 11)[:-1]) = 25, 36, ..., 124 for the default 135-epoch run."""
    import argparse
    import torch
    from tactile_gan_b200 import train as tg_train
    old = tg_train.opt
    try:
        for const, total, want in ((25, 135, [25, 36, 47, 58, 69, 80, 91, 102, 113, 124]),
                                   (1, 2, [1] * 10), (10, 50, [10, 14, 18, 22, 26, 30, 34, 38, 42, 46])):
            tg_train.opt = argparse.Namespace(epoch_constant=const, total_epochs=total)
            p = torch.nn.Parameter(torch.zeros(1))
            opt_ = torch.optim.SGD([p], lr=1e-3)
            sched = tg_train.Train_GAN.get_scheduler(opt_)
            assert sorted(sched.milestones.elements()) == want
            lrs = []
            for _ in range(total):
                lrs.append(opt_.param_groups[0]["lr"])
                opt_.step()
                sched.step()
            # epoch e (1-based, train.py:85-87) runs with lr * 0.8^(number of milestones <= e - 1)
            for e, lr in enumerate(lrs, start=1):
                assert lr == pytest.approx(1e-3 * 0.8 ** sum(1 for m in want if m <= e - 1), rel=1e-12), (const, total, e)
    finally:
        tg_train.opt = old


def test_factories_reject_shapes_the_engines_cannot_serve():
    """limits the reference does not have are raised at construction with a clear message."""
    from tactile_gan_b200.discriminators.discriminators import create_disc
    from tactile_gan_b200.generators.generators import create_gen
    for name in ("UNet", "UNet++", "BCDUNet"):
        with pytest.raises(ValueError):
            create_gen(name, 3, 3, 128)
        with pytest.raises(ValueError):
            create_gen(name, 3, 5, 16)
    with pytest.raises(ValueError):
        create_disc("patch", 4, 4, 16, True)
    with pytest.raises(NameError):
        create_gen("resnet", 3, 3, 16)
    with pytest.raises(NameError):
        create_disc("pixel", 3, 3, 16, True)
    assert create_gen("unet++", 3, 3, 8) is not None and create_disc("patch", 3, 3, 8, True) is not None
