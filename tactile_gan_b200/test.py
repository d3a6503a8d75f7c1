"""Drop-in for the inference part of the reference's test.py (`--folder`, params.txt + final_model.pth):
load_opt / load_model keep the reference signatures (test.py:24-42) and the generator forward
(test.py:202-203) runs on the sm_100a engine, batched. The fuzzy accuracy / Dice / Jaccard of eval_pair
(test.py:113-124) are reduced on the device right after the forward (tg_eval_fuzzy) and written to eval.txt in
the reference's format (test.py:175-186). The matplotlib / seaborn plots, the PNG montage and the thresholded
(non-fuzzy, Otsu) variant stay out of scope: outputs are written as .npy (and PNG when PIL is
available). Additive flags: --synthetic N, --batch."""
import argparse
import json
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

from .generators.generators import create_gen
from .util import mkdir


class Opt:
    def __init__(self, dictionary):
        for k, v in dictionary.items():
            setattr(self, k, v)


def load_opt(path):
    with open(path) as f:
        return Opt(json.load(f))


def load_model(model_path, opt, device):
    """Reference quirk kept: the generator is rebuilt with the default activation=True whatever it was
    trained with (test.py:37)."""
    gen = create_gen(opt.gen, opt.input_dim, opt.output_dim, opt.nf, multigpu=False)
    gen.to(device)
    checkpoint = torch.load(model_path, map_location=device)
    gen.load_state_dict(checkpoint["gen"], strict=False)
    return gen


def load_arrays(path):
    names = {"gen": "genloss", "disc": "discloss", "l1": "l1loss", "gp": "gploss", "per": "perloss"}
    return {k: np.load(os.path.join(path, v + ".npy")) for k, v in names.items()}


def unnormalize(a):
    return a / 2 + 0.5


def fuzzy_sums(out, real):
    """Per-image (sum o*r, sum o^2+r^2, sum min(o,r), sum r) on the device: (N, 4) fp32 tensor."""
    from . import _C
    o, r = out.detach().contiguous().float(), real.detach().contiguous().float()
    if not (o.is_cuda and r.is_cuda):
        raise _C.TgError("eval_pair runs its reductions on CUDA (sm_100a); there is no CPU fallback")
    n = o.shape[0]
    stats = torch.zeros(n, 4, device=o.device)
    _C.call("eval_fuzzy", _C.ptr(o), _C.ptr(r), n, _C.LL(o.numel() // n), _C.ptr(stats))
    return stats


def metrics_from_sums(stats):
    """test.py:117-123: accuracy = sum min(o,r) / sum r, jaccard = I / (D - I), dice = 2 I / D."""
    s = stats.double().cpu().numpy()
    inter, denom, mn, sr = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
    return {"accuracy": mn / sr, "dice": 2 * inter / denom, "jaccard": inter / (denom - inter)}


def eval_pair(real, out, thresh=None, fuzzy=True):
    """reference test.py:113-146 for ONE (C,H,W) pair; only the fuzzy variant (the one test_model uses) is built."""
    if not fuzzy:
        raise NotImplementedError("the thresholded / Otsu variant of eval_pair is not built (unused by test_model)")
    dev = out.device if out.is_cuda else (real.device if real.is_cuda else torch.device("cuda:0"))
    m = metrics_from_sums(fuzzy_sums(out.to(dev).unsqueeze(0), real.to(dev).unsqueeze(0)))
    return {k: float(v[0]) for k, v in m.items()}


def print_evaluation(accuracy, dice, jaccard, output_path):
    """eval.txt in the reference's format (test.py:175-186); the distribution plots are out of scope."""
    a = f"Pixel Accuracy => min:{np.min(accuracy)}, max:{np.max(accuracy)}, avg:{np.mean(accuracy)}, std:{np.std(accuracy)}\n"
    d = f"Dice Coeff => min:{np.min(dice)}, max:{np.max(dice)}, avg:{np.mean(dice)}, std:{np.std(dice)}\n"
    j = f"Jaccard Index => min:{np.min(jaccard)}, max:{np.max(jaccard)}, avg:{np.mean(jaccard)}, std:{np.std(jaccard)}\n"
    with open(os.path.join(output_path, "eval.txt"), 'w') as f:
        f.writelines([a, d, j])
    print(f"Acc: {np.mean(accuracy)}, IoU: {np.mean(jaccard)}, Dice: {np.mean(dice)}")


def test_model(model, dataset, output_path, evaluation=False, device=None):
    """Generator forward over the dataset (any batch size); saves out/<i>.npy (+ .png with PIL). With
    `evaluation` the fuzzy sums are reduced on the device after each forward and read back once at the end;
    returns (accuracy, dice, jaccard) lists like the reference (test.py:188-230)."""
    device = device or next(model.parameters()).device
    mkdir(os.path.join(output_path, "out"))
    try:
        from PIL import Image
    except Exception:
        Image = None
    idx, sums = 0, []
    for batch in dataset:
        real_A, real_B = batch[0], batch[1]
        with torch.no_grad():
            out_dev = model(real_A.to(device).float().contiguous())
            if evaluation:
                sums.append(fuzzy_sums(out_dev, real_B.to(device)))
            out = out_dev.cpu()
        for j in range(out.shape[0]):
            idx += 1
            np.save(os.path.join(output_path, "out", f"{idx}.npy"), out[j].numpy())
            if Image is not None and out.shape[1] == 3:
                img = (out[j].clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype(np.uint8)
                Image.fromarray(img).save(os.path.join(output_path, "out", f"{idx}.png"))
    if not sums:
        return [], [], []
    m = metrics_from_sums(torch.cat(sums))
    return list(m["accuracy"]), list(m["dice"]), list(m["jaccard"])


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--folder", default="pix2obj", help="The folder path including params.txt")
    parser.add_argument("--synthetic", type=int, default=0, help="run on N synthetic pairs instead of opt.data")
    parser.add_argument("--batch", type=int, default=1, help="inference batch size (reference: 1)")
    args = parser.parse_args(argv)
    opt = load_opt(os.path.join(os.getcwd(), "models", args.folder.split("/")[-1], "params.txt"))
    device = torch.device("cuda:0")
    gen = load_model(os.path.join(os.getcwd(), "models", opt.folder_save, "final_model.pth"), opt, device)
    if args.synthetic > 0:
        from .train import SyntheticPairs
        data = SyntheticPairs(args.synthetic, getattr(opt, "image_size", 256), opt.input_dim, opt.output_dim)
    else:
        from .datasets.datasets import get_dataset
        data = get_dataset(os.path.join(os.getcwd(), opt.data, "test", "source"), opt, mode="test")
    loader = DataLoader(dataset=data, batch_size=args.batch, shuffle=False, num_workers=0, drop_last=False)
    output_path = os.path.join(os.getcwd(), "Outputs", opt.folder_save)
    mkdir(output_path)
    accuracy, dice, jaccard = test_model(gen, loader, output_path, evaluation=True, device=device)
    if len(accuracy) > 0:
        print_evaluation(accuracy, dice, jaccard, output_path)


if __name__ == "__main__":
    main()
