"""Drop-in for the inference part of the reference's test.py (`--folder`, params.txt + final_model.pth):
load_opt / load_model keep the reference signatures (test.py:24-42) and the generator forward
(test.py:202-203) runs on the sm_100a engine, batched. The CPU post-processing of the reference (fuzzy
accuracy / Dice / Jaccard, matplotlib / seaborn plots, PNG montage) is out of scope:
outputs are written as .npy (and PNG when PIL is available). Additive flags: --synthetic N, --batch."""
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


def test_model(model, dataset, output_path, evaluation=False, device=None):
    """Generator forward over the dataset (any batch size); saves out/<i>.npy (+ .png with PIL)."""
    device = device or next(model.parameters()).device
    mkdir(os.path.join(output_path, "out"))
    try:
        from PIL import Image
    except Exception:
        Image = None
    idx, l1 = 0, []
    for batch in dataset:
        real_A, real_B = batch[0], batch[1]
        with torch.no_grad():
            out = model(real_A.to(device).float().contiguous()).cpu()
        for j in range(out.shape[0]):
            idx += 1
            np.save(os.path.join(output_path, "out", f"{idx}.npy"), out[j].numpy())
            if Image is not None and out.shape[1] == 3:
                img = (out[j].clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype(np.uint8)
                Image.fromarray(img).save(os.path.join(output_path, "out", f"{idx}.png"))
            if evaluation:
                l1.append(float((out[j] - real_B[j]).abs().mean()))
    return l1


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
    l1 = test_model(gen, loader, output_path, evaluation=True, device=device)
    if l1:
        print(f"mean |out - target| over {len(l1)} samples: {np.mean(l1):.5f}")


if __name__ == "__main__":
    main()
