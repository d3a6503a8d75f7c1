"""Drop-in for the reference's two_step_test.py: two trained generators chained, `gen2(gen1(x))` (two_step_test.py:
21-23), same flags (--s1_dir, --s2_dir, --data), same Outputs/<s1>+<s2>_<data>/ layout and eval.txt. Both forwards run
on the sm_100a engines (CUDA-graph replay per batch shape); the fuzzy metrics are reduced on the device. Like
tactile_gan_b200.test the PNG montages (`sgt/`, `elm/`) are not produced. Additive flags: --synthetic N, --batch."""
import argparse
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

from .test import fuzzy_sums, load_model, load_opt, metrics_from_sums, print_evaluation
from .util import mkdir


class _Chain(torch.nn.Module):
    """gen2 o gen1 as one module, so test_model-style loops take it like a single generator."""

    def __init__(self, gen1, gen2):
        super().__init__()
        self.gen1, self.gen2 = gen1, gen2

    def forward(self, x):
        return self.gen2(self.gen1(x))


def test_two_step(gen1, gen2, dataset, output_path, evaluation=True, device=None):
    """reference two_step_test.py:5-44 -> (accuracy, dice, jaccard) lists; writes out/<i>.npy (+ .png with PIL)."""
    device = device or next(gen1.parameters()).device
    mkdir(os.path.join(output_path, "out"))
    try:
        from PIL import Image
    except Exception:
        Image = None
    chain = _Chain(gen1, gen2)
    idx, sums = 0, []
    for batch in dataset:
        real_A, real_B = batch[0], batch[1]
        with torch.no_grad():
            out_dev = chain(real_A.to(device).float().contiguous())
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
    parser.add_argument("--s1_dir", default="t1_2d_per")
    parser.add_argument("--s2_dir", default="t2_2d_per")
    parser.add_argument("--data", default="data_plot_3")
    parser.add_argument("--synthetic", type=int, default=0, help="run on N synthetic pairs instead of --data")
    parser.add_argument("--batch", type=int, default=1, help="inference batch size (reference: 1)")
    args = parser.parse_args(argv)
    opt1 = load_opt(os.path.join(os.getcwd(), "models", args.s1_dir.split("/")[-1], "params.txt"))
    opt2 = load_opt(os.path.join(os.getcwd(), "models", args.s2_dir.split("/")[-1], "params.txt"))
    device = torch.device("cuda:0")
    gen1 = load_model(os.path.join(os.getcwd(), "models", opt1.folder_save, "final_model.pth"), opt1, device)
    gen2 = load_model(os.path.join(os.getcwd(), "models", opt2.folder_save, "final_model.pth"), opt2, device)
    if args.synthetic > 0:
        from .train import SyntheticPairs
        data = SyntheticPairs(args.synthetic, getattr(opt2, "image_size", 256), opt1.input_dim, opt2.output_dim)
    else:
        from .datasets.datasets import get_dataset
        data = get_dataset(os.path.join(os.getcwd(), args.data, "test", "source"), opt2, mode="test")
    loader = DataLoader(dataset=data, batch_size=args.batch, shuffle=False, num_workers=0, drop_last=False)
    output_path = os.path.join(os.getcwd(), "Outputs", f"{args.s1_dir}+{args.s2_dir}_{args.data}")
    mkdir(output_path)
    accuracy, dice, jaccard = test_two_step(gen1, gen2, loader, output_path, evaluation=True, device=device)
    if len(accuracy) > 0:
        print_evaluation(accuracy, dice, jaccard, output_path)


if __name__ == "__main__":
    main()
