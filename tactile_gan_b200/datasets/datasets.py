"""Dataset factory with the reference's signature (reference: datasets/datasets.py:3-6).

The reference's PairedDataset depends on PIL + albumentations (third-party, unpinned, not installed in
this image); the augmentation pipeline is out of scope. This loader reads the same
directory convention (source/s_*.png, tactile/t_*.tiff) without augmentation when PIL is importable."""
import os

import torch
from torch.utils.data import Dataset


class PairedDataset(Dataset):
    def __init__(self, img_dir, size=256, mode='train', aug=False, target='rgb'):
        try:
            from PIL import Image  # noqa: F401
        except Exception as e:  # pragma: no cover
            raise ImportError("PairedDataset needs PIL; use --synthetic N for dataset-free runs") from e
        if target != 'rgb':
            raise NotImplementedError("--target ch (three grayscale masks) is not rebuilt; see SURVEY 8f")
        self.files = sorted(os.path.join(img_dir, f) for f in os.listdir(img_dir) if f.startswith("s_"))

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        import numpy as np
        from PIL import Image
        src = self.files[i]
        tgt = src.replace("source", "tactile").replace("s_", "t_").replace(".png", ".tiff")
        a = torch.from_numpy(np.asarray(Image.open(src).convert("RGB"), dtype=np.float32) / 255.).permute(2, 0, 1)
        b = torch.from_numpy(np.asarray(Image.open(tgt).convert("RGB"), dtype=np.float32) / 255.).permute(2, 0, 1)
        return (a - 0.5) / 0.5, b          # Normalize(.5,.5) on the source only (PairedDataset.py:52-58,86)


def get_dataset(img_dir, opt, mode='train'):
    return PairedDataset(img_dir, size=256, mode=mode, aug=not getattr(opt, "no_aug", True),
                         target=getattr(opt, "target", "rgb"))
