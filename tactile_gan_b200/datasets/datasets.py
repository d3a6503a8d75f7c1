"""Dataset factory with the reference's signature (reference: datasets/datasets.py:3-6).

Same directory convention as the reference's PairedDataset (source/s_*.png -> tactile/t_*.tiff, or the three
t_*_axes / _grids / _content grayscale masks for --target ch; PairedDataset.py:61-78). Decoding stays on the
CPU workers (PIL); everything after it does not: with `raw=True` (what train.py uses) a sample is the pair of
uint8 HWC arrays, the batch crosses PCIe as uint8 and tactile_gan_b200.augment turns it into the normalised fp32
NCHW tensors on the device -- flip / affine included when `aug` is set (PairedDataset.py:30-44,80-92 ran
albumentations per sample in the workers; that library is unpinned and not installed here, see augment.py).
With `raw=False` the sample is what the reference returns without augmentation (fp32 CHW, source in [-1,1])."""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

IMG_EXT = ('.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.svg', '.tiff')


class PairedDataset(Dataset):
    def __init__(self, img_dir, size=256, mode='train', aug=False, target='rgb', raw=False):
        try:
            from PIL import Image  # noqa: F401
        except Exception as e:  # pragma: no cover
            raise ImportError("PairedDataset needs PIL; use --synthetic N for dataset-free runs") from e
        self.img_dir, self.size, self.mode, self.aug, self.target, self.raw = img_dir, size, mode, aug, target, raw
        images = []
        for root, _, fnames in sorted(os.walk(img_dir)):        # PairedDataset.py:21-27
            for fname in fnames:
                if fname.lower().endswith(IMG_EXT):
                    images.append(os.path.join(root, fname))
        self.images = images
        if aug and not raw:
            raise NotImplementedError("CPU-side augmentation (albumentations) is not rebuilt: use raw=True and "
                                      "tactile_gan_b200.augment on the device")

    def __len__(self):
        return len(self.images)

    def _load(self, i):
        from PIL import Image
        src = np.asarray(Image.open(self.images[i]).convert('RGB'))
        base, ext = self.images[i].replace("source", "tactile").replace("s_", "t_").replace(".png", ".tiff").rsplit(".", 1)
        if self.target == 'rgb':
            tgt = np.asarray(Image.open(f"{base}.{ext}").convert('RGB'))
        else:                                                    # PairedDataset.py:71-76
            tgt = np.stack([np.asarray(Image.open(f"{base}_{part}.{ext}").convert(mode="L"))
                            for part in ("axes", "grids", "content")], 2)
        return src, tgt

    def __getitem__(self, i):
        src, tgt = self._load(i)
        if self.raw:
            return torch.from_numpy(np.ascontiguousarray(src)), torch.from_numpy(np.ascontiguousarray(tgt))
        a = torch.from_numpy(src.astype(np.float32) / 255.).permute(2, 0, 1)
        b = torch.from_numpy(tgt.astype(np.float32) / 255.).permute(2, 0, 1)
        return (a - 0.5) / 0.5, b          # Normalize(.5,.5) on the source only (PairedDataset.py:52-58,86)


def get_dataset(img_dir, opt, mode='train', raw=False):
    """reference datasets/datasets.py:3-6 (+ raw: uint8 samples for the device-side pipeline)."""
    return PairedDataset(img_dir, mode=mode, aug=not getattr(opt, "no_aug", True) if raw else False,
                         target=getattr(opt, "target", "rgb"), raw=raw)
