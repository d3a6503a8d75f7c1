"""UNet generator with the reference's constructor and parameter names (reference: generators/UNet.py).
Parameter containers only -- forward() runs engine.UNetEngine (strided convs and the transposed convs,
as 4 sub-pixel phases, on the tcgen05 implicit-GEMM kernel)."""
import torch.nn as nn

from ..bridge import EngineModule
from .UNet_plusplus import FeatureMapBlock


def _stage(first, out_size):
    norm = lambda: nn.InstanceNorm2d(out_size, affine=True, track_running_stats=False)
    return nn.Sequential(first, norm(), nn.ReLU(True),
                         nn.Conv2d(out_size, out_size, kernel_size=3, stride=1, padding=1, bias=False), norm(),
                         nn.ReLU(True))


class ConvDown(nn.Module):
    """conv k4 s2 p1 -> IN -> ReLU -> conv 3x3 -> IN -> ReLU (reference UNet.py:17-33)."""

    def __init__(self, in_size, out_size, kernel=4, stride=2, padding=1):
        super().__init__()
        self.layer = _stage(nn.Conv2d(in_size, out_size, kernel_size=kernel, stride=stride, padding=padding,
                                      bias=False), out_size)


class DeconvUp(nn.Module):
    """convT k4 s2 p1 -> IN -> ReLU -> conv 3x3 -> IN -> ReLU (reference UNet.py:36-51)."""

    def __init__(self, in_size, out_size, kernel=4, stride=2, padding=1):
        super().__init__()
        self.layer = _stage(nn.ConvTranspose2d(in_size, out_size, kernel, stride, padding, bias=False), out_size)


class UNet(EngineModule):
    engine_kind = "unet"

    def __init__(self, input_dim=3, output_dim=3, num_filter=64, activation=True):
        super().__init__()
        nf = num_filter
        down = [input_dim, nf, nf * 2, nf * 4, nf * 8, nf * 8, nf * 8, nf * 8]
        for i in range(1, 8):
            setattr(self, f"conv{i}", ConvDown(down[i - 1], down[i]))
        # deconv_i consumes cat(d_{i-1}, c_{9-i}) (deconv2: c7 alone) and halves the depth of the pyramid
        up_out = {2: nf * 8, 3: nf * 8, 4: nf * 8, 5: nf * 4, 6: nf * 2, 7: nf, 8: nf}
        prev = nf * 8
        for i in range(2, 9):
            cin = prev if i == 2 else prev + down[9 - i]
            setattr(self, f"deconv{i}", DeconvUp(cin, up_out[i]))
            prev = up_out[i]
        self.downfeature = FeatureMapBlock(nf, output_dim, activation=activation)
