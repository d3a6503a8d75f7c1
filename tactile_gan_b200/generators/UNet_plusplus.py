"""UNet++ generator with the reference's constructor and parameter names
(reference: generators/UNet_plusplus.py:5-86). The torch.nn layers are parameter containers only --
state_dict keys / shapes / init behaviour match the reference -- while forward() runs the sm_100a
engine (tactile_gan_b200/engine.py: tcgen05 implicit-GEMM convs, fused InstanceNorm/ReLU tail)."""
import torch.nn as nn

from ..bridge import EngineModule


class FeatureMapBlock(nn.Module):
    """1x1 projection (+Tanh when `activation`), reference UNet_plusplus.py:5-16."""

    def __init__(self, input_channels, output_channels, activation=True):
        super().__init__()
        self.conv = nn.Conv2d(input_channels, output_channels, kernel_size=1)
        self.activation = activation


class ConvBlock(nn.Module):
    """[conv3x3 -> InstanceNorm(affine) -> ReLU] x 2, reference UNet_plusplus.py:18-34."""

    def __init__(self, in_size, out_size, kernel=3, stride=1, padding=1):
        super().__init__()
        def stage(cin, k, s, p):
            return [nn.Conv2d(cin, out_size, kernel_size=k, stride=s, padding=p, bias=False),
                    nn.InstanceNorm2d(out_size, affine=True, track_running_stats=False), nn.ReLU(True)]

        self.layer = nn.Sequential(*stage(in_size, kernel, stride, padding), *stage(out_size, 3, 1, 1))


class UNet_plusplus(EngineModule):
    engine_kind = "unet++"

    def __init__(self, input_dim=3, output_dim=3, num_filter=64, activation=True):
        super().__init__()
        nf = num_filter
        width = [nf << i for i in range(5)]
        for j in range(5):
            for i in range(5 - j):
                if j == 0:
                    cin = input_dim if i == 0 else width[i - 1]
                else:
                    cin = j * width[i] + width[i + 1]   # j same-level skips + the upsampled deeper node
                setattr(self, f"conv{i}_{j}", ConvBlock(cin, width[i]))
        self.downfeature = FeatureMapBlock(nf, output_dim, activation=activation)
        self._reorder()

    def _reorder(self):
        """Register blocks in the reference's order (column by column, UNet_plusplus.py:43-63) so that
        parameters() -- and therefore optimizer state indices -- line up with reference checkpoints."""
        order = [f"conv{i}_0" for i in range(5)] + [f"conv{i}_1" for i in range(4)] + \
                [f"conv{i}_2" for i in range(3)] + [f"conv{i}_3" for i in range(2)] + ["conv0_4", "downfeature"]
        mods = self._modules
        self._modules = type(mods)((k, mods[k]) for k in order)
