"""Generator factory + GANLoss with the reference's signatures (reference: generators/generators.py)."""
import torch
import torch.nn as nn

from .UNet_plusplus import UNet_plusplus


def create_gen(name, in_nc, out_nc, num_filter, activation=True, multigpu=False):
    """Reference generators/generators.py:8-25. `multigpu` is accepted for signature parity; data
    parallelism is one process per GPU with NCCL gradient allreduce (tactile_gan_b200/step.py), not
    nn.DataParallel."""
    key = name.lower()
    if key == "unet":
        from .UNet import UNet
        return UNet(input_dim=in_nc, output_dim=out_nc, num_filter=num_filter, activation=activation)
    if key == "unet++":
        return UNet_plusplus(input_dim=in_nc, output_dim=out_nc, num_filter=num_filter, activation=activation)
    if key == "bcdunet":
        from .BCDUNet import BCDUNet
        return BCDUNet(input_dim=in_nc, output_dim=out_nc, num_filter=num_filter, bidirectional=True,
                       activation=activation)
    raise NameError(f"{name} not a valid model")
