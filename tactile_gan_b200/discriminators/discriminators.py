"""Discriminator factory with the reference's signature (reference: discriminators/discriminators.py:5-14)."""
from .PatchDiscriminator import PatchDiscriminator


def create_disc(name, in_nc, out_nc, num_filter, return_filter, activation=True, multigpu=False):
    """`multigpu` is accepted for signature parity (data parallelism = one process per GPU + NCCL)."""
    if name.lower() == "patch":
        # the first conv runs on im2col rows of 9 * (in_nc + out_nc) <= 64 channels; the image-gradient fold keeps
        # <= 8 target channels in registers (tg_col2im_grad)
        if 9 * (in_nc + out_nc) > 64 or out_nc > 8 or in_nc < 1 or out_nc < 1:
            raise ValueError(f"PatchDiscriminator on sm_100a needs in_nc + out_nc <= 7 (got {in_nc} + {out_nc})")
        return PatchDiscriminator(in_nc, out_nc, num_filter=num_filter, return_filters=return_filter,
                                  activation=activation)
    raise NameError(f"{name} not a valid model")
