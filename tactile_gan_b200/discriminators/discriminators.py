"""Discriminator factory with the reference's signature (reference: discriminators/discriminators.py:5-14)."""
from .PatchDiscriminator import PatchDiscriminator


def create_disc(name, in_nc, out_nc, num_filter, return_filter, activation=True, multigpu=False):
    """`multigpu` is accepted for signature parity (data parallelism = one process per GPU + NCCL)."""
    if name.lower() == "patch":
        return PatchDiscriminator(in_nc, out_nc, num_filter=num_filter, return_filters=return_filter,
                                  activation=activation)
    raise NameError(f"{name} not a valid model")
