"""Generator factory + GANLoss with the reference's signatures (reference: generators/generators.py)."""
import torch
import torch.nn as nn

from .UNet_plusplus import UNet_plusplus


def create_gen(name, in_nc, out_nc, num_filter, activation=True, multigpu=False):
    """Reference generators/generators.py:8-25. `multigpu` is accepted for signature parity; data
    parallelism is one process per GPU with NCCL gradient allreduce (tactile_gan_b200/step.py), not
    nn.DataParallel."""
    key = name.lower()
    if key in ("unet", "unet++", "bcdunet"):
        # limits of the sm_100a engines the reference does not have (fail here, not deep inside an engine build)
        if not 1 <= num_filter <= 64:
            raise ValueError(f"num_filter must be in 1..64 for the sm_100a generators (got {num_filter}): the "
                             f"FeatureMapBlock kernel reads one 64-channel row")
        if not 1 <= out_nc <= 4 or not 1 <= in_nc <= 64:
            raise ValueError(f"supported channel counts: input 1..64, output 1..4 (got {in_nc} -> {out_nc})")
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


class GANLoss(nn.Module):
    """Reference generators/generators.py:27-121: ls / ce / w / hinge objectives with a cached
    (optionally smoothed, CPU-RNG drawn) real-label tensor. Used by code written against the reference;
    the fused training iteration evaluates the same formulas in tg_gan_loss. Label tensors follow the
    input's device (the reference hard-codes 'cuda')."""

    def __init__(self, gan_mode='hinge', label_smoothing=False, target_real_label=1.0, target_fake_label=0.0,
                 tensor=torch.FloatTensor):
        super().__init__()
        if gan_mode not in ('ls', 'ce', 'w', 'hinge'):
            raise ValueError(f'Unexpected gan mode {gan_mode}')
        self.label_smoothing = label_smoothing
        self.real_label, self.fake_label = target_real_label, target_fake_label
        self.real_label_tensor = self.fake_label_tensor = self.zero_tensor = None
        self.Tensor = tensor
        self.gan_mode = gan_mode

    def get_target_tensor(self, input, target_is_real):
        if target_is_real:
            if self.real_label_tensor is None:
                if self.label_smoothing:
                    t = torch.clamp(torch.normal(self.real_label, .02, size=input.size()), 0, 1)
                else:
                    t = torch.tensor([self.real_label], dtype=torch.float32)
                self.real_label_tensor = t.to(input.device).requires_grad_(False)
            return self.real_label_tensor.expand_as(input)
        if self.fake_label_tensor is None:
            self.fake_label_tensor = torch.tensor([self.fake_label], dtype=torch.float32, device=input.device)
        return self.fake_label_tensor.expand_as(input)

    def get_zero_tensor(self, input):
        if self.zero_tensor is None:
            self.zero_tensor = torch.tensor([0], dtype=torch.float32, device=input.device)
        return self.zero_tensor.expand_as(input)

    def loss(self, input, target_is_real, for_discriminator=True):
        if self.gan_mode == 'ce':
            return nn.functional.binary_cross_entropy_with_logits(input, self.get_target_tensor(input, target_is_real))
        if self.gan_mode == 'ls':
            return nn.functional.mse_loss(input, self.get_target_tensor(input, target_is_real))
        if self.gan_mode == 'hinge':
            if for_discriminator:
                margin = (input - 1) if target_is_real else (-input - 1)
                return -torch.mean(torch.min(margin, self.get_zero_tensor(input)))
            assert target_is_real, "The generator's hinge loss must be aiming for real"
            return -torch.mean(input)
        return -input.mean() if target_is_real else input.mean()

    def __call__(self, input, target_is_real, for_discriminator=True):
        if isinstance(input, list):   # multiscale discriminator outputs
            total = 0
            for pred in input:
                if isinstance(pred, list):
                    pred = pred[-1]
                lt = self.loss(pred, target_is_real, for_discriminator)
                bs = 1 if len(lt.size()) == 0 else lt.size(0)
                total = total + torch.mean(lt.view(bs, -1), dim=1)
            return total / len(input)
        return self.loss(input, target_is_real, for_discriminator)
