"""PatchDiscriminator with the reference's constructor, parameter names and feature hooks
(reference: discriminators/PatchDiscriminator.py:5-43): five 3x3 / pad-0 convolutions with strides
2,2,1,1,1, InstanceNorm(affine)+LeakyReLU(0.2) on the middle three, optional Sigmoid. The torch.nn
layers are parameter containers; forward() runs the sm_100a engine (engine.PatchDInstance)."""
import torch
import torch.nn as nn

from .. import _C
from ..bridge import disc_forward


class PatchDiscriminator(nn.Module):
    def __init__(self, in_channels=3, out_channel=3, num_filter=64, return_filters=True, activation=True):
        super().__init__()
        self.return_filters = return_filters
        self.kw, self.padw = 3, 0
        nf = num_filter
        plan = [(in_channels + out_channel, nf, 2, False), (nf, nf * 2, 2, True), (nf * 2, nf * 4, 1, True),
                (nf * 4, nf * 8, 1, True)]
        layers = []
        for cin, cout, stride, norm in plan:
            layers.append(nn.Conv2d(cin, cout, kernel_size=self.kw, stride=stride, padding=self.padw, bias=not norm))
            if norm:
                layers.append(nn.InstanceNorm2d(cout, affine=True, track_running_stats=False))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
        layers.append(nn.Conv2d(nf * 8, 1, kernel_size=self.kw, stride=1, padding=self.padw))
        if activation:
            layers.append(nn.Sigmoid())
        self.model = nn.Sequential(*layers)
        self.intermediate_outputs = []

    def forward(self, img_A, img_B):
        """(B,1,h,w) prediction; the four LeakyReLU feature maps of this call are kept (detached) for
        get_intermediate_output(), like the reference's forward hooks."""
        pred, feats = disc_forward(self, img_A, img_B)
        self.intermediate_outputs = feats if self.return_filters else []
        return pred

    def get_intermediate_output(self):
        return self.intermediate_outputs[:4]
