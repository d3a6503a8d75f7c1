"""BCDUNet generator with the reference's constructor and parameter names (reference:
generators/BCDUNet.py). The reference constructs ConvLSTM / ConvBLSTM skip modules but never calls them
in forward (BCDUNet.py:154-181): inside BCDUNet they are parameter holders (30 state_dict keys,
Xavier-initialised peepholes). Called on their own they run the device engine of
tactile_gan_b200/convlstm.py (two-source implicit GEMM + fused gate kernel, forward and -- under autograd -- backward;
no eager fallback).
BCDUNet.forward() runs engine.BCDUNetEngine."""
import numpy as np
import torch
import torch.nn as nn

from ..bridge import EngineModule


class ConvLSTMCell(nn.Module):
    """Peephole ConvLSTM cell (reference BCDUNet.py:6-47): gates from one conv over cat[X, H]."""

    def __init__(self, in_channels, out_channels, kernel_size, padding, activation, frame_size):
        super().__init__()
        self.activation = {"tanh": torch.tanh, "relu": torch.relu}[activation]
        self.conv = nn.Conv2d(in_channels + out_channels, 4 * out_channels, kernel_size=kernel_size, padding=padding)
        for name in ("W_ci", "W_co", "W_cf"):
            w = nn.Parameter(torch.empty(out_channels, *frame_size))
            nn.init.xavier_uniform_(w)
            setattr(self, name, w)

    def forward(self, X, H_prev, C_prev):
        from ..convlstm import cell_forward
        return cell_forward(self, X, H_prev, C_prev)


class ConvLSTM(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, padding, activation, frame_size, return_sequence=False):
        super().__init__()
        self.out_channels, self.return_sequence = out_channels, return_sequence
        self.convLSTMcell = ConvLSTMCell(in_channels, out_channels, kernel_size, padding, activation, frame_size)

    def forward(self, X):
        """X: (batch, seq_len, channels, height, width) -> (batch, seq_len, out_channels, H, W), or the last frame's
        H when return_sequence is False (reference BCDUNet.py:61-84)."""
        from ..convlstm import lstm_forward
        return lstm_forward(self, X)


class ConvBLSTM(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, padding, activation, frame_size, return_sequence=False):
        super().__init__()
        self.return_sequence = return_sequence
        args = (in_channels, out_channels // 2, kernel_size, padding, activation, frame_size)
        self.forward_cell = ConvLSTM(*args, return_sequence=True)
        self.backward_cell = ConvLSTM(*args, return_sequence=True)

    def forward(self, x):
        from ..convlstm import blstm_forward
        return blstm_forward(self, x)


class BCDUNet(EngineModule):
    engine_kind = "bcdunet"

    def __init__(self, input_dim=3, output_dim=3, num_filter=64, frame_size=(256, 256), bidirectional=False,
                 activation=True, norm='instance'):
        super().__init__()
        if norm != 'instance':
            raise NotImplementedError("only the InstanceNorm variant (the one create_gen builds) has an engine")
        nf = self.num_filter = num_filter
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.dropout = nn.Dropout(0.5)          # constructed, never applied (reference quirk)
        self.frame_size = np.array(frame_size)
        self.activation = activation

        def conv_block(cin, cout):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.InstanceNorm2d(cout),
                                 nn.ReLU(inplace=True),
                                 nn.Conv2d(cout, cout, kernel_size=3, stride=1, padding=1), nn.InstanceNorm2d(cout),
                                 nn.ReLU(inplace=True))

        widths = [nf, nf * 2, nf * 4, nf * 8]
        cin = input_dim
        for i, wdt in enumerate(widths, 1):
            setattr(self, f"conv{i}", conv_block(cin, wdt))
            cin = wdt
        for k in (3, 2, 1):
            setattr(self, f"upconv{k}", nn.ConvTranspose2d(widths[k], widths[k - 1], kernel_size=2, stride=2))
        for k in (3, 2, 1):
            setattr(self, f"conv{k}m", conv_block(widths[k], widths[k - 1]))
        self.conv0 = nn.Conv2d(nf, output_dim, kernel_size=1)
        lstm = ConvBLSTM if bidirectional else ConvLSTM
        fs = self.frame_size
        self.clstm1 = lstm(nf * 4, nf * 2, (3, 3), (1, 1), 'tanh', list(fs // 4))
        self.clstm2 = lstm(nf * 2, nf, (3, 3), (1, 1), 'tanh', list(fs // 2))
        self.clstm3 = lstm(nf, nf // 2, (3, 3), (1, 1), 'tanh', list(fs))
