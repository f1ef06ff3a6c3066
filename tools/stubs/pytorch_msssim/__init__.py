"""Stand-in for the `pytorch_msssim` package the reference imports at module scope (LSSVC_net.py:4,
dmc_net.py:4) but never uses on the coding path.  Parameter-free, so state_dict layouts are unchanged."""
import torch


class MS_SSIM(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        raise NotImplementedError("MS-SSIM is not part of the coding path")


def ms_ssim(*args, **kwargs):
    raise NotImplementedError("MS-SSIM is not part of the coding path")
