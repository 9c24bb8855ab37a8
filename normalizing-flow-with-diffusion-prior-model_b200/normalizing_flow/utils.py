"""Coupling-network containers and the step glue of the reference's normalizing_flow/utils.py.

``ZeroConv2d`` / ``Conv2dActNorm`` / ``coupling_network`` exist so that parameter names, shapes, init and the
state_dict layout equal the reference's (utils.py:17-90): ``net.0._Conv2dActNorm__conv.weight`` etc.  In the
hot path their arithmetic is executed by the GEMM kernels through ``_engine.coupling_rows``; the modules are
parameter holders.  The aim / torchvision / PIL helpers of the reference file (track_images, save_images,
get_data_transforms) are outside the hot path and are not mirrored.
"""
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _engine as E
from . import _native as N

LatentDim = Tuple[int, int, int]


class ZeroConv2d(nn.Conv2d):
    """3x3 conv whose weight/bias start at zero, followed by a learned per-channel gain exp(3*logs)
    (reference utils.py:17-44)."""

    def __init__(self, in_channels: int, out_channels: int, filter_size: int = 3, stride: int = 1, padding: int = 0,
                 logscale: float = 3.):
        super().__init__(in_channels, out_channels, filter_size, stride=stride, padding=padding)
        with torch.no_grad():
            self.weight.zero_()
            self.bias.zero_()
        self.register_parameter("logs", nn.Parameter(torch.zeros(1, out_channels, 1, 1)))
        self.logscale_factor = logscale
        self._cache = E.SplitCache()

    def forward(self, x):
        """Stand-alone evaluation (off the hot path): im2col + fp32 GEMM + gain, NCHW in / NCHW out."""
        if E.autograd_needed(x, self):
            raise NotImplementedError("ZeroConv2d.forward: stand-alone autograd is not provided; use it inside "
                                      "AffineCoupling / Split or under torch.no_grad()")
        if self.kernel_size != (3, 3) or self.padding != (1, 1) or self.stride != (1, 1) or self.logscale_factor != 3.:
            raise ValueError("ZeroConv2d kernels support filter_size=3, stride=1, padding=1, logscale=3 only")
        x = E.check_input(x)
        B, Cin, H, W = x.shape
        Cout = self.out_channels
        h, ld = E.conv3x3_rows(self, x, Cin * H * W, B, Cin, Cout, H, W)
        out = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x.device)
        N.rows_to_nchw(h, ld, 1, self.bias, self.logs, out, B, Cout, H * W)
        return out


class Conv2dActNorm(nn.Module):
    """conv (no bias, "same" padding) followed by ActNorm (reference utils.py:47-69)."""

    def __init__(self, in_channels: int, out_channels: int, filter_size: int, stride: int = 1, padding: int = None):
        super().__init__()
        from .transforms import ActNorm
        padding = (filter_size - 1) // 2 or padding
        self.__conv = nn.Conv2d(in_channels, out_channels, filter_size, stride=stride, padding=padding, bias=False)
        self.__actnorm = ActNorm(out_channels)

    @property
    def conv(self) -> nn.Conv2d:
        return self.__conv

    @property
    def actnorm(self):
        return self.__actnorm

    def forward(self, x):
        """Stand-alone evaluation (off the hot path; inside AffineCoupling the fused GEMM epilogue is used):
        conv as an exact-fp32 GEMM over pixel rows, data-dependent ActNorm init on first use, NCHW out."""
        if E.autograd_needed(x, self):
            raise NotImplementedError("Conv2dActNorm.forward: stand-alone autograd is not provided; use it inside "
                                      "AffineCoupling or under torch.no_grad()")
        conv, an = self.__conv, self.__actnorm
        x = E.check_input(x)
        B, Cin, H, W = x.shape
        Cout, P = conv.out_channels, H * W
        if conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride == (1, 1):
            h, ld = E.conv3x3_rows(conv, x, Cin * P, B, Cin, Cout, H, W)
        elif conv.kernel_size == (1, 1) and conv.stride == (1, 1):
            Kp, ld = E.round_up(Cin, 16), E.round_up(Cout, 8)
            wp = torch.empty(ld * Kp, dtype=torch.float32, device=x.device)
            N.pack_matrix(conv.weight, wp, 1, Cout, Cin, 0, Cin, 1, Kp, ld)
            A = E.WS.get("As", B * P * Kp, torch.float32, x.device)
            N.nchw_to_rows(x, A, B, Cin, P, Cin * P, Kp)
            h = E.WS.get("hs", B * P * ld, torch.float32, x.device)
            N.gemm_nt(A, Kp, wp, Kp, h, ld, B * P, Cout, Kp)
        else:
            raise ValueError("Conv2dActNorm kernels support 3x3 (pad 1) and 1x1 convolutions with stride 1")
        if not an._initialized():
            E.channel_stats(h, 1, B, Cout, P, ld, an.scale, an.bias)
            an._mark_initialized()
        out = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x.device)
        N.rows_to_nchw(h, ld, 2, an.scale, an.bias, out, B, Cout, P)
        return out


def coupling_network(in_channels: int, n_features: int = 512, out_channels: int = None) -> nn.Sequential:
    """conv3x3+ActNorm, ReLU, conv1x1+ActNorm, ReLU, ZeroConv3x3 (reference utils.py:72-90); indices 0/2/4 hold
    the parameters."""
    return nn.Sequential(
        Conv2dActNorm(in_channels, n_features, 3, padding=1),
        nn.ReLU(inplace=True),
        Conv2dActNorm(n_features, n_features, 1, padding=0),
        nn.ReLU(inplace=True),
        ZeroConv2d(n_features, out_channels or in_channels, padding=1),
    )


def calculate_output_shapes(L: int, in_channels: int, size: int) -> List[LatentDim]:
    """Latent shapes per level, e.g. (3, 3, 32) -> [(6,16,16), (12,8,8), (48,4,4)] (reference utils.py:93-117)."""
    shapes = []
    c, s = in_channels, size
    for _ in range(L - 1):
        if s % 2 != 0:
            raise ValueError("The input dimension is not divisible by 2!")
        c, s = c * 2, s // 2
        shapes.append((c, s, s))
    shapes.append((c * 4, s // 2, s // 2))
    return shapes


def init_optimizer(name: str, params, lr: float) -> torch.optim.Optimizer:
    """reference utils.py:120-137."""
    if name == "adam":
        cls = torch.optim.Adam
    elif name == "adamw":
        cls = torch.optim.AdamW
    else:
        raise ValueError("Unknown optimizer")
    return cls(params, lr=lr if lr else params[0]["lr"])


def _kernel_ok(t: torch.Tensor) -> bool:
    return t.is_cuda and t.dtype == torch.float32 and t.numel() > 0


@torch.no_grad()
def preprocess_batch(batch: torch.Tensor, n_bits: int, n_bins: int, noise: torch.Tensor = None) -> torch.Tensor:
    """[0,1] images -> n_bits levels centred on 0 (reference utils.py:175-196).  CUDA tensors take one
    ``nfdpm_preprocess`` launch, bit-identical to the reference's op sequence; with ``noise`` (U[0,1), same shape) the
    dequantisation add of the training step (trainer.py:155, ``batch + rand_like(batch) / n_bins``) is fused into the
    same pass.  Host tensors (the reference also calls this before ``.to(device)`` in places) keep the torch expression."""
    if _kernel_ok(batch) and 1 <= n_bits <= 8 and (noise is None or (_kernel_ok(noise) and noise.shape == batch.shape)):
        x = batch.contiguous()
        out = torch.empty_like(x)
        N.preprocess(x, None if noise is None else noise.contiguous(), out, n_bits, n_bins)
        return out
    out = batch * 255
    if n_bits < 8:
        out = torch.floor(out / 2 ** (8 - n_bits))
    out = out / n_bins - 0.5
    return out if noise is None else out + noise / n_bins


@torch.no_grad()
def postprocess_batch(batch: torch.Tensor, n_bins: int) -> torch.Tensor:
    """Model space -> uint8 pixels on the CPU (reference utils.py:199-210).  CUDA tensors are quantised on the device
    (``nfdpm_postprocess_u8``) and only the uint8 image crosses PCIe."""
    if _kernel_ok(batch):
        x = batch.contiguous()
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        N.postprocess_u8(x, out, n_bins)
        return out.to("cpu")
    return torch.clip(torch.floor((batch + 0.5) * n_bins) * (256.0 / n_bins), 0, 255).to("cpu", torch.uint8)


def calculate_loss(log_likelihood: torch.Tensor, n_bins: float, n_pixel: float):
    """Mean bits per dimension (reference utils.py:244-256)."""
    return ((np.log(n_bins) * n_pixel - log_likelihood) * (np.log2(np.e) / n_pixel)).mean(dim=0)


def initialize_with_zeros(n: int, batch_size: int, device: torch.device):
    """n fp64 zero accumulators of length batch_size (reference utils.py:259-272)."""
    if n == 1:
        return torch.zeros(batch_size, device=device, dtype=torch.float64)
    return (torch.zeros(batch_size, device=device, dtype=torch.float64) for _ in range(n))


@torch.no_grad()
def data_dependent_nf_initialization(flow, dataloader, device: torch.device, n_bits: int, n_bins: int) -> None:
    """One forward pass over the first batch so every ActNorm initialises itself (reference utils.py:275-292)."""
    flow.eval()
    sample = next(iter(dataloader))
    batch = sample[0].to(device) if isinstance(sample, list) else sample.to(device)
    batch = preprocess_batch(batch, n_bits, n_bins, noise=torch.rand_like(batch))
    ll, lp = initialize_with_zeros(2, batch.size(0), device)
    flow.transform(batch, ll, lp)


def get_item(sequence: Sequence, index: int):
    """sequence[index] or None when out of range (reference utils.py:295-300)."""
    try:
        return sequence[index]
    except IndexError:
        return None
