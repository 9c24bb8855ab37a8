"""Host-side encode / decode of the split bf16 pair layout (NFDPM_BF16X2, include/nfdpm_b200.h) for the tests:
a logical row of ld columns is 2*ld bf16, group g = k // 32 holding [hi(32g..32g+31) | lo(32g..32g+31)];
torch carries the 4-byte-per-logical-element words as int32 (normalizing_flow._native.SPLIT)."""
import torch


def encode(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, ld] (ld % 32 == 0) -> int32 [rows, ld] holding hi = bf16(x), lo = bf16(x - hi)."""
    rows, ld = x.shape
    assert ld % 32 == 0
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    w = torch.stack([hi.reshape(rows, ld // 32, 32), lo.reshape(rows, ld // 32, 32)], dim=2).contiguous()
    return w.view(torch.int32).reshape(rows, ld)


def decode(t: torch.Tensor, rows: int, ld: int):
    """int32 [rows*ld] / [rows, ld] -> (hi, lo) fp32 [rows, ld]."""
    w = t.reshape(rows, ld).view(torch.bfloat16).reshape(rows, ld // 32, 2, 32).float()
    return w[:, :, 0].reshape(rows, ld), w[:, :, 1].reshape(rows, ld)


def value(t: torch.Tensor, rows: int, ld: int) -> torch.Tensor:
    hi, lo = decode(t, rows, ld)
    return hi.double() + lo.double()
