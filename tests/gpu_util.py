"""Helpers shared by the GPU parity tests (everything goes through the C ABI via instaorder_b200._lib)."""
import numpy as np
import torch

from instaorder_b200 import _lib


def bf16_bits_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_rn(a):
    """numpy round-to-nearest-even fp32 -> bf16 -> fp32 (the rounding the kernels apply)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    return r.astype(np.uint32).view(np.float32).reshape(a.shape)


def unpack_pair_tensor(t_u8, P, D):
    """device uint8 pair tensor -> fp32 numpy [P, 5, D, D] (interior) + the raw [P, D+6, pitch, 8] fp32 view."""
    pitch = _lib.lib().io_pair_tensor_row_pitch(D)
    n = P * (D + 6) * pitch * 8
    raw = t_u8[: n * 2].view(torch.bfloat16).float().cpu().numpy().reshape(P, D + 6, pitch, 8)
    inner = raw[:, 3:3 + D, 3:3 + D, :5].transpose(0, 3, 1, 2)
    return np.ascontiguousarray(inner), raw


def conv_reference(x, w, scale, bias, residual, stride, relu):
    """fp32 torch reference of conv + folded BN (+ residual) (+ relu) on bf16-rounded operands.
    x: [B,H,W,Cin] bf16 cuda; w: [Cout,Cin,k,k] fp32 cuda (already scaled + bf16-rounded)."""
    import torch.nn.functional as F
    k = w.shape[-1]
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=stride, padding=k // 2)
    y = y + bias[None, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = torch.relu(y)
    return y


def pack_weight(w):
    """[Cout,Cin,k,k] fp32 -> [Cout, k*k*Cin] bf16 (tap-major, channel-minor)."""
    co, ci, k, _ = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, k * k * ci).to(torch.bfloat16).contiguous()
