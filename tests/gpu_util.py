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


# ---------------------------------------------------------------------------------------------------------------
# bf16-storage emulation of the training step (torch autograd on the GPU): the same network as
# oracle/train_oracle.py with every tensor the CUDA path STORES (conv outputs, activations, their gradients, the GEMM
# weights) rounded to bf16 at the point where it is stored.  Separates kernel bugs from precision effects.
# ---------------------------------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundFwdOnly(torch.autograd.Function):
    """weights: the GEMM consumes a bf16 copy, the gradient goes to the fp32 master unrounded."""
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def emulated_forward_train(P, S, x, eps=1e-5, momentum=0.1, record=None, forced=None, dev_log=None):
    """record: optional dict that receives name + '.y' (raw conv output) / name + '.a' (stored activation), NCHW.
    forced: optional dict with the same keys holding the CUDA path's own stored tensors (NCHW fp32): every stored
    tensor is compared with what torch computes from the *forced* inputs (deviation appended to dev_log as
    (key, max |diff| in units of the bf16 spacing of max(|value|, 1))) and then replaced by the forced value
    (straight-through), so that every layer -- and the whole backward pass -- is evaluated at exactly the CUDA path's
    forward state: per-layer "teacher forcing", no compounding of rounding differences."""
    import torch.nn.functional as F
    rb, rw = _RoundBF16.apply, _RoundFwdOnly.apply

    def rec(key, t):
        if record is not None:
            record[key] = t.detach()
        if forced is not None:
            mine = forced[key]
            if dev_log is not None:
                ulp = torch.clamp(t.detach().abs(), min=1.0) * 2.0 ** -8
                dev_log.append((key, float(((mine - t.detach()).abs() / ulp).max())))
            t = t + (mine - t).detach()
        return t

    def conv(t, name, **kw):
        return rec(name + ".y", rb(F.conv2d(t, rw(P[name + ".weight"]), **kw)))

    def bn(t, name):
        return F.batch_norm(t, S[name + ".running_mean"], S[name + ".running_var"], P[name + ".weight"],
                            P[name + ".bias"], True, momentum, eps)

    x = x.to(torch.bfloat16).float()
    t = rec("conv1.a", rb(F.relu(bn(conv(x, "conv1", stride=2, padding=3), "bn1"))))
    t = F.max_pool2d(t, 3, 2, 1)
    for li, blocks in enumerate((3, 4, 6, 3), start=1):
        for b in range(blocks):
            p = "layer%d.%d" % (li, b)
            stride = 2 if (b == 0 and li > 1) else 1
            idt = t
            o = rec(p + ".conv1.a", rb(F.relu(bn(conv(t, p + ".conv1"), p + ".bn1"))))
            o = rec(p + ".conv2.a", rb(F.relu(bn(conv(o, p + ".conv2", stride=stride, padding=1), p + ".bn2"))))
            o = bn(conv(o, p + ".conv3"), p + ".bn3")
            if b == 0:
                idt = rec(p + ".downsample.0.a", rb(bn(conv(t, p + ".downsample.0", stride=stride),
                                                       p + ".downsample.1")))
            t = rec(p + ".conv3.a", rb(F.relu(o + idt)))
    feat = torch.flatten(F.adaptive_avg_pool2d(t, 1), 1)
    out = {}
    for head in ("fc", "fc_occ", "fc_depth"):
        if head + ".weight" in P:
            out[head] = F.linear(feat, P[head + ".weight"], P[head + ".bias"])
    return out


def stored_keys():
    """Keys of every tensor the training forward stores, in execution order."""
    keys = ["conv1.y", "conv1.a"]
    for li, blocks in enumerate((3, 4, 6, 3), start=1):
        for b in range(blocks):
            p = "layer%d.%d" % (li, b)
            keys += [p + ".conv1.y", p + ".conv1.a", p + ".conv2.y", p + ".conv2.a", p + ".conv3.y"]
            if b == 0:
                keys += [p + ".downsample.0.y", p + ".downsample.0.a"]
            keys += [p + ".conv3.a"]
    return keys
