"""Bring-up probe for the MN-major UMMA operand descriptors: runs one weight-gradient and one data-gradient case per
(LBO, SBO) candidate in a subprocess (the values are read once per process) and prints the error against torch."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from instaorder_b200 import _lib
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
def wg(B, H, W, Cin, Cout, k, stride):
    x = torch.randn((B, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    dy = torch.randn((B, H // stride, W // stride, Cout), generator=g, device="cuda").to(torch.bfloat16)
    dw = torch.zeros((Cout, k * k * Cin), device="cuda")
    _lib.check(L.io_conv_wgrad(x.data_ptr(), B, H, W, Cin, dy.data_ptr(), Cout, k, stride, dw.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, k, k), dy.float().permute(0, 3, 1, 2),
                                      stride=stride, padding=k // 2).permute(0, 2, 3, 1).reshape(Cout, -1)
    return float((dw - ref).abs().max() / ref.abs().max())
def dg(B, H, W, Cin, Cout, k):
    w = (torch.randn((Cout, Cin, k, k), generator=g, device="cuda") / (Cout * k * k) ** 0.5).to(torch.bfloat16)
    dy = torch.randn((B, H, W, Cout), generator=g, device="cuda").to(torch.bfloat16)
    dx = torch.zeros((B, H, W, Cin), device="cuda", dtype=torch.bfloat16)
    zero = torch.zeros(2048, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    _lib.check(L.io_conv_dgrad(dy.data_ptr(), B, H, W, Cin, Cout, k, wp.data_ptr(), zero.data_ptr(), None, dx.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), w.float(), dy.float().permute(0, 3, 1, 2), padding=k // 2).permute(0, 2, 3, 1)
    return float((dx.float() - ref).abs().max() / ref.abs().max())
print("wgrad 64->64 1x1 %%.4g | 256->512 1x1 %%.4g | 128 3x3 %%.4g | dgrad 64->64 %%.4g | 256->512 %%.4g | 128 3x3 %%.4g" %% (
    wg(2, 16, 16, 64, 64, 1, 1), wg(2, 16, 16, 256, 512, 1, 1), wg(2, 16, 16, 128, 128, 3, 1),
    dg(2, 16, 16, 64, 64, 1), dg(2, 16, 16, 256, 512, 1), dg(2, 16, 16, 128, 128, 3)))
''' % ROOT

if __name__ == "__main__":
    cands = [(8192, 1024), (1024, 8192), (8192, 2048), (128, 1024), (1024, 1024), (16, 1024)]
    for lbo, sbo in cands:
        env = dict(os.environ, INSTAORDER_MN_LBO=str(lbo), INSTAORDER_MN_SBO=str(sbo))
        try:
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
            out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
            if r.returncode != 0:
                out += " | rc=%d %s" % (r.returncode, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "")
        except subprocess.TimeoutExpired:
            out = "timeout"
        print("LBO %5d SBO %5d : %s" % (lbo, sbo, out), flush=True)
