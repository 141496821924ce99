"""HBM-bound fp32-residual GEMMs of the UNet (to_out / FF2 projections, level 1-2): time and effective bandwidth."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
for (M, N, K) in ((258048, 320, 320), (258048, 320, 1280), (64512, 640, 640), (64512, 640, 2560), (16128, 1280, 1280)):
    xs = [(torch.randn(M, K, device="cuda") * 0.5).to(AD) for _ in range(2)]
    w = (torch.randn(N, K, device="cuda") * 0.03).to(AD)
    outs = [torch.zeros(M, N, device="cuda") for _ in range(2)]
    b = torch.zeros(N, device="cuda")
    for i in range(4):
        ops.linear(xs[i % 2], w, ops.make_ep(outs[i % 2], bias=b, res1=outs[i % 2]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        ops.linear(xs[i % 2], w, ops.make_ep(outs[i % 2], bias=b, res1=outs[i % 2]))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    by = M * K * 2 + 2 * M * N * 4
    print(f"M{M} N{N} K{K} f32 residual in/out: {ms:.4f} ms  {by / ms / 1e6:.0f} GB/s  {2.0 * M * N * K / ms / 1e9:.0f} TF", flush=True)
