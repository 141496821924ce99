import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
n, H, W = 28, 72, 128
for (C, Co) in ((320, 320), (640, 640), (320, 256), (320, 512)):
    M = n * H * W
    x = (torch.randn(n, H, W, C, device="cuda") * 0.5).to(AD)
    w9 = (torch.randn(Co, 9 * C, device="cuda") * 0.02).to(AD)
    out = torch.empty(M, Co, device="cuda", dtype=AD)
    ep = ops.make_ep(out, bias=torch.zeros(Co, device="cuda"))
    ms = timeit(lambda: ops.conv2d_3x3(x, w9, ep)); fl = 2.0 * M * C * Co * 9
    print(f"conv3x3 {C}->{Co}@72x128: {ms:.3f} ms {fl/ms/1e9:.0f} TF", flush=True)
    xg = (torch.randn(M, 9 * C, device="cuda") * 0.5).to(AD)
    ms = timeit(lambda: ops.linear(xg, w9, ep))
    print(f"linear  M{M} K{9*C} N{Co}: {ms:.3f} ms {fl/ms/1e9:.0f} TF", flush=True)
    del xg
