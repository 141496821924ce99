import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
M, K, N = 258048, 320, 960
mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
x, w = (torch.randn(M, K, device="cuda") * 0.5).to(AD), (torch.randn(N, K, device="cuda") * 0.05).to(AD)
if mode == "f16":
    out = torch.empty(M, N, device="cuda", dtype=AD)
    ep = ops.make_ep(out, bias=torch.zeros(N, device="cuda"))
else:
    N = 320
    w = w[:N].contiguous()
    out = torch.zeros(M, N, device="cuda")
    ep = ops.make_ep(out, bias=torch.zeros(N, device="cuda"), res1=out)
for _ in range(3):
    ops.linear(x, w, ep)
torch.cuda.synchronize()
