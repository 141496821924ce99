import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
M, K = 258048, 320
mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
x = (torch.randn(M, K, device="cuda") * 0.5).to(AD)
if mode == "f16":
    N = 960
    w = (torch.randn(N, K, device="cuda") * 0.05).to(AD)
    out = torch.empty(M, N, device="cuda", dtype=AD)
    ep = ops.make_ep(out, bias=torch.zeros(N, device="cuda"))
elif mode == "geglu":
    N = 2560
    w = (torch.randn(N, K, device="cuda") * 0.05).to(AD)
    out = torch.empty(M, N // 2, device="cuda", dtype=AD)
    ep = ops.make_ep(out, bias=torch.zeros(N, device="cuda"), geglu=True)
else:
    N = 320
    w = (torch.randn(N, K, device="cuda") * 0.05).to(AD)
    out = torch.zeros(M, N, device="cuda")
    ep = ops.make_ep(out, bias=torch.zeros(N, device="cuda"), res1=out)
for _ in range(3):
    ops.linear(x, w, ep)
torch.cuda.synchronize()
