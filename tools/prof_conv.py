import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
n, H, W, C, Co = 28, 72, 128, 320, 320
x = (torch.randn(n, H, W, C, device="cuda") * 0.5).to(AD)
w = (torch.randn(Co, 9 * C, device="cuda") * 0.02).to(AD)
out = torch.zeros(n * H * W, Co, device="cuda")
st = torch.zeros(n * 64, device="cuda", dtype=torch.float64)
ep = ops.make_ep(out, bias=torch.zeros(Co, device="cuda"), res1=out, gn_stats=(st, 10, 32, H * W))
for _ in range(3):
    ops.conv2d_3x3(x, w, ep)
torch.cuda.synchronize()
