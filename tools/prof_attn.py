import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
f, tok, h = 28, 9216, 5
qkv = (torch.randn(f, tok, 3 * h * 64, device="cuda") * 0.5).to(AD)
out = torch.empty(f, tok, h * 64, device="cuda", dtype=AD)
for _ in range(3):
    ops.attention_spatial(qkv, f, tok, h, out)
torch.cuda.synchronize()
