"""Times the spatial attention kernel at the three UNet levels (CUDA events, L2 flushed by the 1.3 GB working set rotation).
GCD_FA_EMU=0|4|2 selects the share of exponentials evaluated on the FMA pipe (read once per process)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
torch.manual_seed(0)
res = {}
for (f, tok, h) in ((28, 9216, 5), (28, 2304, 10), (28, 576, 20)):
    qkvs = [(torch.randn(f, tok, 3 * h * 64, device="cuda") * 1.0).to(AD) for _ in range(3)]
    out = torch.empty(f, tok, h * 64, device="cuda", dtype=AD)
    for i in range(3):
        ops.attention_spatial(qkvs[i], f, tok, h, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 12
    e0.record()
    for i in range(n):
        ops.attention_spatial(qkvs[i % 3], f, tok, h, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * f * h * tok * tok * 64
    res[f"{f}x{tok}x{h}"] = (round(ms, 4), round(fl / ms / 1e9, 1))
    # correctness vs torch SDPA (fp32)
    q, k, v = qkvs[0].float().view(f, tok, 3, h, 64).permute(2, 0, 3, 1, 4)[:, :2].unbind(0)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(2, tok, h * 64)
    ops.attention_spatial(qkvs[0], f, tok, h, out)
    err = ((out[:2].float() - ref).norm() / ref.norm()).item()
    res[f"{f}x{tok}x{h}"] += (f"{err:.2e}",)
print("GCD_FA_EMU", os.environ.get("GCD_FA_EMU", "default"), res)
