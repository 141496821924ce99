"""The HBM-bound kernels of one level-1 UNet block in isolation (BASELINE.json config 5): GroupNorm(+SiLU) statistics and apply,
LayerNorm, temporal attention — for `ncu --set full` (achieved DRAM throughput) and CUDA-event GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
n, HW, C, T, heads = 28, 9216, 320, 14, 5
x = torch.randn(n * HW, C, device="cuda")
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
y = torch.empty(n * HW, C, device="cuda", dtype=AD)
st = torch.zeros(n * 64, device="cuda", dtype=torch.float64)
qkv = (torch.randn(n // T, T, HW, 3 * C, device="cuda") * 0.5).to(AD)
o = torch.empty(n // T, T, HW, C, device="cuda", dtype=AD)


def timed(name, fn, nbytes):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s", flush=True)


E = n * HW * C
timed("groupnorm stats+apply f32->act (+SiLU)", lambda: ops.groupnorm(x, n, HW, C, g, b, 1e-5, True, y, st), E * 10)
timed("groupnorm apply only (fused statistics)", lambda: ops.groupnorm(x, n, HW, C, g, b, 1e-5, True, y, st, have_stats=True), E * 6)
timed("layernorm f32->act", lambda: ops.layernorm(x, g, b, y), E * 6)
timed("temporal attention T=14", lambda: ops.attention_temporal(qkv, n // T, T, HW, heads, o), E * 8)
