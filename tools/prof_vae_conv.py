"""The full-resolution VAE decoder convolutions (14 frames x 576 x 1024 x 128 channels): spatial 3x3 and temporal (3,1,1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
T, H, W, C = 14, 576, 1024, 128
which = sys.argv[1] if len(sys.argv) > 1 else "all"
x = (torch.randn(T, H, W, C, device="cuda") * 0.5).to(AD)
w9 = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(AD)
w3 = (torch.randn(C, 3 * C, device="cuda") * 0.02).to(AD)
o16 = torch.empty(T * H * W, C, device="cuda", dtype=AD)
o32 = torch.zeros(T * H * W, C, device="cuda")
b = torch.zeros(C, device="cuda")
st = torch.zeros(64 * 64, device="cuda", dtype=torch.float64)
cases = {
    "s3x3_f16": lambda: ops.conv2d_3x3(x, w9, ops.make_ep(o16, bias=b)),
    "s3x3_res": lambda: ops.conv2d_3x3(x, w9, ops.make_ep(o32, bias=b, res1=o32)),
    "t3_f16": lambda: ops.conv_t3(x.view(1, T, H * W, C), w3, ops.make_ep(o16, bias=b)),
    "s3x3_f16_gn14": lambda: ops.conv2d_3x3(x, w9, ops.make_ep(o16, bias=b, gn_stats=(st.zero_(), 4, 32, H * W))),
    "s3x3_res_gn1": lambda: ops.conv2d_3x3(x, w9, ops.make_ep(o32, bias=b, res1=o32, gn_stats=(st.zero_(), 4, 32, T * H * W))),
    "t3_f16_gn1": lambda: ops.conv_t3(x.view(1, T, H * W, C), w3, ops.make_ep(o16, bias=b, gn_stats=(st.zero_(), 4, 32, T * H * W))),
    "t3_res": lambda: ops.conv_t3(x.view(1, T, H * W, C), w3, ops.make_ep(o32, bias=b, res1=o32)),
}
for name, fn in cases.items():
    if which != "all" and which != name:
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    K = (9 if name[0] == "s" else 3) * C
    print(f"{name}: {ms:.3f} ms  {2.0 * T * H * W * C * K / ms / 1e9:.0f} TF  mode={os.environ.get('GCD_TC_MODE', 'auto')}", flush=True)
