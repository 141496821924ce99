"""VAE Encoder of 14 conditioning frames at 576x1024: time + per-class breakdown."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops, spec, synthetic, flops
from gcd_b200.vae import Encoder
cfg = spec.VAE_ENCODER
enc = Encoder(**spec.encoder_ctor_kwargs(cfg))
enc.load_state_dict(synthetic.seeded_state(spec.encoder_param_shapes(cfg), seed=0))
enc = enc.cuda()
x = torch.rand(14, 3, 576, 1024, device="cuda") * 2 - 1
qw, qb = torch.randn(8, 8, 1, 1, device="cuda") * 0.4, torch.zeros(8, device="cuda")
for _ in range(2):
    z = enc.encode_mode(x, qw, qb, 0.18215)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    z = enc.encode_mode(x, qw, qb, 0.18215)
e1.record(); torch.cuda.synchronize()
print(f"encode 14x576x1024 -> {tuple(z.shape)}: {e0.elapsed_time(e1) / 3:.1f} ms; finite={bool(torch.isfinite(z).all())}")
ops.DETAIL = True
with ops.profile() as prof:
    enc.encode_mode(x, qw, qb, 0.18215)
rows = sorted(prof.summary().items(), key=lambda kv: -kv[1]["ms"])
tf = sum(d["flops"] for _, d in rows) / 1e12
print(f"tensor-core TFLOP: {tf:.1f}")
for k, d in rows[:14]:
    print(f"{d['ms']:8.2f} ms  x{d['launches']:4d}  {d['flops']/max(d['ms'],1e-9)/1e9:7.0f} TF  {d['bytes']/max(d['ms'],1e-9)/1e6:7.0f} GB/s  {k}")
