"""Run each kernel twice on identical inputs at BASELINE shapes and report bitwise mismatches (race detector)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import ops
AD = ops.act_dtype()
dev = "cuda"
def rand(*s, dt=AD, sc=0.5): return (torch.randn(*s, device=dev) * sc).to(dt)
def report(name, a, b):
    d = (a.float() - b.float()).abs().max().item()
    print(f"{'OK ' if d == 0 else 'DIFF'} {name}: max abs diff {d:.3e}  (max |x| {a.float().abs().max().item():.3e})", flush=True)
M = 258048
# linears
for (K, N, mode) in [(320, 320, "res32"), (320, 960, "f16"), (320, 2560, "geglu"), (1280, 320, "res32"), (1280, 320, "blend"), (640, 640, "res32")]:
    x, w, b = rand(M, K), rand(N, K, sc=K ** -0.5), torch.randn(N, device=dev)
    outs = []
    for rep in range(2):
        if mode == "res32":
            o = torch.ones(M, N, device=dev)
            ops.linear(x, w, ops.make_ep(o, bias=b, res1=o))
        elif mode == "f16":
            o = torch.empty(M, N, device=dev, dtype=AD); ops.linear(x, w, ops.make_ep(o, bias=b))
        elif mode == "geglu":
            o = torch.empty(M, N // 2, device=dev, dtype=AD); ops.linear(x, w, ops.make_ep(o, bias=b, geglu=True))
        else:
            r1, r2 = torch.ones(M, N, device=dev), torch.full((M, N), 2.0, device=dev)
            o = torch.empty(M, N, device=dev, dtype=AD)
            ops.linear(x, w, ops.make_ep(o, bias=b, a_acc=0.4, res1=r1, a_res1=0.4, res2=r2, a_res2=0.6))
        outs.append(o)
    torch.cuda.synchronize(); report(f"linear K{K} N{N} {mode}", outs[0], outs[1])
# convs with stats
n, H, W = 28, 72, 128
for (C, Co, f32) in [(320, 320, True), (320, 320, False), (640, 320, True)]:
    x, w, b = rand(n, H, W, C), rand(Co, 9 * C, sc=(9 * C) ** -0.5), torch.randn(Co, device=dev)
    outs, sts = [], []
    for rep in range(2):
        o = torch.ones(n * H * W, Co, device=dev) if f32 else torch.empty(n * H * W, Co, device=dev, dtype=AD)
        st = torch.zeros(n * 64, device=dev, dtype=torch.float64)
        ops.conv2d_3x3(x, w, ops.make_ep(o, bias=b, res1=o if f32 else None, gn_stats=(st, Co // 32, 32, H * W)))
        outs.append(o); sts.append(st)
    torch.cuda.synchronize(); report(f"conv {C}->{Co} f32={f32}", outs[0], outs[1]); report("   its fused stats", sts[0], sts[1])
x = rand(2, 14, H * W, 320); w = rand(320, 3 * 320, sc=960 ** -0.5)
outs = []
for rep in range(2):
    o = torch.ones(28 * H * W, 320, device=dev); ops.conv_t3(x, w, ops.make_ep(o, a_acc=0.3, res1=o)); outs.append(o)
torch.cuda.synchronize(); report("conv_t3 320 inplace", outs[0], outs[1])
# attention
qkv = rand(28, 9216, 960)
outs = []
for rep in range(2):
    o = torch.empty(28, 9216, 320, device=dev, dtype=AD); ops.attention_spatial(qkv, 28, 9216, 5, o); outs.append(o)
torch.cuda.synchronize(); report("attn_spatial 9216", outs[0], outs[1])
outs = []
for rep in range(2):
    o = torch.empty(28, 9216, 320, device=dev, dtype=AD); ops.attention_temporal(qkv, 2, 14, 9216, 5, o); outs.append(o)
torch.cuda.synchronize(); report("attn_temporal", outs[0], outs[1])
# norms
x = torch.randn(M, 320, device=dev); g, b = torch.ones(320, device=dev), torch.zeros(320, device=dev)
outs = []
for rep in range(2):
    o = torch.empty(M, 320, device=dev, dtype=AD); st = torch.empty(28 * 64, device=dev, dtype=torch.float64)
    ops.groupnorm(x, 28, 9216, 320, g, b, 1e-5, True, o, st); outs.append(o)
torch.cuda.synchronize(); report("groupnorm", outs[0], outs[1])
outs = []
for rep in range(2):
    o = torch.empty(M, 320, device=dev, dtype=AD); ops.layernorm(x, g, b, o); outs.append(o)
torch.cuda.synchronize(); report("layernorm", outs[0], outs[1])
