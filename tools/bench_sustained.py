"""Sustained (power-capped) throughput of single ops: each runs back to back for `--seconds` (default 4) so that the GPU settles at
its power-limited clock like inside bench.py's step — the short bursts of tools/bench_ops.py run at 1.9 GHz and overstate what a
compute-bound kernel delivers in situ. Reports TFLOP/s (or GB/s), the median SM clock and power (nvidia-smi, sampled during the run)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gcd_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("sel", nargs="*", default=["conv", "geglu", "attn"])
args = ap.parse_args()
AD = ops.act_dtype()
dev = "cuda"


def rand(*s, dt=AD, sc=0.5):
    return (torch.randn(*s, device=dev) * sc).to(dt)


def smi_start():
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"],
                         stdout=f, stderr=subprocess.DEVNULL)
    return p, f


def smi_stop(p, f):
    p.terminate(); p.wait(timeout=5)
    f.flush(); f.seek(0)
    rows = [l.strip().split(", ") for l in f.read().strip().splitlines() if l.strip()]
    os.unlink(f.name)
    rows = rows[len(rows) // 3:]                       # settled part
    if not rows:
        return None, None
    clk = sorted(float(r[0]) for r in rows)[len(rows) // 2]
    pw = sorted(float(r[1]) for r in rows)[len(rows) // 2]
    return clk, pw


def sustained(name, fn, flops=None, nbytes=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    p, f = smi_start()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < args.seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    clk, pw = smi_stop(p, f)
    d = {"name": name, "ms": round(ms, 4), "sm_mhz": clk, "watts": pw}
    if flops:
        d["tflops"] = round(flops / ms / 1e9, 1)
    if nbytes:
        d["gbs"] = round(nbytes / ms / 1e6, 1)
    print(json.dumps(d), flush=True)


if "conv" in args.sel:
    n, H, W, C, Co = 28, 72, 128, 320, 320
    x, w = rand(n, H, W, C), rand(Co, 9 * C, sc=0.02)
    out = torch.zeros(n * H * W, Co, device=dev)
    ep = ops.make_ep(out, bias=torch.zeros(Co, device=dev), res1=out, a_res1=0.5)
    sustained("conv3x3 28x72x128 320->320 f32 residual", lambda: ops.conv2d_3x3(x, w, ep), 2.0 * n * H * W * C * Co * 9)
    n, H, W, C, Co = 28, 36, 64, 640, 640
    x2, w2 = rand(n, H, W, C), rand(Co, 9 * C, sc=0.02)
    out2 = torch.empty(n * H * W, Co, device=dev, dtype=AD)
    ep2 = ops.make_ep(out2, bias=torch.zeros(Co, device=dev))
    sustained("conv3x3 28x36x64 640->640 f16 out", lambda: ops.conv2d_3x3(x2, w2, ep2), 2.0 * n * H * W * C * Co * 9)
if "geglu" in args.sel:
    M, K, N = 258048, 320, 2560
    xg, wg = rand(M, K), rand(N, K, sc=0.05)
    og = torch.empty(M, N // 2, device=dev, dtype=AD)
    epg = ops.make_ep(og, bias=torch.zeros(N, device=dev), geglu=True)
    sustained("geglu M258048 K320 N2560", lambda: ops.linear(xg, wg, epg), 2.0 * M * K * N)
if "attn" in args.sel:
    f, tok, h = 28, 9216, 5
    qkv = rand(f, tok, 3 * h * 64, sc=1.0)
    oa = torch.empty(f, tok, h * 64, device=dev, dtype=AD)
    sustained("attn_spatial 28x9216x5", lambda: ops.attention_spatial(qkv, f, tok, h, oa), 4.0 * f * h * tok * tok * 64)
