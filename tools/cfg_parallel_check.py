"""torchrun --nproc-per-node 2 tools/cfg_parallel_check.py [--full]
CFG-parallel sampling (one clip split over two GPUs: uc on rank 0, c on rank 1) vs the single-GPU fused sampler:
parity on the tiny network, and with --full the single-clip latency at the Kubric size (25 steps + decode)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datetime
import torch
import torch.distributed as dist
from gcd_b200 import spec, synthetic
from gcd_b200.pipeline import GCDHotPath

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=600))
assert world == 2
grp = dist.new_group([0, 1])
full = "--full" in sys.argv


def build(ucfg, vcfg, steps):
    pipe = GCDHotPath(ucfg, vcfg, num_steps=steps, device="cuda")
    pipe.load_state(synthetic.seeded_state(spec.unet_param_shapes(ucfg)), synthetic.seeded_state(spec.decoder_param_shapes(vcfg)))
    return pipe


def inputs(cfg, T, H, W):
    x, c, uc, _ = synthetic.seeded_inputs(cfg, 1, T, H, W)
    cu = lambda d: {k: v.cuda() for k, v in d.items()}
    return x.cuda(), cu(c), cu(uc)


out = {}
# ---- parity, tiny network, 4 steps
pipe = build(spec.UNET_TINY, spec.VAE_TINY, 4)
x, c, uc = inputs(spec.UNET_TINY, 14, 16, 24)
ref = pipe.sample_latents(x.clone(), c, uc)
pipe.set_cfg_parallel(grp)
par = pipe.sample_latents(x.clone(), c, uc)
pipe.set_cfg_parallel(None)
err = ((par - ref).norm() / ref.norm()).item()
both = [torch.empty_like(par) for _ in range(2)]
dist.all_gather(both, par)
out["tiny_rel_l2_vs_single_gpu"] = err
out["ranks_bit_identical"] = bool(torch.equal(both[0], both[1]))
assert err < 5e-3 and out["ranks_bit_identical"], out
del pipe
torch.cuda.empty_cache()

if full:
    pipe = build(spec.UNET_KUBRIC, spec.VAE_DECODER, 25)
    x, c, uc = inputs(spec.UNET_KUBRIC, 14, 72, 128)

    def clip(par_mode):
        pipe.set_cfg_parallel(grp if par_mode else None)
        z = pipe.sample_latents(x.clone(), c, uc)
        if rank == 0 or not par_mode:
            pipe.decode_first_stage(z)
        return z

    for mode in (False, True):
        clip(mode)                                   # warm-up
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            z = clip(mode)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 2], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["cfg_parallel_ms_per_clip" if mode else "single_gpu_ms_per_clip"] = t.item()
        if mode:
            zr = z.clone()
        else:
            z0 = z.clone()
    out["full_rel_l2_vs_single_gpu"] = ((zr - z0).norm() / z0.norm()).item()
    out["speedup"] = out["single_gpu_ms_per_clip"] / out["cfg_parallel_ms_per_clip"]
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
