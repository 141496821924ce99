"""GPU parity of the CUDA VideoUNet against (a) golden outputs produced by the REFERENCE's own VideoUNet
(tests/golden/unet_*.pt, made by oracle/pin_against_reference.py) and (b) the CPU oracle run live."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

# stated tolerance: relative L2 error of the network output vs the fp32 reference, fp16 tensor-core operands
TOL_REL_L2 = 4e-3


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def _net_inputs(gold):
    from oracle import gcd_oracle as O, weights
    cfg, B, T, H, W = gold["cfg"], gold["B"], gold["T"], gold["H"], gold["W"]
    x, c, uc, ioi = weights.seeded_inputs(cfg, B, T, H, W)
    sigma = torch.full((2 * B * T,), gold["sigma"])
    c_cat = {k: torch.cat((uc[k], c[k]), 0) for k in c}
    cs, co, ci, cn = O.vscaling_edm_cnoise(sigma.view(-1, 1, 1, 1))
    xin = torch.cat((torch.cat([x, x]) * gold["x_mul"] * ci, c_cat["concat"]), 1)
    return xin, cn.reshape(-1), c_cat, ioi


@pytest.mark.parametrize("tag", ["tiny", "kubric", "pardom"])
def test_unet_forward_vs_reference_golden(tag):
    from gcd_b200 import spec
    from gcd_b200.unet import VideoUNet
    from oracle import weights
    path = os.path.join(GOLD, f"unet_{tag}.pt")
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    gold = torch.load(path)
    cfg = gold["cfg"]
    net = VideoUNet(**spec.unet_ctor_kwargs(cfg))
    missing = net.load_state_dict(weights.seeded_state(spec.unet_param_shapes(cfg), seed=0), strict=True)
    net = net.cuda()
    xin, t, c_cat, ioi = _net_inputs(gold)
    out = net(xin.cuda(), t.cuda(), context=c_cat["crossattn"].cuda(), y=c_cat["vector"].cuda(),
              num_video_frames=gold["T"], image_only_indicator=ioi.cuda())
    assert out.shape == gold["net_out"].shape and out.dtype == torch.float32
    e = relerr(out, gold["net_out"])
    print(f"unet[{tag}] rel-L2 vs reference golden: {e:.3e}")
    assert e < TOL_REL_L2


def test_unet_forward_vs_oracle_live():
    """Different spatial size / frame count than the golden, oracle computed on the host CPU in the same test."""
    from gcd_b200 import spec
    from gcd_b200.unet import VideoUNet
    from oracle import gcd_oracle as O, weights
    cfg = spec.UNET_TINY
    sd = weights.seeded_state(spec.unet_param_shapes(cfg), seed=3)
    net = VideoUNet(**spec.unet_ctor_kwargs(cfg))
    net.load_state_dict(sd)
    net = net.cuda()
    B, T, H, W = 2, 4, 8, 24
    x, c, uc, ioi = weights.seeded_inputs(cfg, B, T, H, W, seed=99)
    xin = torch.cat((x * 0.3, c["concat"]), 1)
    t = torch.linspace(-0.3, 1.2, B * T)
    ioi = torch.zeros(B, T)
    ref = O.unet_forward(sd, cfg, xin, t, c["crossattn"], c["vector"], T, ioi)
    out = net(xin.cuda(), t.cuda(), context=c["crossattn"].cuda(), y=c["vector"].cuda(), num_video_frames=T,
              image_only_indicator=ioi.cuda())
    e = relerr(out, ref)
    print(f"unet[tiny live] rel-L2 vs oracle: {e:.3e}")
    assert e < TOL_REL_L2
    # half-precision input (the reference runs under fp16 autocast): output dtype follows x (video_model.py:534)
    out16 = net(xin.cuda().half(), t.cuda(), context=c["crossattn"].cuda(), y=c["vector"].cuda(), num_video_frames=T,
                image_only_indicator=ioi.cuda())
    assert out16.dtype == torch.float16


@pytest.mark.parametrize("B,T,H,W", [(3, 5, 8, 40), (1, 16, 24, 8), (2, 1, 16, 16)])
def test_unet_ragged_shapes_vs_oracle(B, T, H, W):
    """Shapes whose tiles do not divide evenly (W=40 -> 8-wide tiles, 5/16/1 frames, 3 clips) against the CPU oracle."""
    from gcd_b200 import spec
    from gcd_b200.unet import VideoUNet
    from oracle import gcd_oracle as O, weights
    cfg = spec.UNET_TINY
    sd = weights.seeded_state(spec.unet_param_shapes(cfg), seed=5)
    net = VideoUNet(**spec.unet_ctor_kwargs(cfg))
    net.load_state_dict(sd)
    net = net.cuda()
    x, c, uc, _ = weights.seeded_inputs(cfg, B, T, H, W, seed=7)
    xin = torch.cat((x * 0.5, c["concat"]), 1)
    t = torch.linspace(-0.2, 1.0, B * T)
    ioi = torch.zeros(B, T)
    ref = O.unet_forward(sd, cfg, xin, t, c["crossattn"], c["vector"], T, ioi)
    out = net(xin.cuda(), t.cuda(), context=c["crossattn"].cuda(), y=c["vector"].cuda(), num_video_frames=T,
              image_only_indicator=ioi.cuda())
    e = relerr(out, ref)
    print(f"unet[tiny B{B} T{T} {H}x{W}] rel-L2 vs oracle: {e:.3e}")
    assert e < TOL_REL_L2


def test_unet_full_size_properties():
    """BASELINE-size (28 frames, 72x128, full Kubric width) checks that need no oracle: run-to-run reproducibility and
    clip independence (a clip's result does not depend on what else is in the batch — GroupNorm, temporal attention
    and context[::T] are all per clip)."""
    from gcd_b200 import spec, synthetic
    from gcd_b200.unet import VideoUNet
    cfg = spec.UNET_KUBRIC
    net = VideoUNet(**spec.unet_ctor_kwargs(cfg))
    net.load_state_dict(synthetic.seeded_state(spec.unet_param_shapes(cfg), seed=0))
    net = net.cuda()
    T, H, W = 14, 72, 128
    xa, ca, _, _ = synthetic.seeded_inputs(cfg, 1, T, H, W, seed=11)
    xb, cb, _, _ = synthetic.seeded_inputs(cfg, 1, T, H, W, seed=12)
    mk = lambda x, c: (torch.cat((x * 0.3, c["concat"]), 1).cuda(), c["crossattn"].cuda(), c["vector"].cuda())
    (xa_, ctxa, ya), (xb_, ctxb, yb) = mk(xa, ca), mk(xb, cb)
    t1, t2 = torch.full((T,), 0.57).cuda(), torch.full((2 * T,), 0.57).cuda()
    ioi1, ioi2 = torch.zeros(1, T).cuda(), torch.zeros(2, T).cuda()
    both = net(torch.cat((xa_, xb_)), t2, context=torch.cat((ctxa, ctxb)), y=torch.cat((ya, yb)), num_video_frames=T,
               image_only_indicator=ioi2)
    again = net(torch.cat((xa_, xb_)), t2, context=torch.cat((ctxa, ctxb)), y=torch.cat((ya, yb)), num_video_frames=T,
                image_only_indicator=ioi2)
    only_a = net(xa_, t1, context=ctxa, y=ya, num_video_frames=T, image_only_indicator=ioi1)
    assert torch.isfinite(both).all() and both.abs().max() > 0
    # same launch twice: every kernel is order-deterministic except the fp64 statistics atomics (1e-16 level)
    assert relerr(again, both) < 1e-6
    # clip 0 alone vs clip 0 next to clip 1: mathematically identical. The only batch-dependent arithmetic is the
    # summation order of GroupNorm statistics on concatenated tensors (~1e-7), which the 16-bit intermediate roundings
    # then decorrelate to the network's rounding-noise floor (tools/bisect_batch.py: the whole down path and middle block
    # are bit-identical, the difference appears after the first concat GroupNorm and stays below the parity error).
    assert relerr(both[:T], only_a) < 2e-3
