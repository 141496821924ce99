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
