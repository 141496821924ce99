"""Evaluation-loop image metrics: oracle pinned to the reference's `masked_ssim` (CPU), CUDA kernel vs
the oracle and the golden (GPU). Tolerances: SSIM / PSNR are float32 window sums in a different order than scipy's running sums:
|dSSIM| < 2e-6, |dPSNR| < 1e-4 dB."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "metrics.pt")


def test_oracle_matches_reference_golden():
    from oracle import metrics_oracle as M
    g = torch.load(GOLD)
    for t in range(g["gt"].shape[0]):
        a, m = M.ssim_pair(g["pred"][t].numpy(), g["gt"][t].numpy(), g["mask"][t].numpy())
        assert abs(a - g["ssim_all"][t]) < 1e-12
        assert (np.isnan(m) and np.isnan(g["ssim_masked"][t])) or abs(m - g["ssim_masked"][t]) < 1e-12
        assert abs(M.psnr(g["pred"][t].numpy(), g["gt"][t].numpy()) - g["psnr"][t]) < 1e-12


@pytest.mark.gpu
def test_frame_metrics_kernel_vs_reference_golden_and_oracle():
    from gcd_b200 import metrics
    from oracle import metrics_oracle as M
    g = torch.load(GOLD)
    r = metrics.frame_metrics(g["pred"].cuda(), g["gt"].cuda(), g["mask"].cuda())
    for t in range(g["gt"].shape[0]):
        assert abs(r["ssim"][t].item() - g["ssim_all"][t]) < 2e-6
        assert abs(r["psnr"][t].item() - g["psnr"][t]) < 1e-4
        if np.isnan(g["ssim_masked"][t]):
            assert torch.isnan(r["ssim_masked"][t])
        else:
            assert abs(r["ssim_masked"][t].item() - g["ssim_masked"][t]) < 2e-6
    # ragged size (not a multiple of the 32x16 tile), masked PSNR, identical images -> SSIM 1 / PSNR inf
    gen = torch.Generator().manual_seed(5)
    a = torch.rand(2, 3, 45, 71, generator=gen)
    b = (a + 0.1 * torch.randn(2, 3, 45, 71, generator=gen)).clamp(0, 1)
    m = torch.zeros(2, 45, 71, dtype=torch.bool)
    m[0, 4:41, 6:66] = True; m[0, 20:22, 30:33] = False        # a region with a hole (erosion eats 3 pixels around it)
    m[1, :, :35] = True                                         # touches three image borders (erosion border_value = 0)
    r = metrics.frame_metrics(a.cuda(), b.cuda(), m.cuda())
    for t in range(2):
        sa, sm = M.ssim_pair(a[t].numpy(), b[t].numpy(), m[t].numpy())
        assert abs(r["ssim"][t].item() - sa) < 2e-6 and abs(r["ssim_masked"][t].item() - sm) < 2e-6
        mb = m[t][None].expand(3, -1, -1)
        assert abs(r["psnr_masked"][t].item() - M.psnr(a[t][mb].numpy(), b[t][mb].numpy())) < 1e-4   # test.py:395-399
    r = metrics.frame_metrics(a.cuda(), a.cuda())
    assert torch.isinf(r["psnr"]).all() and (r["ssim"] - 1).abs().max() < 1e-6


@pytest.mark.gpu
def test_calculate_metrics_matches_reference_aggregation():
    """scripts/test.py:346-496: per-sample means over frames (nanmean), visible / occluded split, diversity."""
    from gcd_b200 import metrics
    from oracle import metrics_oracle as M
    gen = torch.Generator().manual_seed(9)
    S, T, H, W = 2, 3, 40, 64
    gt = torch.rand(T, 3, H, W, generator=gen)
    rep = gt.clone()
    rep[:, :, :, 40:] = 0.0                                    # right part occluded in the re-projection
    rep[2] = 0.0                                               # one frame fully occluded: visible metrics are NaN there
    pred = (gt[None] + 0.05 * torch.randn(S, T, 3, H, W, generator=gen)).clamp(0, 1)
    out, unc = metrics.calculate_metrics(gt.cuda(), rep.cuda(), pred.cuda())
    occ = rep.abs().sum(1) <= 1e-7
    for s in range(S):
        ps = [M.psnr(pred[s, t].numpy(), gt[t].numpy()) for t in range(T)]
        ss = [M.ssim_pair(pred[s, t].numpy(), gt[t].numpy())[0] for t in range(T)]
        sv = [M.ssim_pair(pred[s, t].numpy(), gt[t].numpy(), (~occ[t]).numpy())[1] for t in range(T)]
        assert abs(out["mean_psnr"][s].item() - np.mean(ps)) < 1e-4 and abs(out["mean_ssim"][s].item() - np.mean(ss)) < 2e-6
        assert abs(out["mean_ssim_vis"][s].item() - np.nanmean(sv)) < 2e-6 and np.isnan(sv[2])
    ref_div = np.nanmean(np.nanmean(np.std(pred.numpy(), axis=0), axis=1), axis=(1, 2))
    assert np.allclose(out["frame_diversity"].numpy(), ref_div, atol=1e-6)
    assert unc.shape == (T, H, W)
