"""CPU restatement of the evaluation-loop image metrics (TEST INFRASTRUCTURE — only tests may import this).

* PSNR: scikit-image 0.22.0 (`requirements_versions.txt:40`, not vendored under /root/reference)
  `skimage.metrics.peak_signal_noise_ratio(a, b, data_range=1)` = 10 log10(1 / mean((a-b)^2)), the mean taken in float64.
* SSIM: scikit-image 0.22.0 `structural_similarity(a, b, data_range=1, channel_axis=0)` with its defaults (7x7 uniform window,
  sample covariance, K1 = 0.01, K2 = 0.03, 3-pixel border cropped, mean over channels), which the reference itself restates in
  gcd-model/scripts/eval_utils.py:571-664 (`masked_ssim`: element [0] of its result is the scikit-image value, element [1] the
  mean over `binary_erosion(mask, iterations=3)`).
Pinned: tests/golden/metrics.pt holds outputs of the REFERENCE's `masked_ssim` source executed in the build container
(oracle/pin_metrics.py; the four scikit-image helper functions it imports are one-liners supplied by the pin script)."""
import numpy as np
from scipy.ndimage import binary_erosion, uniform_filter


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2, dtype=np.float64)
    return 10.0 * np.log10(1.0 / mse)


def ssim_pair(im1, im2, mask=None, win=7, K1=0.01, K2=0.03):
    """im1, im2: (3, H, W) float32 in [0, 1]; mask: (H, W) bool or None. Returns (ssim_all, ssim_masked or nan)."""
    res = []
    pad = (win - 1) // 2
    for ch in range(im1.shape[0]):
        x, y = im1[ch].astype(np.float32), im2[ch].astype(np.float32)
        NP = win ** 2
        cov = NP / (NP - 1)
        ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
        uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
        vx, vy, vxy = cov * (uxx - ux * ux), cov * (uyy - uy * uy), cov * (uxy - ux * uy)
        C1, C2 = K1 ** 2, K2 ** 2
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        Sc = S[pad:-pad, pad:-pad]
        allv = np.mean(Sc, dtype=np.float64)
        mv = np.nan
        if mask is not None:
            mc = binary_erosion(mask.astype(bool), iterations=pad)[pad:-pad, pad:-pad]
            mv = np.mean(Sc[mc], dtype=np.float64) if mc.any() else np.nan
        res.append((allv, mv))
    return tuple(np.mean(res, axis=0))
