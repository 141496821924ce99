// gcd_b200 — per-frame image metrics of the evaluation loop: PSNR and SSIM, full-frame and over a
// mask, of decoded frames against ground truth — gcd-model/scripts/test.py:346-496 (calculate_metrics), whose SSIM is
// scikit-image 0.22.0 `structural_similarity(data_range=1, channel_axis=0)` and gcd-model/scripts/eval_utils.py:571-664
// (masked_ssim: the same map averaged over an eroded mask): 7x7 uniform window, sample covariance (NP/(NP-1)), K1 = 0.01,
// K2 = 0.03, the 3-pixel border strip ignored, mean over the three channels.
// HBM-bound streaming work (2 x 4 B per pixel-channel read once; 10 MB per 576x1024 frame pair): one block per 32x16 output tile,
// the 38x22 halo of both images staged in shared memory, every thread sums its own 7x7 window (fp32, like the float32 images
// of the reference), block reduction in fp64, one fp64 atomic per block and quantity.
#include "common.cuh"
#include "../../include/gcd_b200.h"
#include <atomic>
extern std::atomic<int64_t> g_launches;

namespace {
constexpr int TX = 32, TY = 16, R = 3, HX = TX + 2 * R, HY = TY + 2 * R;

// out per frame: [0] sum sq. diff (all), [1] count, [2] sum sq. diff (mask), [3] count (mask), [4] sum SSIM (cropped map),
// [5] count, [6] sum SSIM over the eroded mask, [7] count  — sums over the 3 channels
__global__ void __launch_bounds__(TX * TY)
frame_metrics_kernel(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ mask, int H, int W,
                     double* __restrict__ out) {
    __shared__ float sa[HY][HX + 1], sb[HY][HX + 1];
    __shared__ uint8_t sm[HY][HX + 1];
    __shared__ double red[8][TX * TY / 32];
    const int f = blockIdx.z / 3, c = blockIdx.z % 3;
    const float* pa = a + ((int64_t)f * 3 + c) * H * W;
    const float* pb = b + ((int64_t)f * 3 + c) * H * W;
    const uint8_t* pm = mask ? mask + (int64_t)f * H * W : nullptr;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    for (int i = threadIdx.x; i < HX * HY; i += TX * TY) {
        const int ly = i / HX, lx = i % HX, y = y0 + ly - R, x = x0 + lx - R;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;
        sa[ly][lx] = in ? pa[(int64_t)y * W + x] : 0.f;
        sb[ly][lx] = in ? pb[(int64_t)y * W + x] : 0.f;
        sm[ly][lx] = (in && pm) ? pm[(int64_t)y * W + x] : 0;      // binary_erosion border_value = 0
    }
    __syncthreads();
    const int lx = threadIdx.x % TX, ly = threadIdx.x / TX, x = x0 + lx, y = y0 + ly;
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (x < W && y < H) {
        const float d = sa[ly + R][lx + R] - sb[ly + R][lx + R];
        v[0] = (double)d * d; v[1] = 1.0;
        if (pm && sm[ly + R][lx + R]) { v[2] = v[0]; v[3] = 1.0; }
        if (x >= R && x < W - R && y >= R && y < H - R) {          // crop(S, pad): windows fully inside the image
            float ux = 0.f, uy = 0.f, uxx = 0.f, uyy = 0.f, uxy = 0.f;
#pragma unroll
            for (int dy = 0; dy < 7; dy++)
#pragma unroll
                for (int dx = 0; dx < 7; dx++) {
                    const float p = sa[ly + dy][lx + dx], q = sb[ly + dy][lx + dx];
                    ux += p; uy += q; uxx += p * p; uyy += q * q; uxy += p * q;
                }
            const float inv = 1.0f / 49.0f, cov = 49.0f / 48.0f;
            ux *= inv; uy *= inv; uxx *= inv; uyy *= inv; uxy *= inv;
            const float vx = cov * (uxx - ux * ux), vy = cov * (uyy - uy * uy), vxy = cov * (uxy - ux * uy);
            const float C1 = 1e-4f, C2 = 9e-4f;
            const float S = ((2.f * ux * uy + C1) * (2.f * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
            v[4] = S; v[5] = 1.0;
            if (pm) {          // binary_erosion(mask, iterations = 3) with the cross element = all pixels within L1 distance 3
                bool ok = true;
#pragma unroll
                for (int dy = -R; dy <= R; dy++)
#pragma unroll
                    for (int dx = -R; dx <= R; dx++)
                        if (abs(dx) + abs(dy) <= R) ok = ok && sm[ly + R + dy][lx + R + dx];
                if (ok) { v[6] = S; v[7] = 1.0; }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0.0;
        for (int w = 0; w < TX * TY / 32; w++) s += red[threadIdx.x][w];
        if (s != 0.0) atomicAdd(out + (int64_t)f * 8 + threadIdx.x, s);
    }
}
}  // namespace

extern "C" int gcd_frame_metrics(const float* pred, const float* gt, const uint8_t* mask, int frames, int H, int W, double* out,
                                 void* stream) {
    GCD_REQUIRE(pred && gt && out && frames > 0 && H >= 7 && W >= 7, "frame_metrics: need [frames, 3, H >= 7, W >= 7] images");
    GCD_REQUIRE((int64_t)frames * 3 <= 65535, "frame_metrics: too many frames for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    GCD_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)frames * 8 * sizeof(double), st));
    dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, frames * 3);
    frame_metrics_kernel<<<grid, TX * TY, 0, st>>>(pred, gt, mask, H, W, out);
    GCD_CUDA_CHECK(cudaGetLastError());
    g_launches++;
    return 0;
}
