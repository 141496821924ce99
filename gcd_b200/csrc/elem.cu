// gcd_b200 — element-wise / layout kernels and the fused EDM-Euler sampler step (all HBM-bound, tiny next to the UNet).
//   sampler step : gcd-model/sgm/modules/diffusionmodules/sampling.py:86-121 (euler_step, sampler_step),
//                  sampling_utils.py:34-35 (to_d), denoiser.py:23-49 + denoiser_scaling.py:53-61 (VScalingWithEDMcNoise),
//                  guiders.py:79-100 (LinearPredictionGuider), wrappers.py:23-34 (OpenAIWrapper concat).
#include "common.cuh"
#include "../../include/gcd_b200.h"
#include <atomic>
extern std::atomic<int64_t> g_launches;

#define LAUNCH_1D(kernel, n, st, ...)                                              \
    do {                                                                           \
        int64_t _n = (n);                                                          \
        if (_n > 0) {                                                              \
            int64_t _b = (_n + 255) / 256;                                         \
            GCD_REQUIRE(_b < (1ll << 31), #kernel ": too many elements");          \
            kernel<<<(unsigned)_b, 256, 0, (cudaStream_t)(st)>>>(__VA_ARGS__);     \
            GCD_CUDA_CHECK(cudaGetLastError());                                    \
            g_launches++;                                                          \
        }                                                                          \
    } while (0)

extern "C" int gcd_memset_async(void* p, int value, int64_t bytes, void* stream) {
    GCD_CUDA_CHECK(cudaMemsetAsync(p, value, (size_t)bytes, (cudaStream_t)stream));
    return 0;
}

__global__ void cast_kernel(const float* __restrict__ in, int64_t n4, act_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = reinterpret_cast<const float4*>(in)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
}
extern "C" int gcd_cast_f32_to_act(const float* in, int64_t n, void* out, void* stream) {
    GCD_REQUIRE(n % 4 == 0, "cast: n must be a multiple of 4");
    LAUNCH_1D(cast_kernel, n / 4, stream, in, n / 4, (act_t*)out);
    return 0;
}

__global__ void upsample2x_kernel(const float* __restrict__ in, int n, int H, int W, int C4, act_t* __restrict__ out) {
    // one thread per 4 output channels of one OUTPUT pixel
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)n * (2 * H) * (2 * W) * C4;
    if (i >= total) return;
    int c = (int)(i % C4);
    int64_t p = i / C4;
    int wo = (int)(p % (2 * W));
    int64_t q = p / (2 * W);
    int ho = (int)(q % (2 * H));
    int img = (int)(q / (2 * H));
    float4 v = reinterpret_cast<const float4*>(in)[(((int64_t)img * H + (ho >> 1)) * W + (wo >> 1)) * C4 + c];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
}
extern "C" int gcd_upsample2x_to_act(const float* in, int n, int H, int W, int C, void* out, void* stream) {
    GCD_REQUIRE(C % 4 == 0, "upsample: C must be a multiple of 4");
    LAUNCH_1D(upsample2x_kernel, (int64_t)n * 4 * H * W * (C / 4), stream, in, n, H, W, C / 4, (act_t*)out);
    return 0;
}

__global__ void concat_kernel(const float4* __restrict__ a, int Ca4, const float4* __restrict__ b, int Cb4, int64_t rows,
                              float4* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int Ct = Ca4 + Cb4;
    if (i >= rows * Ct) return;
    int c = (int)(i % Ct);
    int64_t r = i / Ct;
    out[i] = (c < Ca4) ? a[r * Ca4 + c] : b[r * Cb4 + (c - Ca4)];
}
extern "C" int gcd_concat_channels(const float* a, int Ca, const float* b, int Cb, int64_t rows, float* out, void* stream) {
    GCD_REQUIRE(Ca % 4 == 0 && Cb % 4 == 0, "concat: channel counts must be multiples of 4");
    LAUNCH_1D(concat_kernel, rows * ((Ca + Cb) / 4), stream, (const float4*)a, Ca / 4, (const float4*)b, Cb / 4, rows,
              (float4*)out);
    return 0;
}

__global__ void silu_kernel(const act_t* __restrict__ in, int64_t n, act_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f2act(ptx::silu(act2f(in[i])));
}
extern "C" int gcd_silu_act(const void* in, int64_t n, void* out, void* stream) {
    LAUNCH_1D(silu_kernel, n, stream, (const act_t*)in, n, (act_t*)out);
    return 0;
}

// util.py:207-231 — emb[i, k] = cos(t_i f_k), emb[i, half+k] = sin(t_i f_k), f_k = exp(-ln(max_period) k / half)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int n, int dim, float neg_log_mp,
                                          act_t* __restrict__ out_act, float* __restrict__ out_f32) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim / 2;
    if (i >= n * half) return;
    int k = i % half, r = i / half;
    float f = expf(neg_log_mp * (float)k / (float)half);
    float a = t[r] * f;
    float c = cosf(a), s = sinf(a);
    if (out_act) { out_act[(int64_t)r * dim + k] = f2act(c); out_act[(int64_t)r * dim + half + k] = f2act(s); }
    if (out_f32) { out_f32[(int64_t)r * dim + k] = c; out_f32[(int64_t)r * dim + half + k] = s; }
}
extern "C" int gcd_timestep_embedding(const float* t, int n, int dim, float max_period, void* out_act, float* out_f32,
                                      void* stream) {
    GCD_REQUIRE(dim % 2 == 0, "timestep_embedding: odd dim unsupported");
    LAUNCH_1D(timestep_embedding_kernel, (int64_t)n * (dim / 2), stream, t, n, dim, -logf(max_period), (act_t*)out_act,
              out_f32);
    return 0;
}

// encoders/modules.py:247-287 (SphericalEmbedder): x[n,3] = (azimuth, elevation, radius) ->
//   feat = [cos a, sin a, cos 2a, sin 2a, cos 4a, sin 4a, cos e, sin e, cos 2e, sin 2e, cos 4e, sin 4e, r],  out = feat W^T + b
__global__ void spherical_embed_kernel(const float* __restrict__ x, int n, const float* __restrict__ w,
                                       const float* __restrict__ b, int dim, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const int r = i / dim, o = i % dim;
    const float a = x[r * 3], e = x[r * 3 + 1], rad = x[r * 3 + 2];
    float f[13];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float m = (float)(1 << k);
        f[2 * k] = cosf(a * m); f[2 * k + 1] = sinf(a * m);
        f[6 + 2 * k] = cosf(e * m); f[7 + 2 * k] = sinf(e * m);
    }
    f[12] = rad;
    float acc = 0.f;                       // same accumulation order as a 13-long dot product followed by the bias add
#pragma unroll
    for (int k = 0; k < 13; k++) acc = fmaf(f[k], w[o * 13 + k], acc);
    out[i] = acc + b[o];
}
extern "C" int gcd_spherical_embed(const float* x, int n, const float* w, const float* b, int dim, float* out, void* stream) {
    GCD_REQUIRE(x && w && b && out && n > 0 && dim > 0, "spherical_embed: bad arguments");
    LAUNCH_1D(spherical_embed_kernel, (int64_t)n * dim, stream, x, n, w, b, dim, out);
    return 0;
}

// ------------------------------------------------------------------------------------------------ sampler
// One thread per (image of the doubled batch, pixel): writes 64 act channels (128 B).
__global__ void sampler_prep_kernel(const float* __restrict__ x, const float* __restrict__ ucc, const float* __restrict__ cc,
                                    int BT, int HW, float c_in, act_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)2 * BT * HW) return;
    int hw = (int)(i % HW);
    int img = (int)(i / HW);
    int bt = img % BT;
    const float* cat = (img >= BT) ? cc : ucc;
    float v[8];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        v[c] = x[((int64_t)bt * 4 + c) * HW + hw] * c_in;
        v[4 + c] = cat ? cat[((int64_t)bt * 4 + c) * HW + hw] : 0.f;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + i * 64);
    dst[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
#pragma unroll
    for (int j = 1; j < 8; j++) dst[j] = make_uint4(0, 0, 0, 0);
}
extern "C" int gcd_sampler_prep(const float* x, const float* uc_concat, const float* c_concat, int BT, int H, int W,
                                float c_in, void* out, void* stream) {
    LAUNCH_1D(sampler_prep_kernel, (int64_t)2 * BT * H * W, stream, x, uc_concat, c_concat, BT, H * W, c_in, (act_t*)out);
    return 0;
}

__global__ void sampler_update_kernel(float* __restrict__ x, const float* __restrict__ net, int ld, int BT, int T, int HW,
                                      float c_out, float c_skip, float sigma, float dt, const float* __restrict__ scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)BT * HW) return;
    int hw = (int)(i % HW);
    int bt = (int)(i / HW);
    const float sc = scale[bt % T];
    const float* nu = net + ((int64_t)bt * HW + hw) * ld;
    const float* nc = net + ((int64_t)(BT + bt) * HW + hw) * ld;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float* px = x + ((int64_t)bt * 4 + c) * HW + hw;
        const float xv = *px;
        const float du = nu[c] * c_out + xv * c_skip;   // denoiser.py:40-49
        const float dc = nc[c] * c_out + xv * c_skip;
        const float den = du + sc * (dc - du);           // guiders.py:79-87
        const float d = (xv - den) / sigma;              // sampling_utils.py:34-35
        *px = xv + dt * d;                               // sampling.py:86-87
    }
}
extern "C" int gcd_sampler_update(float* x, const float* net_out, int ld_net, int BT, int T, int H, int W, float c_out,
                                  float c_skip, float sigma, float dt, const float* scale, void* stream) {
    LAUNCH_1D(sampler_update_kernel, (int64_t)BT * H * W, stream, x, net_out, ld_net, BT, T, H * W, c_out, c_skip, sigma,
              dt, scale);
    return 0;
}

// ------------------------------------------------------------------------------------------------ layout glue
// NCHW float32 [N, C, HW] -> channels-last act [N, HW, Cpad] (channels >= C zero-filled); one thread per 8 out channels
__global__ void nchw_to_act_nhwc_kernel(const float* __restrict__ in, int N, int C, int HW, int Cpad,
                                        act_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g8 = Cpad / 8;
    if (i >= (int64_t)N * HW * g8) return;
    int g = (int)(i % g8);
    int64_t pix = i / g8;
    int hw = (int)(pix % HW);
    int n = (int)(pix / HW);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int c = g * 8 + j;
        v[j] = c < C ? in[((int64_t)n * C + c) * HW + hw] : 0.f;
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
extern "C" int gcd_nchw_to_act_nhwc(const float* in, int N, int C, int HW, int Cpad, void* out, void* stream) {
    GCD_REQUIRE(Cpad % 8 == 0 && Cpad >= C, "nchw_to_act_nhwc: bad Cpad");
    LAUNCH_1D(nchw_to_act_nhwc_kernel, (int64_t)N * HW * (Cpad / 8), stream, in, N, C, HW, Cpad, (act_t*)out);
    return 0;
}
// channels-last float32 [N, HW, ld] (first C columns) -> NCHW float32 [N, C, HW]
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int ld, int N, int C, int HW, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * C * HW) return;
    int hw = (int)(i % HW);
    int64_t t = i / HW;
    int c = (int)(t % C);
    int n = (int)(t / C);
    out[i] = in[((int64_t)n * HW + hw) * ld + c];
}
extern "C" int gcd_nhwc_to_nchw_f32(const float* in, int ld, int N, int C, int HW, float* out, void* stream) {
    LAUNCH_1D(nhwc_to_nchw_kernel, (int64_t)N * C * HW, stream, in, ld, N, C, HW, out);
    return 0;
}
__global__ void silu_f32_to_act_kernel(const float* __restrict__ in, int64_t n, act_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f2act(ptx::silu(in[i]));
}
extern "C" int gcd_silu_f32_to_act(const float* in, int64_t n, void* out, void* stream) {
    LAUNCH_1D(silu_f32_to_act_kernel, n, stream, in, n, (act_t*)out);
    return 0;
}

// ------------------------------------------------------------------------------------------------ VAE tail
// AE3DConv.time_mix_conv (temporal_ae.py:86-107): Conv3d(3->3, kernel (3,1,1), padding (1,0,0)) over the frame axis of
// the 3-channel conv_out result. in: channels-last float32 [B*T, HW, ld] (first 3 cols); out: NCHW float32 [B*T,3,HW].
__global__ void vae_time_mix_kernel(const float* __restrict__ in, int ld, int B, int T, int HW,
                                    const float* __restrict__ w /*[3][3][3] co,ci,kt*/, const float* __restrict__ b,
                                    float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T * HW) return;
    int hw = (int)(i % HW);
    int bt = (int)(i / HW);
    int t = bt % T;
    float acc[3] = {b[0], b[1], b[2]};
#pragma unroll
    for (int kt = 0; kt < 3; kt++) {
        int tt = t + kt - 1;
        if (tt < 0 || tt >= T) continue;
        const float* px = in + ((int64_t)(bt + kt - 1) * HW + hw) * ld;
#pragma unroll
        for (int ci = 0; ci < 3; ci++) {
            float v = px[ci];
#pragma unroll
            for (int co = 0; co < 3; co++) acc[co] += w[(co * 3 + ci) * 3 + kt] * v;
        }
    }
#pragma unroll
    for (int co = 0; co < 3; co++) out[((int64_t)bt * 3 + co) * HW + hw] = acc[co];
}
extern "C" int gcd_vae_time_mix(const float* in, int ld, int B, int T, int HW, const float* w, const float* b, float* out,
                                void* stream) {
    LAUNCH_1D(vae_time_mix_kernel, (int64_t)B * T * HW, stream, in, ld, B, T, HW, w, b, out);
    return 0;
}
