// gcd_b200 — temporal self-attention over the T (=14) frames of a clip, per (spatial position, head), head_dim 64.
// Reference: gcd-model/sgm/modules/video_attention.py:109-140 (VideoTransformerBlock.attn1 after the
// "(b t) s c -> (b s) t c" rearrange) with CrossAttention math attention.py:255-344: softmax(q k^T / sqrt(64)) v.
// The rearrange is pure indexing here: tokens stay in [clip, t, s, C] order.
// Work is 0.04 TFLOP per UNet forward: CUDA-core kernel, one thread per (clip, s, head, query frame); the 14 threads of
// a (s, head) group read the same K/V rows (warp-broadcast loads). HBM-bound: reads qkv once, writes out once.
#include "common.cuh"
#include "../../include/gcd_b200.h"
#include <atomic>
extern std::atomic<int64_t> g_launches;

template <int GS>  // lanes per (s, head) group: 16 (T <= 16) or 32
__global__ void __launch_bounds__(128)
attn_temporal_kernel(const act_t* __restrict__ qkv, int clips, int T, int tokens, int heads, act_t* __restrict__ out) {
    const int C = heads * 64;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GS;   // (clip, s, head)
    const int t = threadIdx.x % GS;
    const int64_t ngroups = (int64_t)clips * tokens * heads;
    if (gid >= ngroups || t >= T) return;
    const int h = (int)(gid % heads);
    const int64_t cs = gid / heads;
    const int s = (int)(cs % tokens);
    const int b = (int)(cs / tokens);
    const int64_t frame_stride = (int64_t)tokens * 3 * C;                       // elements between frames
    const act_t* base = qkv + ((int64_t)b * T * tokens + s) * 3 * C + h * 64;   // frame 0 of this (b, s, head)

    float q[64];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(base + (int64_t)t * frame_stride);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = qp[i];
            float2 f;
            f = unpack2(v.x); q[i * 8 + 0] = f.x; q[i * 8 + 1] = f.y;
            f = unpack2(v.y); q[i * 8 + 2] = f.x; q[i * 8 + 3] = f.y;
            f = unpack2(v.z); q[i * 8 + 4] = f.x; q[i * 8 + 5] = f.y;
            f = unpack2(v.w); q[i * 8 + 6] = f.x; q[i * 8 + 7] = f.y;
        }
    }
    float sc[32];
    float m = -INFINITY;
    for (int j = 0; j < T; j++) {
        const uint4* kp = reinterpret_cast<const uint4*>(base + (int64_t)j * frame_stride + C);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = kp[i];
            float2 f;
            f = unpack2(v.x); acc += q[i * 8 + 0] * f.x + q[i * 8 + 1] * f.y;
            f = unpack2(v.y); acc += q[i * 8 + 2] * f.x + q[i * 8 + 3] * f.y;
            f = unpack2(v.z); acc += q[i * 8 + 4] * f.x + q[i * 8 + 5] * f.y;
            f = unpack2(v.w); acc += q[i * 8 + 6] * f.x + q[i * 8 + 7] * f.y;
        }
        acc *= 0.125f;
        sc[j] = acc;
        m = fmaxf(m, acc);
    }
    float l = 0.f;
    for (int j = 0; j < T; j++) { sc[j] = __expf(sc[j] - m); l += sc[j]; }
    const float inv = 1.0f / l;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; i++) o[i] = 0.f;
    for (int j = 0; j < T; j++) {
        const uint4* vp = reinterpret_cast<const uint4*>(base + (int64_t)j * frame_stride + 2 * C);
        const float pj = sc[j] * inv;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = vp[i];
            float2 f;
            f = unpack2(v.x); o[i * 8 + 0] += pj * f.x; o[i * 8 + 1] += pj * f.y;
            f = unpack2(v.y); o[i * 8 + 2] += pj * f.x; o[i * 8 + 3] += pj * f.y;
            f = unpack2(v.z); o[i * 8 + 4] += pj * f.x; o[i * 8 + 5] += pj * f.y;
            f = unpack2(v.w); o[i * 8 + 6] += pj * f.x; o[i * 8 + 7] += pj * f.y;
        }
    }
    uint4* op = reinterpret_cast<uint4*>(out + (((int64_t)b * T + t) * tokens + s) * C + h * 64);
#pragma unroll
    for (int i = 0; i < 8; i++)
        op[i] = make_uint4(pack2(o[i * 8 + 0], o[i * 8 + 1]), pack2(o[i * 8 + 2], o[i * 8 + 3]),
                           pack2(o[i * 8 + 4], o[i * 8 + 5]), pack2(o[i * 8 + 6], o[i * 8 + 7]));
}

extern "C" int gcd_attention_temporal(const void* qkv, int clips, int T, int tokens, int heads, void* out, void* stream) {
    GCD_REQUIRE(T >= 1 && T <= 32, "attention_temporal: T=%d unsupported (1..32)", T);
    const int GS = T <= 16 ? 16 : 32;
    const int64_t nthreads = (int64_t)clips * tokens * heads * GS;
    const int64_t blocks = (nthreads + 127) / 128;
    GCD_REQUIRE(blocks < (1ll << 31), "attention_temporal: problem too large");
    if (GS == 16)
        attn_temporal_kernel<16><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>((const act_t*)qkv, clips, T, tokens,
                                                                                     heads, (act_t*)out);
    else
        attn_temporal_kernel<32><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>((const act_t*)qkv, clips, T, tokens,
                                                                                     heads, (act_t*)out);
    GCD_CUDA_CHECK(cudaGetLastError());
    g_launches++;
    return 0;
}
