// gcd_b200 — FlashAttention-style spatial self-attention for sm_100a, head_dim 64, tcgen05 + TMEM + TMA.
// Reference math: gcd-model/sgm/modules/attention.py:283-344 (CrossAttention.forward, context=None):
//   out = softmax(q k^T / sqrt(64)) v, heads outer in the channel dim ("b n (h d) -> b h n d").
//
// One CTA per (128-query tile, head, frame); two CTAs co-reside per SM so one CTA's softmax overlaps the other's MMAs.
//   warp 0    : TMA producer — Q tile once, then K_j / V_j (64 keys each) through a 3-stage mbarrier ring.
//   warp 1    : MMA issuer   — S_j = Q K_j^T  (M128 N64 K64, both operands K-major)   -> TMEM S[j&1]
//                              O_j = P_j V_j  (M128 N64 K64, P K-major from smem, V MN-major) -> TMEM O[j&1]
//   warp 2    : TMEM allocator (256 columns: S0 S1 O0 O1)
//   warps 4-7 : softmax — thread r owns query row r: tcgen05.ld S row, online softmax in fp32 (exp2 domain),
//               P written 128B-swizzled to smem as the next MMA's A operand, partial O accumulated in registers
//               with the running rescale; final normalise and store.
// Roofline note: per S element 256 tensor FLOPs vs one MUFU ex2 (16/clk/SM) -> MUFU-bound at ~half the bf16 peak.
#include "common.cuh"
#include "../../include/gcd_b200.h"
#include <atomic>
extern std::atomic<int64_t> g_launches;
using namespace ptx;

namespace fa {
constexpr int BQ = 128, BK = 64, D = 64;
constexpr int KV_STAGES = 3;
constexpr int Q_BYTES = BQ * D * 2;          // 16 KB
constexpr int KV_BYTES = BK * D * 2;         // 8 KB each for K and V
constexpr int P_BYTES = BQ * BK * 2;         // 16 KB
constexpr int SMEM = Q_BYTES + KV_STAGES * 2 * KV_BYTES + 2 * P_BYTES + 1024 + 256;

struct Params {
    int tokens, heads, nblk;
    act_t* out;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(256, 2)
attn_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Q_BYTES;                         // [stage][8 KB]
    uint8_t* sV = sK + KV_STAGES * KV_BYTES;            // [stage][8 KB]
    uint8_t* sP = sV + KV_STAGES * KV_BYTES;            // [2][16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
    uint64_t* bar_q = bars;            // 1
    uint64_t* kv_full = bars + 1;      // 3
    uint64_t* kv_empty = bars + 4;     // 3
    uint64_t* s_full = bars + 7;       // 2
    uint64_t* s_empty = bars + 9;      // 2
    uint64_t* p_full = bars + 11;      // 2
    uint64_t* o_full = bars + 13;      // 2
    uint64_t* o_empty = bars + 15;     // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, head = blockIdx.y, frame = blockIdx.z;
    const int C = p.heads * D;
    const int nblk = p.nblk;

    if (warp == 0 && lane == 0) { prefetch_tmap(&mapQ); prefetch_tmap(&mapKV); }
    if (warp == 1 && lane == 0) {
        mbar_init(bar_q, 1);
        for (int i = 0; i < KV_STAGES; i++) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && lane == 0) {
            // ------------------------------------------------ TMA producer
            mbar_expect_tx(bar_q, Q_BYTES);
            tma_load_3d(&mapQ, sQ, bar_q, head * D, q0, frame);
            for (int j = 0; j < nblk; j++) {
                const int s = j % KV_STAGES;
                mbar_wait(&kv_empty[s], ((j / KV_STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], 2 * KV_BYTES);
                tma_load_3d(&mapKV, sK + s * KV_BYTES, &kv_full[s], C + head * D, j * BK, frame);
                tma_load_3d(&mapKV, sV + s * KV_BYTES, &kv_full[s], 2 * C + head * D, j * BK, frame);
            }
        } else if (warp == 1 && lane == 0) {
            // ------------------------------------------------ MMA issuer
            constexpr uint32_t idesc_qk = make_idesc_f16(128, BK, 0, 0);   // A, B K-major
            constexpr uint32_t idesc_pv = make_idesc_f16(128, D, 0, 1);    // B (=V) MN-major
            const uint64_t qd = make_desc_sw128(smem_u32(sQ), 16, 1024);
            auto issue_qk = [&](int j) {
                const int s = j % KV_STAGES;
                mbar_wait(&kv_full[s], (j / KV_STAGES) & 1);
                mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint64_t kd = make_desc_sw128(smem_u32(sK + s * KV_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma_f16_ss(tmem_base + (j & 1) * 64, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_qk, k != 0);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(bar_q, 0);
            issue_qk(0);
            for (int j = 0; j < nblk; j++) {
                if (j + 1 < nblk) issue_qk(j + 1);
                const int s = j % KV_STAGES;
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                mbar_wait(&o_empty[j & 1], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint64_t pd = make_desc_sw128(smem_u32(sP + (j & 1) * P_BYTES), 16, 1024);
                // V tile [64 keys][64 d]: MN-major B operand; 16 keys (2 groups of 8 rows, SBO=1024) per K step
                const uint64_t vd = make_desc_sw128(smem_u32(sV + s * KV_BYTES), 1024, 1024);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma_f16_ss(tmem_base + 128 + (j & 1) * 64, pd + (uint64_t)(k * 2), vd + (uint64_t)(k * 128), idesc_pv,
                                k != 0);
                umma_commit(&o_full[j & 1]);
                umma_commit(&kv_empty[s]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        // ---------------------------------------------------- softmax / accumulate / epilogue
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
        float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
        float o[D];
#pragma unroll
        for (int i = 0; i < D; i++) o[i] = 0.f;

        auto accumulate_o = [&](int jb, float alpha) {
            mbar_wait(&o_full[jb & 1], (jb >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int hlf = 0; hlf < 2; hlf++) {
                uint32_t v[32];
                tmem_ld32(lane_addr + 128 + (jb & 1) * 64 + hlf * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i++) o[hlf * 32 + i] = o[hlf * 32 + i] * alpha + __uint_as_float(v[i]);
            }
            tc_fence_before();
            mbar_arrive(&o_empty[jb & 1]);
        };

        for (int j = 0; j < nblk; j++) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sv[64];
            tmem_ld32(lane_addr + (j & 1) * 64, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld32(lane_addr + (j & 1) * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_empty[j & 1]);

            const int kv_left = p.tokens - j * BK;   // valid keys in this block
            float mx = -INFINITY;
            if (kv_left >= BK) {
#pragma unroll
                for (int i = 0; i < BK; i++) mx = fmaxf(mx, __uint_as_float(sv[i]));
            } else {
#pragma unroll
                for (int i = 0; i < BK; i++) {
                    if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
                    mx = fmaxf(mx, __uint_as_float(sv[i]));
                }
            }
            const float m_new = fmaxf(m_run, mx * SCALE_LOG2);
            const float alpha = ex2(m_run - m_new);
            m_run = m_new;
            float lsum = 0.f;
            // P row -> smem (128B swizzle: 16B chunk index XOR (row & 7)), as fp16/bf16 pairs
            uint8_t* prow = sP + (j & 1) * P_BYTES + r * 128;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    e[i] = ex2(__uint_as_float(sv[c * 8 + i]) * SCALE_LOG2 - m_new);
                    lsum += e[i];
                }
                *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) =
                    make_uint4(pack2(e[0], e[1]), pack2(e[2], e[3]), pack2(e[4], e[5]), pack2(e[6], e[7]));
            }
            l_run = l_run * alpha + lsum;
            fence_proxy_async_smem();
            mbar_arrive(&p_full[j & 1]);
            if (j >= 1) accumulate_o(j - 1, alpha_prev);
            alpha_prev = alpha;
        }
        accumulate_o(nblk - 1, alpha_prev);
        if (q0 + r < p.tokens) {
            const float inv = 1.0f / l_run;
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((int64_t)frame * p.tokens + q0 + r) * C + head * D);
#pragma unroll
            for (int c = 0; c < 8; c++)
                dst[c] = make_uint4(pack2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv), pack2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv),
                                    pack2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv), pack2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}
}  // namespace fa

extern "C" int gcd_attention_spatial(const void* qkv, int frames, int tokens, int heads, void* out, void* stream) {
    GCD_REQUIRE(qkv && out && frames > 0 && tokens > 0 && heads > 0, "attention_spatial: bad arguments");
    GCD_REQUIRE(frames <= 65535 && heads <= 65535, "attention_spatial: grid too large");
    static bool configured = false;
    if (!configured) {
        GCD_CUDA_CHECK(cudaFuncSetAttribute(fa::attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fa::SMEM));
        configured = true;
    }
    const int C = heads * 64;
    CUtensorMap mQ, mKV;
    uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)tokens, (uint64_t)frames};
    uint64_t str[2] = {(uint64_t)3 * C * 2, (uint64_t)tokens * 3 * C * 2};
    uint32_t boxq[3] = {64, fa::BQ, 1}, boxk[3] = {64, fa::BK, 1};
    int rc = gcd_make_tmap(&mQ, qkv, 3, dims, str, boxq, nullptr, 128, 0);
    if (rc) return rc;
    rc = gcd_make_tmap(&mKV, qkv, 3, dims, str, boxk, nullptr, 128, 0);
    if (rc) return rc;
    fa::Params p;
    p.tokens = tokens; p.heads = heads; p.nblk = (tokens + fa::BK - 1) / fa::BK; p.out = (act_t*)out;
    dim3 grid((tokens + fa::BQ - 1) / fa::BQ, heads, frames);
    fa::attn_kernel<<<grid, 256, fa::SMEM, (cudaStream_t)stream>>>(mQ, mKV, p);
    GCD_CUDA_CHECK(cudaGetLastError());
    g_launches++;
    return 0;
}
