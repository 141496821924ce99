// gcd_b200 — FlashAttention-style spatial self-attention for sm_100a, head_dim 64, tcgen05 + TMEM + TMA.
// Reference math: gcd-model/sgm/modules/attention.py:283-344 (CrossAttention.forward, context=None):
//   out = softmax(q k^T / sqrt(64)) v, heads outer in the channel dim ("b n (h d) -> b h n d").
//
// One CTA per (128-query tile, head, frame); two CTAs co-reside per SM so one CTA's softmax overlaps the other's MMAs.
//   warp 0    : TMA producer — Q tile once, then K_j / V_j (64 keys each) through a 3-stage mbarrier ring.
//   warp 1    : MMA issuer   — S_j = Q K_j^T  (M128 N64 K64, both operands K-major)            -> TMEM S[j&1]
//                              O  += P_j V_j  (M128 N64 K64, P K-major from smem, V MN-major)    -> TMEM O (accumulates)
//   warp 2    : TMEM allocator (256 columns: S0 S1 O)
//   warps 4-7 : softmax — thread r owns query row r: tcgen05.ld S row, online softmax in fp32 in the exp2 domain with
//               packed f32x2 FMA/ADD and 3-input max, P written 128B-swizzled to smem as the next MMA's A operand.
//               O stays in TMEM: it is rescaled (tcgen05.ld / st) only when a row's running max grew by more than 2^8
//               ("lazy rescale"); otherwise the stale max is kept, which is exact after the final 1/l normalisation.
// Roofline note: per S element 256 tensor FLOPs vs one MUFU ex2 (16/clk/SM) -> MUFU-bound at ~half the bf16 peak
// unless part of the exponentials is evaluated on the FMA pipe (EMU_EVERY below).
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
constexpr float RESCALE_TAU = 8.0f;          // log2 units
#ifndef GCD_FA_EMU_EVERY
#define GCD_FA_EMU_EVERY 0                   // 0: all exponentials on MUFU; n: every n-th pair on the FMA pipe
#endif

struct Params {
    int tokens, heads, nblk;
    act_t* out;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 2^x for a pair on the FMA pipe: Cody-Waite split (round-to-nearest via the 1.5*2^23 trick) + degree-4 polynomial on
// [-0.5, 0.5] (rel. err ~4e-6, far below the 16-bit P rounding), exponent inserted with integer adds.
__device__ __forceinline__ uint64_t ex2_emu2(uint64_t x) {
    const uint64_t MAGIC = pk2(12582912.0f, 12582912.0f), NMAGIC = pk2(-12582912.0f, -12582912.0f);
    float x0, x1;
    upk2(x, x0, x1);
    x = pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
    const uint64_t t = add2(x, MAGIC);                 // integer part in the low mantissa bits
    const uint64_t n = add2(t, NMAGIC);                // rounded x
    const uint64_t f = fma2(n, pk2(-1.0f, -1.0f), x);
    uint64_t p = fma2(f, pk2(9.6181291e-3f, 9.6181291e-3f), pk2(5.5504109e-2f, 5.5504109e-2f));
    p = fma2(p, f, pk2(2.4022651e-1f, 2.4022651e-1f));
    p = fma2(p, f, pk2(6.9314718e-1f, 6.9314718e-1f));
    p = fma2(p, f, pk2(1.0f, 1.0f));
    float p0, p1, t0, t1;
    upk2(p, p0, p1);
    upk2(t, t0, t1);
    p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
    return pk2(p0, p1);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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
    uint64_t* pv_done = bars + 13;     // 1: phase j completes when P_j V_j has been accumulated into O
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

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
        }
        mbar_init(pv_done, 1);
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
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);      // P_j staged and O rescaled if needed
                tc_fence_after();
                const uint64_t pd = make_desc_sw128(smem_u32(sP + (j & 1) * P_BYTES), 16, 1024);
                // V tile [64 keys][64 d]: MN-major B operand; 16 keys (2 groups of 8 rows, SBO=1024) per K step
                const uint64_t vd = make_desc_sw128(smem_u32(sV + s * KV_BYTES), 1024, 1024);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma_f16_ss(tmem_base + 128, pd + (uint64_t)(k * 2), vd + (uint64_t)(k * 128), idesc_pv, (j | k) != 0);
                umma_commit(pv_done);
                umma_commit(&kv_empty[s]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        // ---------------------------------------------------- softmax warps
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
        const uint64_t SC2 = pk2(SCALE_LOG2, SCALE_LOG2);
        float m_run = -INFINITY, l_run = 0.f;

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
            if (kv_left < BK) {
#pragma unroll
                for (int i = 0; i < BK; i++)
                    if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
            }
            float mx = max3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
#pragma unroll
            for (int i = 3; i + 1 < BK; i += 2) mx = max3(mx, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
            mx = fmaxf(mx, __uint_as_float(sv[BK - 1]));
            const float m_new = fmaxf(m_run, mx * SCALE_LOG2);

            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);               // O holds blocks 0..j-1; P buffer (j&1) is free
                if (__any_sync(0xffffffffu, m_new - m_run > RESCALE_TAU)) {
                    tc_fence_after();
                    const float alpha = ex2(m_run - m_new);
                    const uint64_t A2 = pk2(alpha, alpha);
#pragma unroll
                    for (int hlf = 0; hlf < 2; hlf++) {
                        uint32_t v[32];
                        tmem_ld32(lane_addr + 128 + hlf * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float a, b;
                            upk2(mul2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), A2), a, b);
                            v[i] = __float_as_uint(a); v[i + 1] = __float_as_uint(b);
                        }
                        tmem_st32(lane_addr + 128 + hlf * 32, v);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    l_run *= alpha;
                    m_run = m_new;
                }
            } else {
                m_run = m_new;
            }
            // p = 2^(s*scale - m_run), row sum, P row -> smem (128B swizzle: 16B chunk index XOR (row & 7))
            const uint64_t NM2 = pk2(-m_run, -m_run);
            uint64_t lsum2 = pk2(0.f, 0.f);
            uint8_t* prow = sP + (j & 1) * P_BYTES + r * 128;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int e = c * 8 + 2 * i;
                    const uint64_t x = fma2(pk2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), SC2, NM2);
                    uint64_t pe;
#if GCD_FA_EMU_EVERY > 0
                    if (((c * 4 + i) % GCD_FA_EMU_EVERY) == 0) {
                        pe = ex2_emu2(x);
                    } else
#endif
                    {
                        float x0, x1;
                        upk2(x, x0, x1);
                        pe = pk2(ex2(x0), ex2(x1));
                    }
                    lsum2 = add2(lsum2, pe);
                    float p0, p1;
                    upk2(pe, p0, p1);
                    w[i] = pack2(p0, p1);
                }
                *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            {
                float a, b;
                upk2(lsum2, a, b);
                l_run += a + b;
            }
            fence_proxy_async_smem();
            mbar_arrive(&p_full[j & 1]);
        }
        // ---- epilogue: O / l
        mbar_wait(pv_done, (nblk - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l_run;
        const bool valid = q0 + r < p.tokens;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((int64_t)frame * p.tokens + q0 + r) * C + head * D);
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
            uint32_t v[32];
            tmem_ld32(lane_addr + 128 + hlf * 32, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    dst[hlf * 4 + c] = make_uint4(
                        pack2(__uint_as_float(v[c * 8 + 0]) * inv, __uint_as_float(v[c * 8 + 1]) * inv),
                        pack2(__uint_as_float(v[c * 8 + 2]) * inv, __uint_as_float(v[c * 8 + 3]) * inv),
                        pack2(__uint_as_float(v[c * 8 + 4]) * inv, __uint_as_float(v[c * 8 + 5]) * inv),
                        pack2(__uint_as_float(v[c * 8 + 6]) * inv, __uint_as_float(v[c * 8 + 7]) * inv));
            }
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
