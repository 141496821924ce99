// gcd_b200 — tcgen05 implicit-GEMM kernel (Linear / Conv2d 3x3 & 1x1 / Conv3d (3,1,1) / batched matmul).
//
// One persistent, warp-specialised kernel per output-tile width BN in {128,160,256}:
//   warp 0   : TMA producer  — per (tap, 64-channel chunk): one 4-D box of the channels-last activation tensor
//              (out-of-range coordinates zero-filled by TMA = conv zero padding) + one 2-D/3-D box of the weights,
//              both landing 128B-swizzled in shared memory, signalled on an mbarrier ring.
//   warp 1   : MMA issuer    — a single thread issues tcgen05.mma (M=128, N=BN, K=16, fp32 accumulate in TMEM),
//              tcgen05.commit frees smem stages and publishes the accumulator.
//   warp 2   : TMEM allocator (512 columns = 2 accumulator stages so the epilogue overlaps the next tile's MMAs).
//   warps 4-7: epilogue      — tcgen05.ld accumulator rows, fused bias / per-frame vector / activation / GEGLU /
//              residual blend, vectorised global stores.
// Replaces cuDNN/cuBLAS calls behind nn.Conv2d / nn.Conv3d / nn.Linear in
//   gcd-model/sgm/modules/diffusionmodules/openaimodel.py:213-357 (ResBlock), :110-210 (Up/Downsample),
//   gcd-model/sgm/modules/attention.py:87-113,255-344 (FeedForward/GEGLU, CrossAttention projections),
//   gcd-model/sgm/modules/diffusionmodules/model.py:94-201 (VAE ResnetBlock/AttnBlock).
#include "common.cuh"
#include "../../include/gcd_b200.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

using namespace ptx;

// ------------------------------------------------------------------------------------------------ host globals
static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};
void gcd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* gcd_last_error(void) { return g_err; }
extern "C" int gcd_version(void) { return 100; }
extern "C" int gcd_act_dtype(void) { return (int)GCD_UMMA_FMT; }
extern "C" int64_t gcd_launch_count(void) { return g_launches.load(); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}
int gcd_make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides, int swizzle128) {
    PFN_encodeTiled enc = get_encode();
    GCD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; i++) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = elem_strides ? elem_strides[i] : 1;
    }
    for (int i = 0; i + 1 < rank; i++) s[i] = strides_bytes[i];
#ifdef GCD_ACT_BF16
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
    CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GCD_REQUIRE(r == CUDA_SUCCESS,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u stride0 %llu", (int)r,
                rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
                (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), b[0],
                rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0, (unsigned long long)s[0]);
    return 0;
}

// ------------------------------------------------------------------------------------------------ kernel
struct TcParams {
    int ntx, nty, ntz;          // M-tiles along x, y, z
    int lTW, lTH;               // log2 tile extents (TW*TH*TN == 128)
    int Xo, Yo, Zo;             // output extents
    int in_mul;
    int ntaps, kchunks;
    int8_t tdx[9], tdy[9], tdz[9];
    int N, n_tiles;
    int w_batched;
    // epilogue
    const float* bias;
    const float* rowvec;
    int rpv, ldv;
    const void* r1;
    const void* r2;
    int ld1, ld2, r1f32, r2f32;
    float a0, a1, a2;
    void* out;
    int ldo, of32, geglu, act, vec_ok;
};

template <int BN>
struct TcCfg {
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGES = (BN == 256) ? 4 : 5;
    static constexpr int EPI_LD = 36;                       // floats per staged row (32 + 4 pad: conflict-free)
    static constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;   // one 32x32 fp32 chunk per epilogue warp
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---- coalesced epilogue helpers: a lane owns 4 consecutive output columns of one row -------------------------
__device__ __forceinline__ void load_res4(float* x, const void* R, int64_t off, bool f32, float a, int nv, bool vec) {
    if (f32) {
        const float* r = reinterpret_cast<const float*>(R) + off;
        if (vec) {
            float4 t = *reinterpret_cast<const float4*>(r);
            x[0] += a * t.x; x[1] += a * t.y; x[2] += a * t.z; x[3] += a * t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (j < nv) x[j] += a * r[j];
        }
    } else {
        const act_t* r = reinterpret_cast<const act_t*>(R) + off;
        if (vec) {
            uint2 t = *reinterpret_cast<const uint2*>(r);
            float2 f0 = unpack2(t.x), f1 = unpack2(t.y);
            x[0] += a * f0.x; x[1] += a * f0.y; x[2] += a * f1.x; x[3] += a * f1.y;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (j < nv) x[j] += a * act2f(r[j]);
        }
    }
}
__device__ __forceinline__ void store4(void* out, int64_t off, bool f32, const float* x, int nv, bool vec) {
    if (f32) {
        float* o = reinterpret_cast<float*>(out) + off;
        if (vec) *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (j < nv) o[j] = x[j];
        }
    } else {
        act_t* o = reinterpret_cast<act_t*>(out) + off;
        if (vec) *reinterpret_cast<uint2*>(o) = make_uint2(pack2(x[0], x[1]), pack2(x[2], x[3]));
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (j < nv) o[j] = f2act(x[j]);
        }
    }
}

// One 32-accumulator-column chunk of one warp's 32 rows, read back from the staging tile in coalesced order.
// LPR lanes cover one row (4 output columns per lane); LPR = 8 (plain) or 4 (GEGLU: 16 value + 16 gate columns).
// Phase 1 issues every residual load of the chunk (memory-level parallelism), phase 2 computes and stores.
template <int LPR>
__device__ __forceinline__ void epi_chunk(const TcParams& p, const float* stg, int ld, int q, int lane, int nbase,
                                          int tx, int ty, int tz, int TW, int TH, int TN) {
    constexpr int RPS = 32 / LPR, STEPS = LPR;
    constexpr bool GEGLU = (LPR == 4);
    const int c4 = (lane % LPR) * 4;
    const int rsub = lane / LPR;
    const int Nout = GEGLU ? (p.N >> 1) : p.N;
    const int n = nbase + c4;                                  // accumulator (value) column of this lane
    const int ncol = GEGLU ? (nbase >> 1) + c4 : n;            // output column
    const int nv = Nout - ncol;
    if (nv <= 0) return;
    const bool vec = p.vec_ok && nv >= 4;
    float bv[4] = {0.f, 0.f, 0.f, 0.f}, bg[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (n + j < p.N) bv[j] = __ldg(p.bias + n + j);
            if (GEGLU) bg[j] = __ldg(p.bias + n + 16 + j);
        }
    }
    int64_t rows[STEPS];
    float r1v[STEPS][4], r2v[STEPS][4];
#pragma unroll
    for (int s = 0; s < STEPS; s++) {
        const int r = q * 32 + s * RPS + rsub;
        const int x = tx * TW + (r & (TW - 1));
        const int y = ty * TH + ((r >> p.lTW) & (TH - 1));
        const int z = tz * TN + (r >> (p.lTW + p.lTH));
        const bool valid = (x < p.Xo) && (y < p.Yo) && (z < p.Zo);
        rows[s] = valid ? ((int64_t)z * p.Yo + y) * p.Xo + x : -1;
#pragma unroll
        for (int j = 0; j < 4; j++) { r1v[s][j] = 0.f; r2v[s][j] = 0.f; }
        if (valid) {
            if (p.r1) load_res4(r1v[s], p.r1, rows[s] * p.ld1 + ncol, p.r1f32, p.a1, nv, vec);
            if (p.r2) load_res4(r2v[s], p.r2, rows[s] * p.ld2 + ncol, p.r2f32, p.a2, nv, vec);
        }
    }
#pragma unroll
    for (int s = 0; s < STEPS; s++) {
        if (rows[s] < 0) continue;
        const int rl = s * RPS + rsub;
        float xf[4];
        {
            float4 t = *reinterpret_cast<const float4*>(stg + rl * ld + c4);
            xf[0] = t.x + bv[0]; xf[1] = t.y + bv[1]; xf[2] = t.z + bv[2]; xf[3] = t.w + bv[3];
        }
        const float* rv = p.rowvec ? p.rowvec + (rows[s] / p.rpv) * (int64_t)p.ldv + n : nullptr;
        if (rv) {
#pragma unroll
            for (int j = 0; j < 4; j++) if (n + j < p.N) xf[j] += __ldg(rv + j);
        }
        if (GEGLU) {
            float4 g = *reinterpret_cast<const float4*>(stg + rl * ld + 16 + c4);
            float gg[4] = {g.x + bg[0], g.y + bg[1], g.z + bg[2], g.w + bg[3]};
            if (rv) {
#pragma unroll
                for (int j = 0; j < 4; j++) gg[j] += __ldg(rv + 16 + j);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) xf[j] *= gelu_erf(gg[j]);
        }
        if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 4; j++) xf[j] = silu(xf[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) xf[j] = p.a0 * xf[j] + r1v[s][j] + r2v[s][j];
        store4(p.out, rows[s] * p.ldo + ncol, p.of32, xf, nv, vec);
    }
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
    using Cfg = TcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
    float* sEpi = reinterpret_cast<float*>(smem + Cfg::STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::EPI_BYTES);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
    uint64_t* tfull = bars + 2 * Cfg::STAGES;    // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < Cfg::STAGES; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = p.ntx * p.nty * p.ntz;
    const int total = m_tiles * p.n_tiles;
    const int kiters = p.ntaps * p.kchunks;
    const int TW = 1 << p.lTW, TH = 1 << p.lTH;
    const int TN = 128 >> (p.lTW + p.lTH);

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                const int mt = tile / p.n_tiles;
                const int tx = mt % p.ntx;
                const int ty = (mt / p.ntx) % p.nty;
                const int tz = mt / (p.ntx * p.nty);
                const int x0 = tx * TW * p.in_mul, y0 = ty * TH * p.in_mul, z0 = tz * TN;
                const int wy = p.w_batched ? ty : 0;
                int kidx = 0;
                for (int tap = 0; tap < p.ntaps; tap++) {
                    const int cx = x0 + p.tdx[tap], cy = y0 + p.tdy[tap], cz = z0 + p.tdz[tap];
                    for (int kc = 0; kc < p.kchunks; kc++, kidx += 64) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
                        tma_load_4d(&mapA, sA + stage * Cfg::A_BYTES, &full[stage], kc * 64, cx, cy, cz);
                        tma_load_3d(&mapB, sB + stage * Cfg::B_BYTES, &full[stage], kidx, nt * BN, wy);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc_f16(128, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, it++) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * 256;
                for (int ki = 0; ki < kiters; ki++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t ad = make_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES), 16, 1024);
                    const uint64_t bd = make_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; k++)   // 4 x (K=16) inside the 64-wide swizzle atom: +32 B each
                        umma_f16_ss(tacc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (ki | k) != 0);
                    umma_commit(&empty[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[as]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // TMEM rows arrive one-per-thread; each 32-column chunk is transposed through a per-warp smem tile so that
        // global loads/stores are coalesced: a lane then owns 4 consecutive columns, LPR lanes cover one row.
        const int q = warp & 3;              // TMEM lane quadrant of this warp
        float* stg = sEpi + q * 32 * Cfg::EPI_LD;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, it++) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int nt = tile % p.n_tiles;
            const int mt = tile / p.n_tiles;
            const int tx = mt % p.ntx;
            const int ty = (mt / p.ntx) % p.nty;
            const int tz = mt / (p.ntx * p.nty);
            const int n0 = nt * BN;

            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                if (n0 + c0 >= p.N) continue;     // warp-uniform
                {
                    float4* dst = reinterpret_cast<float4*>(stg + lane * Cfg::EPI_LD);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
                __syncwarp();
                if (p.geglu) epi_chunk<4>(p, stg, Cfg::EPI_LD, q, lane, n0 + c0, tx, ty, tz, TW, TH, TN);
                else         epi_chunk<8>(p, stg, Cfg::EPI_LD, q, lane, n0 + c0, tx, ty, tz, TW, TH, TN);
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host launcher
static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

template <int BN>
static int launch_tc(const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        GCD_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        int dev = 0;
        GCD_CUDA_CHECK(cudaGetDevice(&dev));
        GCD_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const int total = p.ntx * p.nty * p.ntz * p.n_tiles;
    const int grid = total < num_sms ? total : num_sms;
    tc_gemm_kernel<BN><<<grid, 256, Cfg::SMEM, st>>>(mA, mB, p);
    GCD_CUDA_CHECK(cudaGetLastError());
    g_launches++;
    return 0;
}

extern "C" int gcd_tc_run(const gcd_tc_op* op, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GCD_REQUIRE(op && op->A && op->W && op->ep.out, "gcd_tc_run: null pointer");
    GCD_REQUIRE(op->ntaps >= 1 && op->ntaps <= 9, "gcd_tc_run: ntaps %d out of range", op->ntaps);
    GCD_REQUIRE(op->C % 8 == 0 && op->C > 0, "gcd_tc_run: C=%d must be a positive multiple of 8", op->C);
    GCD_REQUIRE(op->sx % 8 == 0 && op->sy % 8 == 0 && op->sz % 8 == 0 && op->ldw % 8 == 0,
                "gcd_tc_run: strides must be multiples of 8 elements (16 B)");
    GCD_REQUIRE(((uintptr_t)op->A & 15) == 0 && ((uintptr_t)op->W & 15) == 0, "gcd_tc_run: operands must be 16B aligned");
    GCD_REQUIRE(op->in_mul == 1 || op->in_mul == 2, "gcd_tc_run: in_mul must be 1 or 2");
    const gcd_epilogue& e = op->ep;
    GCD_REQUIRE(!e.geglu || (op->N % 32 == 0), "gcd_tc_run: GEGLU needs N %% 32 == 0");
    GCD_REQUIRE(!e.rowvec || e.rows_per_vec > 0, "gcd_tc_run: rows_per_vec must be > 0");

    TcParams p;
    memset(&p, 0, sizeof(p));
    // ---- tile search
    int TW, TH, TN;
    if (op->gemm_tile) {
        TW = 128; TH = 1; TN = 1;
    } else {
        TW = 1;
        while (TW < 128 && op->Xo % (TW * 2) == 0) TW *= 2;
        long best = -1;
        int bestTH = 1;
        for (int th = 1; th * TW <= 128; th *= 2) {
            int tn = 128 / (TW * th);
            long cov = (long)((op->Yo + th - 1) / th) * th * (long)((op->Zo + tn - 1) / tn) * tn;
            if (best < 0 || cov <= best) { best = cov; bestTH = th; }
        }
        TH = bestTH;
        TN = 128 / (TW * TH);
    }
    p.lTW = ilog2(TW);
    p.lTH = ilog2(TH);
    p.ntx = (op->Xo + TW - 1) / TW;
    p.nty = (op->Yo + TH - 1) / TH;
    p.ntz = (op->Zo + TN - 1) / TN;
    p.Xo = op->Xo; p.Yo = op->Yo; p.Zo = op->Zo;
    p.in_mul = op->in_mul;
    p.ntaps = op->ntaps;
    p.kchunks = (op->C + 63) / 64;
    for (int i = 0; i < op->ntaps; i++) { p.tdx[i] = op->tap_dx[i]; p.tdy[i] = op->tap_dy[i]; p.tdz[i] = op->tap_dz[i]; }
    p.N = op->N;
    p.w_batched = op->w_batch_stride != 0;
    GCD_REQUIRE(!p.w_batched || TH == 1, "gcd_tc_run: batched weights need gemm tiling");

    int BN;
    if (op->N % 256 == 0) BN = 256;
    else if (op->N % 160 == 0) BN = 160;
    else BN = 128;
    p.n_tiles = (op->N + BN - 1) / BN;

    p.bias = e.bias; p.rowvec = e.rowvec; p.rpv = e.rows_per_vec; p.ldv = e.ld_rowvec;
    p.r1 = e.res1; p.r2 = e.res2; p.ld1 = e.ld_res1; p.ld2 = e.ld_res2; p.r1f32 = e.res1_f32; p.r2f32 = e.res2_f32;
    p.a0 = e.a_acc; p.a1 = e.a_res1; p.a2 = e.a_res2;
    p.out = e.out; p.ldo = e.ld_out; p.of32 = e.out_f32; p.geglu = e.geglu; p.act = e.act;
    // vector path: 16-column runs must be 16B aligned for every tensor touched
    auto ok = [](const void* ptr, int ld, int f32) {
        if (!ptr) return true;
        int a = f32 ? 4 : 8;
        return (ld % a == 0) && (((uintptr_t)ptr & 15) == 0);
    };
    p.vec_ok = ok(e.out, e.ld_out, e.out_f32) && ok(e.res1, e.ld_res1, e.res1_f32) && ok(e.res2, e.ld_res2, e.res2_f32) &&
               (!e.bias || ((uintptr_t)e.bias & 15) == 0);

    // ---- tensor maps
    CUtensorMap mA, mB;
    {
        uint64_t dims[4] = {(uint64_t)op->C, (uint64_t)op->Xi, (uint64_t)op->Yi, (uint64_t)op->Zi};
        uint64_t str[3] = {(uint64_t)op->sx * 2, (uint64_t)op->sy * 2, (uint64_t)op->sz * 2};
        uint32_t box[4] = {64, (uint32_t)(TW * op->in_mul), (uint32_t)(TH * op->in_mul), (uint32_t)TN};
        uint32_t es[4] = {1, (uint32_t)op->in_mul, (uint32_t)op->in_mul, 1};
        if (op->in_mul == 2) {   // box spans 2*T-1 source elements -> exactly T samples
            box[1] = 2 * TW - 1 > 0 ? 2 * TW - 1 : 1;
            box[2] = 2 * TH - 1 > 0 ? 2 * TH - 1 : 1;
            if (TW == 1) es[1] = 1;
            if (TH == 1) es[2] = 1;
        }
        // degenerate extents: TMA needs stride > 0 & multiple of 16B even when extent is 1
        for (int i = 0; i < 3; i++) if (str[i] == 0) str[i] = 16;
        int rc = gcd_make_tmap(&mA, op->A, 4, dims, str, box, es, 1);
        if (rc) return rc;
    }
    {
        uint64_t nb = p.w_batched ? (uint64_t)op->Yo : 1;
        uint64_t dims[3] = {(uint64_t)op->ntaps * (uint64_t)(p.kchunks * 64), (uint64_t)op->N, nb};
        // K extent: the weight matrix must really have ntaps*kchunks*64 columns unless C%64==0 holds (host packs it so)
        dims[0] = (uint64_t)op->ntaps * (uint64_t)op->C;
        GCD_REQUIRE(op->C % 64 == 0 || op->ntaps == 1, "gcd_tc_run: multi-tap ops need C %% 64 == 0 (got %d)", op->C);
        uint64_t str[2] = {(uint64_t)op->ldw * 2, (uint64_t)(p.w_batched ? op->w_batch_stride : op->ldw * (int64_t)op->N) * 2};
        uint32_t box[3] = {64, (uint32_t)BN, 1};
        int rc = gcd_make_tmap(&mB, op->W, 3, dims, str, box, nullptr, 1);
        if (rc) return rc;
    }
    switch (BN) {
        case 256: return launch_tc<256>(mA, mB, p, st);
        case 160: return launch_tc<160>(mA, mB, p, st);
        default: return launch_tc<128>(mA, mB, p, st);
    }
}
