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
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
};

template <bool OUT_F32>
__device__ __forceinline__ void store_run(void* out, int64_t off, const float* x, int n, bool vec) {
    if (OUT_F32) {
        float* o = reinterpret_cast<float*>(out) + off;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                if (j < n) *reinterpret_cast<float4*>(o + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        } else {
            for (int j = 0; j < n; j++) o[j] = x[j];
        }
    } else {
        act_t* o = reinterpret_cast<act_t*>(out) + off;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 16; j += 8)
                if (j < n)
                    *reinterpret_cast<uint4*>(o + j) = make_uint4(pack2(x[j], x[j + 1]), pack2(x[j + 2], x[j + 3]),
                                                                  pack2(x[j + 4], x[j + 5]), pack2(x[j + 6], x[j + 7]));
        } else {
            for (int j = 0; j < n; j++) o[j] = f2act(x[j]);
        }
    }
}

// add a_res * R[off .. off+16) into x (16 values, n valid)
__device__ __forceinline__ void add_res16(float* x, const void* R, int64_t off, bool f32, float a, int n, bool vec) {
    if (f32) {
        const float* r = reinterpret_cast<const float*>(R) + off;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float4 t = *reinterpret_cast<const float4*>(r + j);
                x[j] += a * t.x; x[j + 1] += a * t.y; x[j + 2] += a * t.z; x[j + 3] += a * t.w;
            }
        } else {
            for (int j = 0; j < n; j++) x[j] += a * r[j];
        }
    } else {
        const act_t* r = reinterpret_cast<const act_t*>(R) + off;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
                uint4 t = *reinterpret_cast<const uint4*>(r + j);
                float2 f;
                f = unpack2(t.x); x[j] += a * f.x; x[j + 1] += a * f.y;
                f = unpack2(t.y); x[j + 2] += a * f.x; x[j + 3] += a * f.y;
                f = unpack2(t.z); x[j + 4] += a * f.x; x[j + 5] += a * f.y;
                f = unpack2(t.w); x[j + 6] += a * f.x; x[j + 7] += a * f.y;
            }
        } else {
            for (int j = 0; j < n; j++) x[j] += a * act2f(r[j]);
        }
    }
}

// Finishes 16 consecutive output columns [ncol, ncol+16) of one row: act, blend with residuals, store.
__device__ __forceinline__ void finish16(const TcParams& p, float* x, int64_t row, int ncol, int nvalid) {
    if (nvalid <= 0) return;
    const bool vec = p.vec_ok && nvalid >= 16;
    if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = silu(x[j]);
    }
    if (p.a0 != 1.0f) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] *= p.a0;
    }
    if (p.r1) add_res16(x, p.r1, row * p.ld1 + ncol, p.r1f32, p.a1, nvalid, vec);
    if (p.r2) add_res16(x, p.r2, row * p.ld2 + ncol, p.r2f32, p.a2, nvalid, vec);
    if (p.of32)
        store_run<true>(p.out, row * p.ldo + ncol, x, nvalid, vec);
    else
        store_run<false>(p.out, row * p.ldo + ncol, x, nvalid, vec);
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
    using Cfg = TcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
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
        const int q = warp & 3;              // TMEM lane quadrant of this warp
        const int r = q * 32 + lane;         // tile row handled by this thread
        const int Nout = p.geglu ? (p.N >> 1) : p.N;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, it++) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int nt = tile % p.n_tiles;
            const int mt = tile / p.n_tiles;
            const int tx = mt % p.ntx;
            const int ty = (mt / p.ntx) % p.nty;
            const int tz = mt / (p.ntx * p.nty);
            const int x = tx * TW + (r & (TW - 1));
            const int y = ty * TH + ((r >> p.lTW) & (TH - 1));
            const int z = tz * TN + (r >> (p.lTW + p.lTH));
            const bool valid = (x < p.Xo) && (y < p.Yo) && (z < p.Zo);
            const int64_t row = ((int64_t)z * p.Yo + y) * p.Xo + x;
            const float* rv = (p.rowvec && valid) ? p.rowvec + (row / p.rpv) * (int64_t)p.ldv : nullptr;
            const int n0 = nt * BN;

            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                const int n = n0 + c0;
                if (valid && n < p.N) {
                    float xf[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) xf[j] = __uint_as_float(v[j]);
                    const int nv = min(32, p.N - n);
                    if (p.bias) {
                        if (nv == 32 && p.vec_ok) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                                xf[j] += b.x; xf[j + 1] += b.y; xf[j + 2] += b.z; xf[j + 3] += b.w;
                            }
                        } else {
                            for (int j = 0; j < nv; j++) xf[j] += p.bias[n + j];
                        }
                    }
                    if (rv) {
                        if (nv == 32 && (p.ldv & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 b = __ldg(reinterpret_cast<const float4*>(rv + n + j));
                                xf[j] += b.x; xf[j + 1] += b.y; xf[j + 2] += b.z; xf[j + 3] += b.w;
                            }
                        } else {
                            for (int j = 0; j < nv; j++) xf[j] += rv[n + j];
                        }
                    }
                    if (p.geglu) {
                        float g[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) g[j] = xf[j] * gelu_erf(xf[16 + j]);
                        finish16(p, g, row, n >> 1, min(16, Nout - (n >> 1)));
                    } else {
                        finish16(p, xf, row, n, min(16, Nout - n));
                        finish16(p, xf + 16, row, n + 16, min(16, Nout - n - 16));
                    }
                }
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
