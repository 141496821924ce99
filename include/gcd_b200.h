/* gcd_b200 — C ABI of the B200-native GCD denoising hot path.
 *
 * The reference (basilevh/gcd, gcd-model/sgm) is pure Python/PyTorch and has NO FFI of its own
 * (SURVEY.md §8(b)); the drop-in boundary upstream is the `instantiate_from_config` plugin surface
 * (gcd-model/sgm/util.py:168-185). This header is the C boundary *underneath* the plugin classes in
 * gcd_b200/ (VideoUNet, Denoiser, EulerEDMSampler, VideoDecoder): each entry point states which reference
 * torch call(s) it replaces. Conventions (all entry points):
 *   - raw DEVICE pointers, caller-owned, never freed or retained by the library;
 *   - plain ints / POD structs, no torch types;  `stream` is a cudaStream_t passed as void*;
 *   - returns 0 on success, negative on error; message via gcd_last_error() (thread-local);
 *   - kernels are asynchronous on `stream`.
 * 16-bit tensors ("act") are IEEE fp16 unless the library was built with -DGCD_ACT_BF16
 * (query with gcd_act_dtype()). Activations are channels-last: [frames, H, W, C] == [rows, C].
 */
#ifndef GCD_B200_H
#define GCD_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* gcd_last_error(void);
int gcd_version(void);
/* 0 = fp16 operands, 1 = bf16 operands */
int gcd_act_dtype(void);
/* number of kernels launched by this library since load (bench.py `gpu_launches`) */
int64_t gcd_launch_count(void);

/* Fused epilogue applied to every tensor-core op (rows r of the output, columns n):
 *   x      = acc[r,n] + bias[n] + rowvec[(r / rows_per_vec) * ld_rowvec + n]
 *   x      = act(x)                      (act: 0 none, 1 SiLU)
 *   geglu: columns come in interleaved blocks of 16 value / 16 gate;  x = value * gelu_erf(gate)
 *          and the output has N/2 columns          (attention.py:87-94 GEGLU)
 *   out    = a_acc * x + a_res1 * res1[r,n] + a_res2 * res2[r,n]
 * which covers: conv bias, `h + emb_out[..., None, None]` (openaimodel.py:349-355), residual adds
 * (openaimodel.py:357, attention.py:551-572), AlphaBlender (util.py:341-369), len-1 cross-attention bias.
 */
typedef struct gcd_epilogue {
    const float* bias;     /* [N] or NULL */
    const float* rowvec;   /* [n_groups, ld_rowvec] or NULL */
    int32_t rows_per_vec;
    int32_t ld_rowvec;
    const void* res1;      /* [rows, ld_res1] or NULL */
    int32_t ld_res1;
    int32_t res1_f32;      /* 1: float32, 0: act */
    const void* res2;
    int32_t ld_res2;
    int32_t res2_f32;
    float a_acc, a_res1, a_res2;
    void* out;             /* [rows, ld_out] */
    int32_t ld_out;
    int32_t out_f32;
    int32_t geglu;
    int32_t act;
    /* Optional fused GroupNorm statistics of the tensor being written (the NEXT GroupNorm's input):
     * gn_stats[(row / gn_rows_per_img) * gn_groups + col / gn_cpg][0..1] += (sum, sum of squares) of the stored values.
     * Zero the buffer first. Only produced when every 128-row output tile lies inside one statistics image and all
     * tiles are full; otherwise gcd_tc_run returns 1 (success, statistics NOT produced: run gcd_groupnorm_stats). */
    double* gn_stats;
    int32_t gn_cpg;
    int32_t gn_groups;
    int64_t gn_rows_per_img;
} gcd_epilogue;

/* Generic tcgen05 implicit-GEMM:  acc[r, n] = sum_{tap, c} A[(x,y,z)(r) * in_mul + tap_off(tap), c] * W[n, tap*Cin + c]
 * A is a 4-D channels-last act tensor (C innermost, then X, Y, Z) with element strides; out-of-range
 * coordinates read as zero (= conv zero padding). Output rows are (z*Yo + y)*Xo + x.
 * Replaces nn.Linear / nn.Conv2d(3x3, 1x1, stride 1|2) / nn.Conv3d((3,1,1)) / torch.bmm on the hot path. */
typedef struct gcd_tc_op {
    const void* A;
    int32_t C;                 /* channels per tap (multiple of 8) */
    int32_t Xi, Yi, Zi;        /* input extents */
    int64_t sx, sy, sz;        /* input strides in elements (channel stride is 1) */
    int32_t Xo, Yo, Zo;        /* output extents */
    int32_t in_mul;            /* 1, or 2 for stride-2 conv */
    int32_t ntaps;             /* 1..9 */
    int8_t tap_dx[9], tap_dy[9], tap_dz[9];
    int32_t gemm_tile;         /* 1: plain GEMM tiling (128 consecutive x), 0: search conv tile */
    const void* W;             /* [N, ldw] act, K-contiguous, ldw >= ntaps*C ; batched over y if w_batch_stride != 0 */
    int64_t ldw;
    int64_t w_batch_stride;    /* elements between per-y weight matrices (batched GEMM), else 0 */
    int32_t N;                 /* accumulator columns */
    gcd_epilogue ep;
} gcd_tc_op;
/* returns 0 on success, 1 on success without the requested fused GroupNorm statistics, < 0 on error */
int gcd_tc_run(const gcd_tc_op* op, void* stream);

/* ---- normalisation (HBM-bound kernels) ------------------------------------------------------------------ */
/* GroupNorm statistics over `rows_per_group_img` rows x (C/groups) channels: util.py:274-276 GroupNorm32,
 * attention.py:125-128 Normalize. in: [n_img*rows, C] float32 (in_f32=1) or act. stats: double[n_img*groups*2]
 * (sum, sumsq), must be zeroed by the caller (gcd_memset_async) before the call. */
int gcd_groupnorm_stats(const void* in, int in_f32, int64_t n_img, int64_t rows, int C, int groups, double* stats,
                        void* stream);
/* y = (x-mean)*rstd*gamma+beta, optional SiLU, written as act. */
int gcd_groupnorm_apply(const void* in, int in_f32, int64_t n_img, int64_t rows, int C, int groups,
                        const double* stats, const float* gamma, const float* beta, float eps, int silu, void* out,
                        void* stream);
/* LayerNorm over C per row (attention.py BasicTransformerBlock norm1-3, video_attention.py norm_in/1/2/3).
 * in: float32 [rows, C]. Optional `add` [n_add, C] float32 is added first with index (row / add_rows_per) % add_mod
 * (video_attention.py:266-287 time_pos_embed) and the sum written to sum_out (float32) when non-NULL. */
int gcd_layernorm(const float* in, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                  const float* add, int64_t add_rows_per, int64_t add_mod, float* sum_out, void* out, void* stream);

/* ---- attention ------------------------------------------------------------------------------------------- */
/* Spatial self-attention softmax(QK^T/sqrt(64))V, head_dim 64 (attention.py:255-344 CrossAttention with
 * context=None). qkv: act [frames, tokens, 3*heads*64] = (q | k | v), head index outer within each third.
 * out: act [frames, tokens, heads*64]. */
int gcd_attention_spatial(const void* qkv, int frames, int tokens, int heads, void* out, void* stream);
/* Temporal self-attention over T frames per (clip, spatial position) (video_attention.py:109-140 attn1 after the
 * "(b t) s c -> (b s) t c" rearrange, done here by indexing). qkv/out laid out as above with frames = clips*T. */
int gcd_attention_temporal(const void* qkv, int clips, int T, int tokens, int heads, void* out, void* stream);
/* Row softmax of float32 scores [rows, cols] * scale -> act probs (model.py:161-201 AttnBlock, d=512 single head). */
int gcd_softmax_rows(const float* in, int64_t rows, int cols, float scale, void* out, void* stream);

/* ---- element-wise / layout ------------------------------------------------------------------------------- */
int gcd_memset_async(void* p, int value, int64_t bytes, void* stream);
/* float32 [rows, C] -> act [rows, C] */
int gcd_cast_f32_to_act(const float* in, int64_t n, void* out, void* stream);
/* nearest x2 upsample of channels-last float32 [n,H,W,C] -> act [n,2H,2W,C] (openaimodel.py:110-160, model.py:58-71) */
int gcd_upsample2x_to_act(const float* in, int n, int H, int W, int C, void* out, void* stream);
/* channel concat of two float32 channels-last tensors: out[r, :Ca]=a, out[r, Ca:]=b  (video_model.py:525 th.cat) */
int gcd_concat_channels(const float* a, int Ca, const float* b, int Cb, int64_t rows, float* out, void* stream);
/* the same concat of [n_img*rows, Ca] and [n_img*rows, Cb] fused with the GroupNorm statistics of the result (the skip concat
 * of the UNet's output blocks feeds a ResBlock whose first op is GroupNorm32, openaimodel.py / video_model.py:536-540):
 * stats[n_img, groups, 2] float64 (sum, sum of squares), zeroed by the caller, accumulated in a fixed order per block. */
int gcd_concat_channels_stats(const float* a, int Ca, const float* b, int Cb, int64_t n_img, int64_t rows, int groups,
                              float* out, double* stats, void* stream);
/* Same, the concatenated tensor written in the 16-bit activation type only (statistics of the fp32 inputs): for consumers that read
 * 16-bit operands anyway — the UNet's output ResBlocks (cin != cout: 1x1 skip conv + first GroupNorm), openaimodel.py `th.cat`. */
int gcd_concat_channels_stats_act(const float* a, int Ca, const float* b, int Cb, int64_t n_img, int64_t rows, int groups,
                                  void* out, double* stats, void* stream);
/* act(x) on act tensor: SiLU (emb_layers SiLU, openaimodel.py:262-268) */
int gcd_silu_act(const void* in, int64_t n, void* out, void* stream);
/* SiLU on a float32 tensor, written as act (emb_layers' nn.SiLU on `emb`, openaimodel.py:262-268) */
int gcd_silu_f32_to_act(const float* in, int64_t n, void* out, void* stream);
/* NCHW float32 [N,C,HW] -> channels-last act [N,HW,Cpad] (zero-padded channels): entry glue of VideoUNet.forward /
 * VideoDecoder.forward when called through the plugin surface with torch NCHW tensors. */
int gcd_nchw_to_act_nhwc(const float* in, int N, int C, int HW, int Cpad, void* out, void* stream);
/* channels-last float32 [N,HW,ld] (first C columns) -> NCHW float32 [N,C,HW]: exit glue. */
int gcd_nhwc_to_nchw_f32(const float* in, int ld, int N, int C, int HW, float* out, void* stream);
/* AE3DConv.time_mix_conv (temporal_ae.py:86-107): Conv3d(3->3,(3,1,1)) over frames of the VAE's 3-channel output.
 * in: channels-last float32 [B*T, HW, ld] (first 3 columns); w: [3,3,3] (co,ci,kt); out: NCHW float32 [B*T,3,HW]. */
int gcd_vae_time_mix(const float* in, int ld, int B, int T, int HW, const float* w, const float* b, float* out,
                     void* stream);
/* timestep_embedding (util.py:207-231): t[n] float32 -> act [n, dim] = cat(cos, sin)(t * exp(-ln(max_period) k / half)) */
int gcd_timestep_embedding(const float* t, int n, int dim, float max_period, void* out_act, float* out_f32,
                           void* stream);

/* SphericalEmbedder.forward (encoders/modules.py:247-287): x[n,3] = (azimuth, elevation, radius) float32
 * -> out[n,dim] float32 = [cos/sin of 1x,2x,4x azimuth | same for elevation | radius] * w[dim,13]^T + b[dim]. */
int gcd_spherical_embed(const float* x, int n, const float* w, const float* b, int dim, float* out, void* stream);

/* ---- sampler step (sampling.py:101-121, denoiser.py:23-49, guiders.py:79-100, wrappers.py:23-34) ------------ */
/* The scalar coefficients c_in/c_out/c_skip (denoiser_scaling.py:53-61) and dt = sigma_next - sigma_hat are computed
 * by the host with the same torch fp32 CPU ops as the reference and passed in, so scheduler arithmetic is bit-exact.
 * Builds the CFG-doubled network input: out act [2*BT, H, W, 64] channels-last, channels 0..3 = x * c_in,
 * 4..7 = concat cond (zeros for the unconditional half), 8..63 = 0. x: float32 NCHW [BT,4,H,W];
 * cond_concat: float32 NCHW [BT,4,H,W]; uc_concat likewise (may be NULL = zeros). */
int gcd_sampler_prep(const float* x, const float* uc_concat, const float* c_concat, int BT, int H, int W, float c_in,
                     void* out, void* stream);
/* net_out: float32 channels-last [2*BT, H, W, ld_net] (first 4 columns used), x updated IN PLACE (NCHW float32):
 *   den_k = net_k * c_out + x * c_skip ; den = den_u + scale[t](den_c - den_u) ; d = (x - den)/sigma ;
 *   x += d * dt.  scale: float32 [T] (frame index = bt % T). */
int gcd_sampler_update(float* x, const float* net_out, int ld_net, int BT, int T, int H, int W, float c_out,
                       float c_skip, float sigma, float dt, const float* scale, void* stream);

/* Experiments only (tools/autotune_tc.py): force tile width (128 / 160 / 256) and cluster mode (2 = weight multicast, 3 = CTA-pair
 * MMA) of the following gcd_tc_run calls; 0 = automatic. Not thread-safe. */
void gcd_tc_override(int bn, int mode);

/* Evaluation-loop image metrics (gcd-model/scripts/test.py:346-496 calculate_metrics; SSIM = scikit-image 0.22.0
 * structural_similarity(data_range=1, channel_axis=0), masked form = gcd-model/scripts/eval_utils.py:571-664 masked_ssim).
 * pred, gt: float32 [frames, 3, H, W] in [0, 1]; mask: uint8 [frames, H, W] or NULL. out: float64 [frames, 8], zeroed here:
 *   {sum (pred-gt)^2, count, same over mask, count, sum SSIM map (3-pixel border cropped), count, same over the 3x eroded mask, count},
 * sums over the 3 channels. PSNR = 10 log10(count / sum) for data_range 1; SSIM = sum / count. */
int gcd_frame_metrics(const float* pred, const float* gt, const uint8_t* mask, int frames, int H, int W, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
