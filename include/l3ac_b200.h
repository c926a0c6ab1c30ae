/*
 * l3ac_b200 -- C ABI of the B200 (sm_100a) kernels behind the L3AC encode/quantize/decode hot path.
 *
 * The reference (zhai-lw/L3AC) has no FFI: its hot path is the body of two Python methods,
 *   L3AC.encode_audio   l3ac/__init__.py:108-114
 *   L3AC.decode_audio   l3ac/__init__.py:116-121
 * which call stock ATen operators.  This header is the operator-level boundary that replaces
 * those ATen call sites; each entry point cites the reference operator(s) it stands in for
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding and how the
 * reference's `L3AC` class would call it.
 *
 * Conventions
 *  - Plain C: raw device pointers, sizes, a `cudaStream_t` passed as `void*`.  No torch types.
 *  - The library allocates nothing and keeps no global state: outputs and workspaces are
 *    caller-allocated device buffers; all work is enqueued on the given stream (graph-capturable).
 *  - Activations are time-major / channels-last: a (B, T, C) tensor is B*T rows of C contiguous
 *    values.  (The reference is channels-first (B, C, T) in the conv stack; the host layer
 *    converts at the two ends of the path only.)
 *  - Return value: 0 = ok; <0 = argument validation error (L3AC_E*); >0 = cudaError_t.
 */
#ifndef L3AC_B200_H_
#define L3AC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L3AC_OK 0
#define L3AC_EINVAL (-1)        /* bad argument (NULL pointer, size out of range)   */
#define L3AC_EUNSUPPORTED (-2)  /* valid request this build has no kernel for       */
#define L3AC_EDRIVER (-3)       /* CUDA driver entry point unavailable (TMA encode) */

/* activation element types for `out_dtype` */
#define L3AC_F32 0
#define L3AC_BF16 1
#define L3AC_BF16X2 2 /* split pair: out = bf16(x) plane, out_lo = bf16(x - hi) plane (same shape and pitch) */

/* epilogue activations of the GEMM entry points */
#define L3AC_ACT_NONE 0
#define L3AC_ACT_SNAKE 1 /* snake(x; alpha) then optional per-column affine (the folded GRN)      */
#define L3AC_ACT_GEGLU 2 /* columns come in (value, gate) pairs: out[n/2] = v * gelu_erf(g)       */
#define L3AC_ACT_GELU 3
#define L3AC_ACT_TANH 4

typedef void* l3ac_stream_t; /* cudaStream_t */

int l3ac_abi_version(void);
const char* l3ac_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Encoder stem.  Replaces V3FirstBlock.forward (l3ac/tconv/__init__.py:16-22) incl. BaseBlock
 * (l3ac/tconv/base.py:27-45) and trend_pool (l3ac/tconv/base.py:8-14):
 *   5 x [TrendPool(k) -> Conv1d(1->4,k7)] -> cat(20) -> 1x1 20->80 -> GELU -> cat(x) -> 1x1 81->C.
 * audio (B,T) fp32 -> out (B,T,C) fp32.  pool kernels are fixed (1,5,11,21,45).  C must be 24.
 *   branch_w [5][4][7], branch_b [20], w1 [80][20], b1 [80], w2 [C][81], b2 [C]  (all folded fp32)
 * ------------------------------------------------------------------------------------------ */
int l3ac_stem(const float* audio, int B, int T, const float* branch_w, const float* branch_b,
              const float* w1, const float* b1, const float* w2, const float* b2, int C, float* out,
              l3ac_stream_t stream);

/* The same stem on the tensor cores at fp32-class precision (the two 1x1 convs as 3-term split-bf16 MMAs, fp32
 * accumulate; the pooled signals by a 4-neighbour two-level scheme, i.e. a different but equally accurate fp32
 * summation order).  Same arguments; out 8-byte aligned.  Used by the default (split-bf16) encode path. */
int l3ac_stem_tc(const float* audio, int B, int T, const float* branch_w, const float* branch_b,
              const float* w1, const float* b1, const float* w2, const float* b2, int C, float* out,
              l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ConvUnit prologue.  Replaces dw_conv + permute + norm of ConvUnit.forward
 * (l3ac/modules.py:33-35; F.layer_norm at l3ac/layers.py:80): depthwise Conv1d(C,C,k7,pad 3) then
 * LayerNorm over C.  x (B,T,C) fp32 -> out (B,T,C) fp32|bf16|bf16 (hi,lo) pair.  dw_w is [7][C]
 * (tap-major).  out_lo is the low-order plane for L3AC_BF16X2, NULL otherwise.
 * ------------------------------------------------------------------------------------------ */
int l3ac_dwconv7_ln(const float* x, int B, int T, int C, const float* dw_w, const float* dw_b,
                    const float* ln_w, const float* ln_b, float eps, void* out, void* out_lo, int out_dtype,
                    l3ac_stream_t stream);

/* The same ConvUnit prologue through a plan, for the decode side's thin stages (C = 48 / 96, bf16 out): the per-channel
 * parameters (HOST arrays at plan creation: dw_w [7][C], dw_b / ln_w / ln_b [C]) stay in the plan and travel as kernel
 * parameters, so a thread-per-row kernel (128-row tiles staged in shared memory, taps as FFMAs against the constant bank,
 * thread-local LayerNorm statistics) replaces the lane-group kernel.  x (B,T,C) fp32 -> out (B,T,C) bf16, both 16-byte
 * aligned; B <= 65535.  Same arithmetic as l3ac_dwconv7_ln up to the summation order of the LayerNorm statistics. */
typedef struct l3ac_dwconv_plan l3ac_dwconv_plan;
int l3ac_dwconv_plan_create(int C, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b, float eps,
                            l3ac_dwconv_plan** plan_out);
int l3ac_dwconv_plan_destroy(l3ac_dwconv_plan* plan);
int l3ac_dwconv7_ln_plan(const l3ac_dwconv_plan* plan, const float* x, int B, int T, void* out, l3ac_stream_t stream);

/* LayerNorm over the last dim.  Replaces channels-first ChannelNorm (l3ac/layers.py:50-56) after the
 * strided convs and nn.LayerNorm inside local_attention's LocalMHA / FeedForward. */
int l3ac_layernorm(const float* x, long long M, int C, const float* w, const float* b, float eps,
                   void* out, void* out_lo, int out_dtype, l3ac_stream_t stream);

/* fp32 -> (hi, lo) bf16 planes with hi = bf16(x), lo = bf16(x - hi): the operand format of the split GEMM.
 * n must be a multiple of 4. */
int l3ac_split_bf16(const float* x, long long n, void* hi, void* lo, l3ac_stream_t stream);

/* Elementwise snake (l3ac/layers.py:29-33), per-channel alpha [C]. */
int l3ac_snake(const float* x, long long M, int C, const float* alpha, void* out, int out_dtype,
               l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GEMM / conv-as-GEMM with fused epilogue.  Replaces nn.Linear / nn.Conv1d call sites:
 *   pw_conv1 / pw_conv2 (l3ac/modules.py:36,39), strided down convs (modules.py:97; local_trans.py:136),
 *   k3 edge convs (modules.py:110,150), 1x1 up convs (modules.py:161), LegacyUnit convs
 *   (modules.py:55,57), to_qkv / to_out / FeedForward linears (local_attention, call site
 *   local_trans.py:34-39), with snake + GRN (layers.py:29-33,112-115), GEGLU and Residual
 *   (xtract/nn/layers.py:59-62) fused.
 *
 *   out[m, n] = epi( bias[n] + sum_{s<taps} sum_{k<K} A[b, t + tap_shift0 + s*tap_step, k] * W[n, s*K + k] )
 *   with m = b*T + t; rows outside [0,T) read as zero (the conv's zero padding).
 *   epi: act (see L3AC_ACT_*), then `+ residual[m, n]` if residual != NULL.
 *   For L3AC_ACT_SNAKE: v = snake(v, alpha[n]); if scale: v = v*scale[n] + shift[n].
 *   For L3AC_ACT_GEGLU: N counts the interleaved (value,gate) columns; out has N/2 columns.
 * ------------------------------------------------------------------------------------------ */
typedef struct l3ac_gemm_desc {
    const void* A;         /* [B*T, lda]: fp32 (l3ac_gemm_f32) or bf16 (l3ac_gemm_bf16_tc)          */
    const void* W;         /* [N, taps*K]  same element type as A                                   */
    const float* bias;     /* [N] or NULL                                                           */
    const float* alpha;    /* [N] (ACT_SNAKE)                                                       */
    const float* scale;    /* [N] or NULL (ACT_SNAKE)                                               */
    const float* shift;    /* [N] or NULL (ACT_SNAKE)                                               */
    const float* residual; /* [B*T, ldr] fp32 or NULL                                               */
    void* out;             /* [B*T, ldo] fp32 or bf16                                               */
    const void* A_lo;      /* tcgen05 path only: low-order bf16 plane of A (same layout) or NULL       */
    const void* W_lo;      /* tcgen05 path only: low-order bf16 plane of W; given iff A_lo is given    */
    void* out_lo;          /* low-order output plane when out_dtype == L3AC_BF16X2                    */
    long long lda, ldr, ldo; /* row pitches in elements                                            */
    int B, T, K, N;
    int taps, tap_shift0, tap_step;
    int act, out_dtype;
} l3ac_gemm_desc;

/* fp32 SIMT reference-precision path (used for the encoder side and for the fp32 mode). */
int l3ac_gemm_f32(const l3ac_gemm_desc* d, l3ac_stream_t stream);

/* bf16 x bf16 -> fp32 tcgen05/TMEM path fed by TMA.  A and W are bf16; lda and taps*K must be
 * multiples of 8 (16-byte TMA pitch).  With A_lo/W_lo the product is the 3-term split
 * A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (fp32-class accuracy at bf16 tensor-core rate). */
int l3ac_gemm_bf16_tc(const l3ac_gemm_desc* d, l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused ConvUnit MLP (pw_conv1 -> Snake -> GRN -> pw_conv2 -> Residual; l3ac/modules.py:36-44, layers.py:29-33,112-115,
 * xtract/nn/layers.py:59-62) in one tcgen05 kernel: the 4C-wide hidden activation stays in TMEM / shared memory.
 *   out[m,:] = residual[m,:] + b2 + W2 . ( scale * snake(W1 . a[m,:] + b1; alpha) + shift )
 * a (M,C) bf16, w1 (4C,C) bf16, w2 (C,4C) bf16, b1/alpha/ialpha/scale/shift [4C] with ialpha = 1/(alpha + 1e-8),
 * b2 [C], residual/out (M,C) fp32.
 * 16 <= C <= 256, C % 8 == 0 (C = 512 does not fit one SM's shared memory: use the two GEMM calls).
 * ------------------------------------------------------------------------------------------ */
int l3ac_convunit_mlp_tc(const void* a, const void* w1, const float* b1, const float* alpha, const float* ialpha,
                         const float* scale, const float* shift, const void* w2, const float* b2, const float* residual,
                         float* out, long long M, int C, l3ac_stream_t stream);

/* The same kernel with one more output: ch0_out (M) fp32 = channel 0 of `out` as a compact plane (NULL: none).  EnhanceBlock
 * (l3ac/tconv/__init__.py:40-44) derives its four branch signals from channel 0 only; reading it from the (M, C) tensor costs
 * a 128-byte line per row, so the unit in front of an EnhanceBlock emits it and l3ac_enhance_stats takes it as x with C = 1. */
int l3ac_convunit_mlp_tc_ch0(const void* a, const void* w1, const float* b1, const float* alpha, const float* ialpha,
                             const float* scale, const float* shift, const void* w2, const float* b2, const float* residual,
                             float* out, float* ch0_out, long long M, int C, l3ac_stream_t stream);

/* Whole Residual(ConvUnit) (l3ac/modules.py:10-44) for the thin full-rate encoder stage (C = 24, hidden 96) as one fp32
 * kernel: depthwise conv k7 + LayerNorm + pw_conv1 + Snake + GRN affine + pw_conv2 + residual; the hidden activation
 * never leaves registers.  x (B,T,24) fp32; dw_w [7][24]; w1 [96][24]; w2 [24][96]; b1/alpha/scale/shift [96].
 * out_dtype L3AC_F32: out (B,T,24) fp32, out_lo ignored.  L3AC_BF16X2: out / out_lo (B,T,24) bf16 = the split pair
 * (hi = bf16(v), lo = bf16(v - hi)) that the strided down-conv GEMM of the stage consumes. */
int l3ac_convunit_thin_f32(const float* x, int B, int T, int C, const float* dw_w, const float* dw_b, const float* ln_w,
                           const float* ln_b, float eps, const float* w1, const float* b1, const float* alpha,
                           const float* scale, const float* shift, const float* w2, const float* b2, void* out,
                           void* out_lo, int out_dtype, l3ac_stream_t stream);

/* The same fused Residual(ConvUnit) (l3ac/modules.py:10-44) on the tensor cores at fp32-class precision, for the
 * encode-side stages with C = 24 and C = 48: dwconv7 + LayerNorm in registers, pw_conv1 / pw_conv2 as 3-term
 * split-bf16 MMAs (hi*Whi + lo*Whi + hi*Wlo, fp32 accumulate) with the hidden activation kept in registers.
 * Arguments as l3ac_convunit_thin_f32 (fp32 weights: w1 [4C][C], w2 [C][4C]; they are split on the fly);
 * x / out / out_lo 16-byte aligned.  operand_dtype L3AC_BF16X2: 3-term split products (encode side);
 * L3AC_BF16: plain bf16 operands, fp32 accumulate (the decode side's arithmetic; used for its C = 48 unit). */
int l3ac_convunit_thin_tc(const float* x, int B, int T, int C, const float* dw_w, const float* dw_b, const float* ln_w,
                          const float* ln_b, float eps, const float* w1, const float* b1, const float* alpha,
                          const float* scale, const float* shift, const float* w2, const float* b2, void* out,
                          void* out_lo, int out_dtype, int operand_dtype, l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Block-local causal attention.  Replaces LocalAttention.forward of local-attention==1.11.2 as
 * configured at l3ac/local_trans.py:34-38 (causal, look_backward=1, exact_windowsize=False, autopad)
 * plus the DynamicPositionBias gather (local_trans.py:43):
 *   query p attends keys [ (floor(p/w)-1)*w clipped at 0 , p ];  logit = q.k/sqrt(D) + bias[h][p-k].
 * qkv (B,T,3*H*D) fp32 packed [q | k | v], each (H,D)-major -> out (B,T,H*D).  bias_table [H][2w].
 * D must be 32.
 * ------------------------------------------------------------------------------------------ */
int l3ac_local_attention_f32(const float* qkv, const float* bias_table, int B, int T, int H, int D,
                             int window, float* out, l3ac_stream_t stream);

/* Tensor-core variant of the same attention over bf16 q/k/v (the QKV GEMM writes them directly).  qkv_hi (and
 * qkv_lo for the 3-term split product, else NULL) are (B,T,3*H*D) bf16 planes packed [q | k | v]; out (B,T,H*D) is
 * fp32, bf16 or a bf16 (hi, lo) pair (out_lo) -- the A operand of the output projection.  D must be 32. */
int l3ac_local_attention_tc(const void* qkv_hi, const void* qkv_lo, const float* bias_table, int B, int T, int H,
                            int D, int window, void* out, void* out_lo, int out_dtype, l3ac_stream_t stream);

/* The same attention on the 5th-generation tensor cores (tcgen05.mma, S and O accumulators in TMEM; one thread per query
 * row does the softmax on its TMEM lane; V is consumed in its natural layout as an MN-major operand and the row sums come
 * out of the P.V product through a ones column).  Same arguments and results as l3ac_local_attention_tc; additionally
 * out / out_lo must be 16-byte aligned.  This is the product path; the mma.sync kernel above is kept as a cross-check. */
int l3ac_local_attention_umma(const void* qkv_hi, const void* qkv_lo, const float* bias_table, int B, int T, int H,
                              int D, int window, void* out, void* out_lo, int out_dtype, l3ac_stream_t stream);

/* Rotary-position path (en_coder_dynamic_pos = false: LocalMHA(use_rotary_pos_emb=True), l3ac/local_trans.py:29,36;
 * replaces SinusoidalEmbeddings + apply_rotary_pos_emb of local-attention inside LocalAttention.forward).
 * l3ac_rotary_pack: qkv (B,T,3*H*D) fp32 [q | k | v] -> one segment per attention window, (B*ceil(T/window), 2*window,
 * 3*H*D) in fp32 / bf16 / bf16 (hi, lo) pair, keys rotated by their bucket position 0..2w-1 and queries by w..2w-1
 * (cos_table / sin_table: (2*window, D) fp32, host-computed like the reference).  The attention entry points above then
 * run on the segments with T = 2*window and a zero bias table; l3ac_rotary_unpack gathers the (B,T,row_bytes) result rows
 * back from the (B*ceil(T/window), 2*window, row_bytes) segment output.  D must be 32. */
int l3ac_rotary_pack(const float* qkv, int B, int T, int H, int D, int window, const float* cos_table,
                     const float* sin_table, void* out, void* out_lo, int out_dtype, l3ac_stream_t stream);
int l3ac_rotary_unpack(const void* seg, void* out, int B, int T, int window, int row_bytes, l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FSQ bottleneck.  Replaces VQEmbed.forward (l3ac/vq/__init__.py:25-30) = project_in ->
 * SuperFSQ.forward (l3ac/vq/fsq.py:30-68; tanh_act l3ac/vq/fsq_act.py:38-39) -> project_out, and
 * VQEmbed.to_features (l3ac/vq/__init__.py:20-23; fsq.py:70-81).
 *   x (M,F) fp32;  w_in [D][F], b_in [D], w_out [F][D], b_out [F];  levels: HOST array of D ints, D<=8.
 *   q_feature (M,F) fp32, indices (M) int32, level_indices (M,D) fp32 (may be NULL), z (M,D) (may be NULL).
 * l3ac_fsq_quantize_latents takes the D-dim latents directly (bit-exactness check of the quantiser).
 * ------------------------------------------------------------------------------------------ */
int l3ac_fsq_quantize(const float* x, long long M, int F, const float* w_in, const float* b_in,
                      const float* w_out, const float* b_out, const int* levels, int D, float* q_feature,
                      int32_t* indices, float* level_indices, float* z, l3ac_stream_t stream);
int l3ac_fsq_quantize_latents(const float* z, long long M, const int* levels, int D, float* q_z,
                              int32_t* indices, float* level_indices, l3ac_stream_t stream);
int l3ac_fsq_dequantize(const void* indices, int indices_are_i64, long long M, int F, const float* w_out,
                        const float* b_out, const int* levels, int D, float* q_feature,
                        l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Linear x`scale` upsampling along time (nn.Upsample(mode='linear', align_corners=False),
 * l3ac/modules.py:162, l3ac/local_trans.py:121) with the following channels-first ChannelNorm
 * (l3ac/modules.py:163, l3ac/layers.py:50-56) fused when cn_w != NULL.
 * x (B,T,C) fp32 -> out (B,T*scale,C) fp32.
 * ------------------------------------------------------------------------------------------ */
int l3ac_upsample_linear_cn(const float* x, int B, int T, int C, int scale, const float* cn_w,
                            const float* cn_b, float eps, float* out, l3ac_stream_t stream);

/* The up layer's tail fused with the next ConvUnit's prologue (decode side, bf16 operands; C = 48 / 96, scale 2 / 3):
 *   x_up = ChannelNorm(Upsample_linear(y, scale))   (l3ac/modules.py:162-163)   -> (B, T*scale, C) fp32, the next stage's residual stream
 *   a    = LayerNorm(dwconv7(x_up))                 (l3ac/modules.py:33-35)     -> (B, T*scale, C) bf16, the operand of l3ac_convunit_mlp_tc
 * in one kernel: x_up is written once and never read back for the depthwise conv.  The per-channel parameters (HOST arrays
 * at plan creation: cn_w / cn_b [C], dw_w [7][C], dw_b / ln_w / ln_b [C]) stay in the plan and travel as kernel parameters.
 * Same arithmetic as l3ac_upsample_linear_cn followed by l3ac_dwconv7_ln_plan.  y, x_up, a_out 16-byte aligned; B <= 65535. */
typedef struct l3ac_updw_plan l3ac_updw_plan;
int l3ac_updw_plan_create(int C, int scale, const float* cn_w, const float* cn_b, float cn_eps, const float* dw_w, const float* dw_b,
                          const float* ln_w, const float* ln_b, float ln_eps, l3ac_updw_plan** plan_out);
int l3ac_updw_plan_destroy(l3ac_updw_plan* plan);
int l3ac_upsample_cn_dwconv7_ln(const l3ac_updw_plan* plan, const float* y, int B, int T, float* x_up, void* a_out,
                                l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * EnhanceBlock (l3ac/tconv/__init__.py:30-44): 4 x [TrendPool(k in 1,3,5,9) -> Conv1d(1->1,k7,dil 1,2,3,5)]
 * on channel 0 -> InstanceNorm1d(4, affine, eps 1e-5, stats over all T) -> Conv1d(4->C,1x1) -> x + y*x.
 * Two passes: `stats` writes per-(b,chunk) partial sums, `apply` reduces them and gates.
 * (`stats` reads channel 0 of x only: pass a compact channel-0 plane as x with C = 1 when one is available.)
 * partials: l3ac_enhance_partials_floats(B,T) floats.  conv_w [4][7], conv_b [4], in_w/in_b [4],
 * merge_w [C][4], merge_b [C].  out (B,T,C) fp32|bf16.
 * ------------------------------------------------------------------------------------------ */
long long l3ac_enhance_partials_floats(int B, int T);
/* branches (optional, 16-byte aligned): (B,T,4) fp32 -- when given, `stats` also stores the four un-normalised branch
 * signals of every sample and `apply` streams them back instead of recomputing the pooling / convolutions per tile. */
int l3ac_enhance_stats(const float* x, int B, int T, int C, const float* conv_w, const float* conv_b,
                       float* partials, float* branches, l3ac_stream_t stream);
int l3ac_enhance_apply(const float* x, int B, int T, int C, const float* conv_w, const float* conv_b,
                       const float* in_w, const float* in_b, const float* merge_w, const float* merge_b,
                       const float* partials, const float* branches, void* out, int out_dtype, l3ac_stream_t stream);

/* EnhanceBlock gate + the up layer's 1x1 conv in one kernel for the thin decode stages ((C_in, C_out) = (48, 24) / (96, 48), bf16
 * operands): out = bf16(x + y * x) . W^T + b with y from the stats pass' partials / branches (l3ac/tconv/__init__.py:40-44 ->
 * l3ac/modules.py:161).  The gated activation goes straight into mma.sync fragments; it is never written to memory.  Plan from HOST
 * arrays: in_w / in_b [4] (InstanceNorm affine), merge_w [C_in][4], merge_b [C_in], up_w [C_out][C_in] (folded), up_b [C_out].
 * x (B,T,C_in) fp32, partials / branches as written by l3ac_enhance_stats, out (B,T,C_out) fp32; 16-byte aligned; B <= 65535. */
typedef struct l3ac_enhup_plan l3ac_enhup_plan;
int l3ac_enhup_plan_create(int C_in, int C_out, const float* in_w, const float* in_b, const float* merge_w, const float* merge_b,
                           const float* up_w, const float* up_b, l3ac_enhup_plan** plan_out);
int l3ac_enhup_plan_destroy(l3ac_enhup_plan* plan);
int l3ac_enhance_up(const l3ac_enhup_plan* plan, const float* x, int B, int T, const float* partials, const float* branches,
                    float* out, l3ac_stream_t stream);

/* Decoder tail (l3ac/modules.py:192-194): Snake(C) -> Conv1d(C->1,k7,pad 3) -> tanh.
 * x (B,T,C) fp32 -> out (B,T) fp32.  w is [7][C] (tap-major). */
int l3ac_tail_conv_tanh(const float* x, int B, int T, int C, const float* alpha, const float* w, float bias,
                        float* out, l3ac_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused full-rate decoder tail: 3 x Residual(LegacyUnit) (l3ac/modules.py:47-64,174-179; dilations 1,3,9)
 * + Snake -> Conv1d(C->1,k7,pad 3) -> tanh (l3ac/modules.py:192-194) in one kernel.
 * x (B,T,24) fp32 -> out (B,T) fp32.  The k7 and 1x1 conv weights are bf16, pre-packed in mma.m16n8k16
 * B-fragment order (both pointers 16-byte aligned):
 *   conv_frags [3 units][11 ksteps][3 ntiles][32 lanes][4] bf16 over the tap-major K axis k = 24*tap + channel
 *   (168 -> 176, zero padded); pw_frags [3][2][3][32][4] bf16 with K padded 24 -> 32.  The 4 values of lane l are
 *   W[n][k0], W[n][k0+1], W[n][k0+8], W[n][k0+9], n = 8*ntile + l/4, k0 = 16*kstep + 2*(l%4).
 *   conv_bias/pw_bias/alpha0/alpha1 are [3][24]; dilations is a HOST array of 3 ints, each <= 9, whose receptive
 *   field 3*(d0+d1+d2)+3 must be <= 42.  alpha_f [24], w_f [7][24] (tap-major), bias_f: final conv.
 * ------------------------------------------------------------------------------------------ */
int l3ac_decoder_tail(const float* x, int B, int T, int C, const void* conv_frags, const float* conv_bias,
                      const void* pw_frags, const float* pw_bias, const float* alpha0, const float* alpha1,
                      const int* dilations, const float* alpha_f, const float* w_f, float bias_f, float* out,
                      l3ac_stream_t stream);

/* The same fused decoder tail on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM; the bf16 snake
 * outputs are written to shared memory by their row-owner threads and the conv taps are row-shifted operand
 * descriptors of that one tile).  The immutable packed weights live in a plan (handle):
 *   l3ac_tail_plan_create  takes HOST arrays in the reference's layout, folded fp32:
 *       conv_w [3][24 out][24 in][7 taps]  (LegacyUnit.block.1, l3ac/modules.py:54), conv_b [3][24],
 *       pw_w [3][24 out][24 in] (block.3), pw_b [3][24], alpha0 / alpha1 [3][24] (block.0 / block.2 Snake),
 *       dilations [3], alpha_f [24], w_f [7 taps][24], bias_f (Decoder tail, l3ac/modules.py:192-194);
 *     converts them to bf16 operand order and uploads them to the CURRENT device (one cudaMalloc + cudaMemcpy).
 *   l3ac_decoder_tail_tc   x (B,T,24) fp32 -> out (B,T) fp32 on `stream`; allocates nothing; graph-capturable.
 *   l3ac_tail_plan_destroy frees the device blob.
 * Same constraints on the dilations as l3ac_decoder_tail. */
/* The fused Residual(ConvUnit) of the thin encode-side stages (C = 24 / 48; see l3ac_convunit_thin_tc) on the 5th-generation
 * tensor cores: pw_conv1 / pw_conv2 as 3-term split-bf16 tcgen05.mma with TMEM accumulators (fp32-class), one thread per time
 * step does dwconv7 + LayerNorm, snake + GRN affine and the operand splitting.  Weights go into a plan built once from HOST
 * arrays (layouts of l3ac_convunit_thin_f32: dw_w [7][C], w1 [4C][C], w2 [C][4C], b1/alpha/scale/shift [4C]).
 * l3ac_convunit_umma: x (B,T,C) fp32 -> out fp32 (out_dtype L3AC_F32) or the split pair out / out_lo (L3AC_BF16X2); all
 * pointers 16-byte aligned; allocates nothing. */
typedef struct l3ac_convunit_plan l3ac_convunit_plan;
int l3ac_convunit_plan_create(int C, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b, float eps,
                              const float* w1, const float* b1, const float* alpha, const float* scale, const float* shift,
                              const float* w2, const float* b2, l3ac_convunit_plan** plan_out);
int l3ac_convunit_plan_destroy(l3ac_convunit_plan* plan);
int l3ac_convunit_umma(const l3ac_convunit_plan* plan, const float* x, int B, int T, void* out, void* out_lo, int out_dtype,
                       l3ac_stream_t stream);

/* The encoder stem (V3FirstBlock, see l3ac_stem above) on the 5th-generation tensor cores: the two 1x1 convs are 3-term
 * split-bf16 tcgen05.mma with TMEM accumulators (fp32-class results), one thread per sample does the pooling taps, GELU and
 * operand splitting.  Weights go into a plan built once from HOST arrays in the reference's layout (folded fp32):
 *   branch_w [5][4][7], branch_b [20], w1 [80][20], b1 [80], w2 [24][81], b2 [24].
 * l3ac_stem_umma: audio (B,T) fp32 -> out (B,T,24) fp32 (16-byte aligned) on `stream`; allocates nothing. */
typedef struct l3ac_stem_plan l3ac_stem_plan;
int l3ac_stem_plan_create(const float* branch_w, const float* branch_b, const float* w1, const float* b1, const float* w2,
                          const float* b2, int C, l3ac_stem_plan** plan_out);
int l3ac_stem_plan_destroy(l3ac_stem_plan* plan);
int l3ac_stem_umma(const l3ac_stem_plan* plan, const float* audio, int B, int T, float* out, l3ac_stream_t stream);

typedef struct l3ac_tail_plan l3ac_tail_plan;
int l3ac_tail_plan_create(const float* conv_w, const float* conv_b, const float* pw_w, const float* pw_b,
                          const float* alpha0, const float* alpha1, const int* dilations, const float* alpha_f,
                          const float* w_f, float bias_f, int C, l3ac_tail_plan** plan_out);
int l3ac_tail_plan_destroy(l3ac_tail_plan* plan);
int l3ac_decoder_tail_tc(const l3ac_tail_plan* plan, const float* x, int B, int T, float* out, l3ac_stream_t stream);
/* The same kernel with 3-term split-bf16 operands for every contraction (snake outputs and weights as (hi, lo) bf16 pairs:
 * hi*Whi + lo*Whi + hi*Wlo, fp32 accumulation in TMEM) and a libm-accurate tanh: fp32-class results for precision "split".
 * Same plan, same arguments. */
int l3ac_decoder_tail_tc_split(const l3ac_tail_plan* plan, const float* x, int B, int T, float* out, l3ac_stream_t stream);

/* ==========================================================================================
 * Step-level interface: the two methods of the hot path as two calls.
 *   l3ac_encode = L3AC.encode_audio (l3ac/__init__.py:108-114): preprocess -> Encoder -> LocalEncoder /
 *                 CompressedLocalEncoderWithCache -> VQEmbed;
 *   l3ac_decode = L3AC.decode_audio (l3ac/__init__.py:116-121): VQEmbed.to_features -> LocalDecoder /
 *                 CompressedLocalDecoderWithCache -> Decoder.
 * A handle (l3ac_codec) owns the packed weights and the launch sequence over the operator-level entry points above, on the
 * tensor cores: encode side 3-term split-bf16, decode side bf16 (L3AC_PRECISION_BF16) or 3-term split-bf16
 * (L3AC_PRECISION_SPLIT), fp32 accumulation and residual stream.
 *
 * l3ac_create takes the reference's checkpoint as HOST fp32 tensors named "<module>.<state_dict key>" with module in
 * encoder / quantizer / decoder / en_encoder / en_decoder (the five <module>.pt files of l3ac/xtract/nn/module.py:36-54,
 * weight-norm pairs parametrizations.weight.original0/1 included) and the ModelConfig fields of the TOML
 * (l3ac/codec.py:13-36, l3ac/en_codec.py:9-19); it folds / packs them and uploads them to the CURRENT device.
 * Restrictions (L3AC_EUNSUPPORTED otherwise): feature_dim 128, first encoder / last decoder width 24, base_unit 'normal',
 * use_norm, use_snake_act, decoder_last_layer 'legacy', en_coder_dynamic_pos true, en_coder_cache_size 0 -- the four
 * published configs.  l3ac_last_error() names the offending tensor / field (thread-local string).
 *
 * l3ac_encode / l3ac_decode take DEVICE pointers, allocate nothing, never synchronise and enqueue everything on `stream`
 * (graph-capturable).  Activations live in `workspace` (16-byte aligned device memory of at least
 * l3ac_workspace_bytes(codec, B, T) bytes, T in samples; one workspace per concurrently running call).
 *   l3ac_encode: audio (B, T) fp32, any T >= 1 (right-padded to a multiple of the hop like Codec.preprocess) ->
 *                indices (B, T_tok) int32, T_tok = ceil(T / hop); q_feature (B, T_tok, feature_dim) fp32 and
 *                level_indices (B, T_tok, n_levels) fp32 are optional (NULL).
 *   l3ac_decode: indices (B, T_tok) int32 / int64 (indices_are_i64) or, when q_feature != NULL, the quantized features
 *                (B, T_tok, feature_dim) fp32 -> audio (B, T_tok * hop) fp32.
 * l3ac_encode_host / l3ac_decode_host take HOST pointers (pinned memory makes the copies asynchronous): micro-batches of
 * <= 330 s of audio are uploaded, processed and downloaded on up to four internal streams so that copies overlap kernels;
 * device staging and workspaces are owned by the handle and grow on demand; the calls return when the results are in
 * host memory.  Not re-entrant per handle.
 * ========================================================================================== */
#define L3AC_MAX_STAGES 8
#define L3AC_PRECISION_BF16 0  /* decode side bf16 operands (the default of the Python host, precision="bf16")               */
#define L3AC_PRECISION_SPLIT 1 /* decode side 3-term split-bf16 like the encode side: fp32-class waveform (precision="split") */
typedef struct l3ac_codec_config {
    int feature_dim;
    int n_encoder_stages;                 /* len(encoder_dims) = len(compress_rates) + 1 */
    int encoder_dims[L3AC_MAX_STAGES];
    int encoder_depths[L3AC_MAX_STAGES];
    int compress_rates[L3AC_MAX_STAGES];
    int en_coder_depth;
    int en_coder_window_size;
    int en_coder_compress_rate;
    int en_coder_dynamic_pos;
    int n_levels;                         /* vq_config.levels */
    int levels[8];
    int n_decoder_stages;                 /* len(decoder_dims) = len(decode_rates) + 1 */
    int decoder_dims[L3AC_MAX_STAGES];
    int decoder_depths[L3AC_MAX_STAGES];
    int decode_rates[L3AC_MAX_STAGES];
    int precision;                        /* L3AC_PRECISION_* */
} l3ac_codec_config;

typedef struct l3ac_tensor {
    const char* name;   /* "<module>.<state_dict key>", e.g. "decoder.blocks.0.parametrizations.weight.original1" */
    const float* data;  /* host, fp32, contiguous */
    long long numel;
} l3ac_tensor;

typedef struct l3ac_codec l3ac_codec;

int l3ac_create(const l3ac_codec_config* config, const l3ac_tensor* tensors, int n_tensors, l3ac_codec** codec_out);
int l3ac_destroy(l3ac_codec* codec);
const char* l3ac_last_error(void);
int l3ac_hop_length(const l3ac_codec* codec);
long long l3ac_workspace_bytes(const l3ac_codec* codec, int B, int T);
long long l3ac_launch_count(const l3ac_codec* codec); /* kernels launched through this handle so far */
int l3ac_encode(l3ac_codec* codec, const float* audio, int B, int T, void* workspace, long long workspace_bytes,
                float* q_feature, int32_t* indices, float* level_indices, l3ac_stream_t stream);
int l3ac_decode(l3ac_codec* codec, const void* indices, int indices_are_i64, const float* q_feature, int B, int T_tok,
                void* workspace, long long workspace_bytes, float* audio, l3ac_stream_t stream);
/* The quantiser alone with the handle's weights: VQEmbed.forward (trans_feature (B,T_tok,feature_dim) fp32 -> q_feature,
 * indices int32, level_indices fp32 or NULL) and VQEmbed.to_features (indices int32 / int64 -> q_feature);
 * l3ac/vq/__init__.py:20-30.  Device pointers, no workspace. */
int l3ac_quantize(l3ac_codec* codec, const float* trans_feature, int B, int T_tok, float* q_feature, int32_t* indices,
                  float* level_indices, l3ac_stream_t stream);
int l3ac_dequantize(l3ac_codec* codec, const void* indices, int indices_are_i64, int B, int T_tok, float* q_feature,
                    l3ac_stream_t stream);
int l3ac_encode_host(l3ac_codec* codec, const float* audio, int B, int T, int32_t* indices, float* q_feature);
int l3ac_decode_host(l3ac_codec* codec, const int32_t* indices, int B, int T_tok, float* audio);

#ifdef __cplusplus
}
#endif
#endif /* L3AC_B200_H_ */
