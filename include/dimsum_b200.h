/*
 * dimsum_b200 -- C ABI of the B200-native DiMSUM Mamba hot path (libdimsum_b200.so).
 *
 * Plain C: pointers, sizes, element strides, a CUDA stream handle.  No torch / ATen types.
 * Every entry point validates its arguments, launches asynchronously on `stream`, never
 * allocates device memory, never synchronises, and returns 0 or a negative dimsum_status;
 * dimsum_last_error() gives the message of the calling thread's last failure.
 *
 * Which reference interface each entry replaces (paths relative to the reference tree):
 *
 *   dimsum_selective_scan_fwd   selective_scan_fwd()  mamba/csrc/selective_scan/selective_scan.cpp:226-336
 *                               (pybind `selective_scan_cuda.fwd`, :495; params POD selective_scan.h:26-69)
 *   dimsum_selective_scan_bwd   selective_scan_bwd()  selective_scan.cpp:338-492 (`.bwd`, :496; selective_scan.h:71-101)
 *   dimsum_causal_conv1d_fwd    causal_conv1d_fwd() / causal_conv1d_fwd_cond()
 *                               causal-conv1d/csrc/causal_conv1d.cpp:221-281, :283-347 (POD causal_conv1d.h:9-35)
 *   dimsum_causal_conv1d_bwd    causal_conv1d_bwd() / causal_conv1d_bwd_cond()  causal_conv1d.cpp:349-427, :429-510
 *   dimsum_conv_xproj_fwd       causal_conv1d_fwd[_cond] + the x_proj F.linear + the B / C rearranges of MambaInnerFn*.forward,
 *                               mamba/mamba_ssm/ops/selective_scan_interface.py:836-866 (one tcgen05 kernel)
 *   dimsum_attention_fwd        F.scaled_dot_product_attention at dimsum/attention_fusion.py:61-84 and in the shared DiTBlock
 *                               (dimsum/models_dim.py:1532-1554), 256-token sequences, fp32 I/O with TF32 tensor cores
 *   dimsum_token_gather         torch.gather on token orders, mamba/mamba_ssm/modules/mamba_simple.py:634,657 and the
 *                               rearrange/flip/local_scan copies of dimsum/models_dim.py:1498-1524, :660-664,:700-701
 *   dimsum_wavelet_packet_fwd   WaveDiMBlock._dwt_fast  dimsum/models_dim.py:572-586 (+ local_scan :662)
 *   dimsum_wavelet_packet_inv   WaveDiMBlock._idwt_fast dimsum/models_dim.py:588-604 (+ local_reverse :701)
 *   dimsum_modulate / dimsum_gate_residual / dimsum_add_rmsnorm
 *                               the adaLN modulate, gated residual and fused add+RMSNorm passes around every mixer call
 *                               (models_dim.py:34-35,1079-1098,1509-1512; layernorm.py:460), with the token order folded in
 *
 * All strides are in ELEMENTS of the tensor's own dtype.  The innermost (sequence) stride of every
 * (batch, dim, seqlen) tensor is 1, as the reference requires (selective_scan.cpp:252-253).
 */
#ifndef DIMSUM_B200_H_
#define DIMSUM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIMSUM_ABI_VERSION 1

typedef enum {
    DIMSUM_OK = 0,
    DIMSUM_ERR_INVALID = -1,       /* bad argument (shape, stride, alignment, null pointer)      */
    DIMSUM_ERR_UNSUPPORTED = -2,   /* legal in the reference, not implemented here (no fallback) */
    DIMSUM_ERR_CUDA = -3           /* CUDA runtime reported an error at launch                   */
} dimsum_status;

typedef enum { DIMSUM_F32 = 0, DIMSUM_F16 = 1, DIMSUM_BF16 = 2 } dimsum_dtype;

/* ---- selective scan ------------------------------------------------------------------------
 * u, delta, z, out, out_z : (batch, dim, seqlen) io_dtype, arbitrary batch / dim strides
 * A                       : (dim, dstate) fp32
 * B, C                    : (batch, n_groups, dstate, seqlen) io_dtype ("variable" B and C)
 * D, delta_bias           : (dim) fp32 or NULL
 * x                       : (batch, dim, n_chunks, 2*dstate) fp32 contiguous or NULL; chunk_len = 32 and
 *                           n_chunks = ceil(seqlen/32) + 1.  Record c < n_chunks-1 holds, planar, the state h after
 *                           step 32c+16 in [0, dstate) and after step 32c+32 in [dstate, 2*dstate): the 16-step
 *                           checkpoints the backward restarts from.  The LAST record keeps the reference's
 *                           interleaved convention (odd slots = final state), so the reference's
 *                           last_state = x[:, :, -1, 1::2] (selective_scan_interface.py:39) holds.  (The reference
 *                           stores (prod a, h) per 2048-step chunk, selective_scan_fwd_kernel.cuh:251-254; only its
 *                           own backward kernel, replaced here, reads those.)
 * out                     : pre-gate y (needed by backward) or NULL (inference)
 * z / out_z               : both NULL or both set; out_z = y * silu(z)
 * perm                    : NULL, or int32[seqlen]: z is read at, and out_z written to, token perm[l]
 *                           (scan-order gather folded into the kernel; inference only)
 * a_is_arithmetic         : the CALLER asserts that A[d][n] == (n+1) * A[d][0] for every row (the S4D-real
 *                           initialisation A = -(1..N), mamba_simple.py:514-521); the kernel then derives the 16
 *                           decays of a step from one exp.  Not checked on the device.  Honoured only for dstate 16.
 */
typedef struct {
    int64_t batch, dim, seqlen, dstate, n_groups, n_chunks, chunk_len;
    int64_t io_dtype, delta_softplus;
    int64_t a_is_arithmetic;                 /* see above; 0 = general A */
    int64_t u_batch_stride, u_d_stride;
    int64_t delta_batch_stride, delta_d_stride;
    int64_t z_batch_stride, z_d_stride;
    int64_t out_batch_stride, out_d_stride;
    int64_t out_z_batch_stride, out_z_d_stride;
    int64_t A_d_stride, A_dstate_stride;
    int64_t B_batch_stride, B_group_stride, B_dstate_stride;
    int64_t C_batch_stride, C_group_stride, C_dstate_stride;
    const void *u, *delta, *A, *B, *C, *D, *z, *delta_bias;
    const int32_t *perm;
    void *out, *out_z, *x;
} dimsum_scan_fwd_params;

int dimsum_selective_scan_fwd(const dimsum_scan_fwd_params *p, void *stream);

/* Backward.  dB, dC are fp32 (batch, n_groups, dstate, seqlen) and dA (dim,dstate), dD, ddelta_bias (dim)
 * fp32; ALL FIVE must be zero-initialised by the caller (they are accumulated with atomics, as in the
 * reference: selective_scan.cpp:460-466).  du, ddelta, dz in io_dtype with their own strides (dz may be a
 * view of a larger buffer, selective_scan_interface.py:933-934).  out_z_recompute: optional recomputed
 * gated output (reference `recompute_out_z`).  x: the checkpoint tensor produced by the forward.
 */
typedef struct {
    int64_t batch, dim, seqlen, dstate, n_groups, n_chunks, chunk_len;
    int64_t io_dtype, delta_softplus;
    int64_t u_batch_stride, u_d_stride;
    int64_t delta_batch_stride, delta_d_stride;
    int64_t z_batch_stride, z_d_stride;
    int64_t out_batch_stride, out_d_stride;
    int64_t dout_batch_stride, dout_d_stride;
    int64_t du_batch_stride, du_d_stride;
    int64_t ddelta_batch_stride, ddelta_d_stride;
    int64_t dz_batch_stride, dz_d_stride;
    int64_t out_z_batch_stride, out_z_d_stride;
    int64_t A_d_stride, A_dstate_stride;
    int64_t B_batch_stride, B_group_stride, B_dstate_stride;
    int64_t C_batch_stride, C_group_stride, C_dstate_stride;
    int64_t dB_batch_stride, dB_group_stride, dB_dstate_stride;
    int64_t dC_batch_stride, dC_group_stride, dC_dstate_stride;
    const void *u, *delta, *A, *B, *C, *D, *z, *delta_bias, *dout, *out, *x;
    void *du, *ddelta, *dz, *out_z_recompute;
    float *dA, *dB, *dC, *dD, *ddelta_bias;
} dimsum_scan_bwd_params;

int dimsum_selective_scan_bwd(const dimsum_scan_bwd_params *p, void *stream);

/* ---- causal depthwise conv1d (+ optional SiLU) ----------------------------------------------
 * x, out : (batch, dim, seqlen) io_dtype; weight (dim, width) and bias (dim) in w_dtype; width 2..4.
 * perm   : NULL or int32[seqlen]; the conv runs over the permuted sequence x[:, :, perm[l]]
 *          (token-order gather folded into the load, no permuted copy of x).
 */
typedef struct {
    int64_t batch, dim, seqlen, width;
    int64_t io_dtype, w_dtype, silu;
    int64_t x_batch_stride, x_d_stride;
    int64_t out_batch_stride, out_d_stride;
    int64_t w_d_stride, w_width_stride;
    const void *x, *weight, *bias;
    const int32_t *perm;
    void *out;
} dimsum_conv_fwd_params;

int dimsum_causal_conv1d_fwd(const dimsum_conv_fwd_params *p, void *stream);

/* dweight (dim,width) and dbias (dim) are fp32 and must be zero-initialised (atomics, causal_conv1d.cpp:402-405). */
typedef struct {
    int64_t batch, dim, seqlen, width;
    int64_t io_dtype, w_dtype, silu;
    int64_t x_batch_stride, x_d_stride;
    int64_t dout_batch_stride, dout_d_stride;
    int64_t dx_batch_stride, dx_d_stride;
    int64_t w_d_stride, w_width_stride;
    const void *x, *weight, *bias, *dout;
    void *dx;
    float *dweight, *dbias;
} dimsum_conv_bwd_params;

int dimsum_causal_conv1d_bwd(const dimsum_conv_bwd_params *p, void *stream);

/* ---- causal conv1d + SiLU fused in front of the x_proj contraction (tcgen05 tensor cores) ------------------
 * Replaces the first half of MambaInnerFn*.forward (mamba/mamba_ssm/ops/selective_scan_interface.py:836-866):
 *     conv1d_out = causal_conv1d_cuda.causal_conv1d_fwd[_cond](x, conv1d_weight, conv1d_bias, True[, init_states])
 *     x_dbl      = F.linear(rearrange(conv1d_out, "b d l -> (b l) d"), x_proj_weight)
 *     B, C       = rearrange(x_dbl[:, ...], "(b l) dstate -> b 1 dstate l").contiguous()
 * x            : (batch, dim, seqlen) io_dtype, sequence stride 1 (the first half of xz)
 * conv_weight  : (dim, 4) contiguous fp32, conv_bias (dim) fp32 (the model's conv; anything else is DIMSUM_ERR_UNSUPPORTED);
 *                SiLU always applied
 * x_proj_weight: (n_out, dim) io_dtype, row stride xw_row_stride, n_out = dt_rank + 2*dstate (multiple of 8, <= 128)
 * x_proj_weight_lo : precision 1 only: x_proj_weight - tf32_truncate(x_proj_weight), same shape and strides (fp32)
 * u            : (batch, dim, seqlen) io_dtype  = silu(conv(x)), bit-identical to dimsum_causal_conv1d_fwd
 * x_dbl        : x_proj(u), CHANNEL-major, io_dtype: output row e (< n_out) of batch b, token l is written to
 *                x_dbl + b * x_dbl_batch_stride + e * x_dbl_row_stride + l for e < split_rows (or all e when x_dbl_tail is
 *                NULL) and to x_dbl_tail + b * tail_batch_stride + (e - split_rows) * tail_row_stride + l otherwise;
 *                split_rows is a multiple of 8.  The Python host puts the dt rows in a (dt_rank, batch * seqlen) matrix --
 *                the right operand of the dt_proj GEMM, selective_scan_interface.py:841 -- and the B / C rows in a
 *                (batch, 2*dstate, seqlen) tensor whose halves ARE B and C in the layout the scan reads: the two
 *                rearrange(...).contiguous() copies of :852,:862 disappear
 * precision    : 0 = one tensor-core pass (TF32 for fp32 I/O -- what cuBLAS does under allow_tf32 -- bf16 / fp16 operands for
 *                16-bit I/O, fp32 accumulation); 1 = 3xTF32 split for fp32 I/O (fp32-grade, ~1e-6 relative)
 * Requires dim % (128 / element size) == 0, seqlen % (16 / element size) == 0 and 16-byte aligned rows; anything else
 * returns DIMSUM_ERR_UNSUPPORTED and the caller runs the two separate steps.  x and x_proj_weight are read through TMA
 * tensor maps built per call (cuTensorMapEncodeTiled via cudaGetDriverEntryPoint: no link-time libcuda dependency).
 */
typedef struct {
    int64_t batch, dim, seqlen, width, n_out, split_rows;
    int64_t io_dtype, w_dtype, precision;
    int64_t x_batch_stride, x_d_stride;
    int64_t u_batch_stride, u_d_stride;
    int64_t x_dbl_batch_stride, x_dbl_row_stride;
    int64_t tail_batch_stride, tail_row_stride;
    int64_t w_d_stride, w_width_stride, xw_row_stride;
    const void *x, *conv_weight, *conv_bias, *x_proj_weight, *x_proj_weight_lo;
    void *u, *x_dbl, *x_dbl_tail;
} dimsum_conv_xproj_params;

int dimsum_conv_xproj_fwd(const dimsum_conv_xproj_params *p, void *stream);

/* ---- softmax attention (TMA + tcgen05, fp32 I/O, TF32 math, fp32 accumulation and softmax) -------------------------
 * Replaces F.scaled_dot_product_attention(q, k, v) at dimsum/attention_fusion.py:61-84 (CrossAttentionFusion, 8 heads x 64)
 * and in the shared DiTBlock (dimsum/models_dim.py:1532-1554, timm Attention, 16 heads x 64) for the 256-token sequences of
 * the 256px configuration.  q, k, v: (batch, heads, seqlen, 64) views with innermost stride 1 and arbitrary positive,
 * 16-byte aligned batch / head / token strides (the slices of a fused qkv projection are taken in place); out: written as
 * out[b, h, token, :] with its own strides -- pass the strides of a (batch, tokens, heads * 64) buffer to get the layout the
 * output projection reads.  No mask, no dropout (the model uses neither).  head_dim == 64, seqlen_k % 64 == 0 (online softmax
 * over key blocks of 128: any length, 256 tokens at 256px and 1024 at 512px).
 */
typedef struct {
    int64_t batch, heads, seqlen_q, seqlen_k, head_dim, dtype;
    int64_t q_batch_stride, q_head_stride, q_token_stride;
    int64_t k_batch_stride, k_head_stride, k_token_stride;
    int64_t v_batch_stride, v_head_stride, v_token_stride;
    int64_t out_batch_stride, out_head_stride, out_token_stride;
    float scale;
    const void *q, *k, *v;
    void *out;
} dimsum_attention_params;

int dimsum_attention_fwd(const dimsum_attention_params *p, void *stream);

/* ---- token-major gather: dst[b, l, :] = src[b, index[l], :]  for (batch, seqlen, channels) rows ---- */
typedef struct {
    int64_t batch, seqlen, channels, dtype;
    int64_t src_batch_stride, src_token_stride;
    int64_t dst_batch_stride, dst_token_stride;
    const void *src;
    const int32_t *index;
    void *dst;
} dimsum_gather_params;

int dimsum_token_gather(const dimsum_gather_params *p, void *stream);

/* ---- 2-level Haar wavelet packet with the reference's token/channel map -----------------------
 * src, dst : (batch, grid*grid tokens, channels), token-major, channel stride 1; channels % 16 == 0, grid % 4 == 0.
 * fwd : dst[b, pos[token'], c'] = coef ...   where pos = seq_of_token (NULL = identity) places the
 *       transformed token at its position in the window scan (local_scan fused into the store).
 * inv : reads src[b, pos[token'], c'] (local_reverse fused into the load) and reconstructs the image tokens.
 * Both kernels compute the raw +-1 butterfly times `scale`: the forward transform uses scale = 1/16
 * (two Haar levels of 1/2 each, then the reference's division by 4), the inverse uses scale = 1.
 * Gradients reuse the same two kernels (wavelet_layer.py:22-33,50-65): d(fwd) = inv with scale 1/16,
 * d(inv) = fwd with scale 1.
 */
typedef struct {
    int64_t batch, grid, channels, dtype;
    int64_t src_batch_stride, src_token_stride;
    int64_t dst_batch_stride, dst_token_stride;
    const void *src;
    const int32_t *pos;
    void *dst;
    float scale;
} dimsum_wavelet_params;

int dimsum_wavelet_packet_fwd(const dimsum_wavelet_params *p, void *stream);
int dimsum_wavelet_packet_inv(const dimsum_wavelet_params *p, void *stream);

/* ---- token-major elementwise glue around the mixer, with the token order folded in ----------
 * modulate      : dst[b, l, :] = x[b, idx[l], :] * (1 + scale[b, :]) + shift[b, :]          (models_dim.py:34-35 + order)
 * gate_residual : dst[b, l, :] = x[b, l, :] + gate[b, :] * m[b, idx[l], :]                  (models_dim.py:1510-1512 + un-order)
 * x, m, dst: (batch, seqlen, channels) with channel stride 1; shift/scale/gate: (batch, channels) with a row stride
 * (they are chunks of one adaLN GEMM output).  idx: NULL or int32[seqlen].  Mixed precision is allowed, as it occurs
 * under autocast: x_dtype (residual stream, typically fp32), aux_dtype (m, shift, scale, gate: the GEMM outputs) and
 * dst_dtype are independent; channels % 8 == 0.
 */
typedef struct {
    int64_t batch, seqlen, channels;
    int64_t x_dtype, aux_dtype, dst_dtype;
    int64_t x_batch_stride, x_token_stride;
    int64_t m_batch_stride, m_token_stride;
    int64_t dst_batch_stride, dst_token_stride;
    int64_t vec_row_stride;                  /* row stride of shift / scale / gate */
    const void *x, *m, *shift, *scale, *gate;
    const int32_t *idx;
    void *dst;
} dimsum_rowwise_params;

int dimsum_modulate(const dimsum_rowwise_params *p, void *stream);
int dimsum_gate_residual(const dimsum_rowwise_params *p, void *stream);

/* residual-add + RMSNorm (reference: Triton _layer_norm_fwd_1pass_kernel, mamba_ssm/ops/triton/layernorm.py:62-118,
 * maths of rms_norm_ref :32-47): res_out = x + residual (fp32), y = res_out * rsqrt(mean(res_out^2) + eps) * weight.
 * x: (rows, channels) `dtype`; residual / res_out: fp32 or NULL; weight fp32; y `dtype`.
 */
typedef struct {
    int64_t rows, channels, dtype;
    int64_t x_row_stride, y_row_stride;
    const void *x, *residual, *weight;
    void *y, *res_out;
    float eps;
} dimsum_rmsnorm_params;

int dimsum_add_rmsnorm(const dimsum_rmsnorm_params *p, void *stream);

/* residual-add + norm + adaLN modulate in one pass, for the two places a DiMSUM block runs them back to back
 * (dimsum/models_dim.py:1509-1512: `hidden = hidden + x; mlp(modulate(norm_2(hidden), shift, scale))`, and
 * :1079-1098 / :1532-1554: `modulate(LayerNorm(x), shift, scale)` of the shared DiT block and the final layer):
 *     h = x + residual (fp32; residual may be NULL)        res_out = h (fp32, may be NULL)
 *     n = h * rsqrt(mean(h^2) + eps) * weight              (norm_kind 0, RMSNorm, maths of rms_norm_ref layernorm.py:32-47)
 *     n = (h - mean(h)) * rsqrt(var(h) + eps)              (norm_kind 1, LayerNorm without affine, biased variance)
 *     y = n * (1 + scale[row / rows_per_batch]) + shift[row / rows_per_batch]      (shift = scale = NULL: y = n)
 * x (rows, channels) x_dtype; shift / scale (batch, channels) aux_dtype with row stride vec_row_stride; y y_dtype.
 */
typedef struct {
    int64_t rows, channels, rows_per_batch;
    int64_t x_dtype, aux_dtype, y_dtype, norm_kind;
    int64_t x_row_stride, y_row_stride, vec_row_stride;
    const void *x, *residual, *weight, *shift, *scale;
    void *y, *res_out;
    float eps;
} dimsum_norm_modulate_params;

int dimsum_norm_modulate(const dimsum_norm_modulate_params *p, void *stream);

/* GatedMLP inner activation (dimsum/mlp.py:65-70): y[r, :] = gelu_tanh(x[r, :H]) * x[r, H:2H]; x (rows, 2H), y (rows, H). */
typedef struct {
    int64_t rows, hidden, dtype;
    int64_t x_row_stride, y_row_stride;
    const void *x;
    void *y;
} dimsum_gelu_mul_params;

int dimsum_gelu_mul(const dimsum_gelu_mul_params *p, void *stream);

/* Sampler step: classifier-free-guidance combine of DiM.forward_with_cfg (dimsum/models_dim.py:1886-1902:
 * half = uncond + s (cond - uncond) on the first `channels` channels, the two halves of the batch made identical) fused with
 * the fixed-grid Euler update of the velocity ODE (dimsum/transport/integrators.py:98-111: x += dt v).
 *   model_out : (2 * half_batch, >= channels, H, W) in out_dtype, row (sample) stride out_row_stride; cond rows first
 *   x, x_new  : (2 * half_batch, channels, H, W) fp32 contiguous (x_new may alias x); dt: fp32 scalar in DEVICE memory (so the
 *               step can sit inside a CUDA graph whose grid position changes between replays)
 *   v_out     : NULL, or (2 * half_batch, channels, H, W) fp32: also store the guided drift v (both halves)
 */
typedef struct {
    int64_t half_batch, channels, hw, out_dtype, out_row_stride;
    float cfg_scale;
    const void *model_out, *x;
    const float *dt;
    void *x_new, *v_out;
} dimsum_cfg_euler_params;

int dimsum_cfg_euler_step(const dimsum_cfg_euler_params *p, void *stream);

/* ---- backward pieces of the glue (training) ------------------------------------------------ */
/* Column sums over the tokens of each batch row, the reductions autograd needs for adaLN modulate / gated residual
 * (d shift = sum_l g, d scale = sum_l g x, d gate = sum_l g m; reference: autograd through models_dim.py:34-35, 1509-1512):
 *     sum_g[b, c] = sum_l g[b, l, c]            sum_gx[b, c] = sum_l g[b, l, c] * x[b, x_idx[l], c]
 * g, x: (batch, seqlen, channels) with channel stride 1, any of the three dtypes; either output may be NULL (x may be NULL
 * when sum_gx is); outputs (batch, channels) in out_dtype with row stride out_row_stride.  channels % 4 == 0.
 */
typedef struct {
    int64_t batch, seqlen, channels;
    int64_t g_dtype, x_dtype, out_dtype;
    int64_t g_batch_stride, g_token_stride, x_batch_stride, x_token_stride, out_row_stride;
    const void *g, *x;
    void *sum_g, *sum_gx;
    const int32_t *x_idx;      /* optional (seqlen,) token table: x is read at row x_idx[l] (g at row l); NULL = identity */
} dimsum_colsum_params;

int dimsum_token_colsum(const dimsum_colsum_params *p, void *stream);

/* backward of dimsum_gelu_mul: dx[r, :H] = dy * x[r, H:] * gelu_tanh'(x[r, :H]), dx[r, H:] = dy * gelu_tanh(x[r, :H]). */
typedef struct {
    int64_t rows, hidden, dtype;
    int64_t x_row_stride, dy_row_stride, dx_row_stride;
    const void *x, *dy;
    void *dx;
} dimsum_gelu_mul_bwd_params;

int dimsum_gelu_mul_bwd(const dimsum_gelu_mul_bwd_params *p, void *stream);

/* backward of dimsum_add_rmsnorm (autograd through rms_norm_ref, layernorm.py:32-47):
 *     xhat = h * rstd,  rstd = rsqrt(mean(h^2) + eps),  h = x + residual (fp32, the forward's res_out)
 *     dh = (weight * dy - xhat * mean(weight * dy * xhat)) * rstd + dres_in          (dres_in may be NULL)
 *     dweight_partial[cta, c] = sum over the rows of that CTA of dy[r, c] * xhat[r, c]
 * dh is written as dx (dx_dtype) and, when dres_out != NULL, as fp32 dres_out.  h, dres_in, dres_out: (rows, channels) fp32
 * contiguous; dy: y_dtype with row stride dy_row_stride; dweight_partial: (n_partials, channels) fp32, one row per CTA, the
 * caller sums them (deterministic, no atomics).  channels % 4 == 0, channels <= 1024.
 */
typedef struct {
    int64_t rows, channels, n_partials;
    int64_t dy_dtype, dx_dtype;
    int64_t dy_row_stride, dx_row_stride;
    const void *h, *weight, *dy, *dres_in;
    void *dx, *dres_out, *dweight_partial;
    float eps;
} dimsum_rmsnorm_bwd_params;

int dimsum_add_rmsnorm_bwd(const dimsum_rmsnorm_bwd_params *p, void *stream);

/* ---- misc --------------------------------------------------------------------------------- */
int dimsum_abi_version(void);
const char *dimsum_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py `gpu_launches`). */
int64_t dimsum_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DIMSUM_B200_H_ */
