/*
 * fastdm_b200.h -- C ABI of libfastdm_b200.so: the B200-native (sm_100a) DiT transformer-block
 * hot path that replaces FastDM's pybind module `fastdm.cuda_ops`
 * (reference: csrc/torch_bindings.cpp:191-201, csrc/include/ops.h:9-32).
 *
 * Conventions
 *  - plain pointers + sizes only; every pointer is a DEVICE pointer unless stated otherwise.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *  - every entry point returns 0 on success, a negative FDM_ERR_* otherwise, and never throws;
 *    fdm_last_error() returns a thread-local human-readable message for the last failure
 *    (reference convention: TORCH_CHECK -> RuntimeError, csrc/torch_bindings.cpp:31-61; the Python
 *    wrappers in fastdm_b200/ops.py raise RuntimeError from the code + message).
 *  - no state is kept between calls except a per-device cache of TMA descriptors / attribute
 *    settings; all entry points are thread-safe under the caller's usual "one stream at a time".
 *  - there is NO CPU fallback: on a device that is not compute capability 10.x every compute entry
 *    point returns FDM_ERR_ARCH.
 */
#ifndef FASTDM_B200_H_
#define FASTDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDM_OK 0
#define FDM_ERR_ARG (-1)     /* bad shape / alignment / null pointer                       */
#define FDM_ERR_ARCH (-2)    /* device is not sm_100                                        */
#define FDM_ERR_CUDA (-3)    /* a CUDA runtime / driver call failed (message has the detail) */
#define FDM_ERR_UNSUPPORTED (-4)

/* element types of activations / outputs */
#define FDM_BF16 0
#define FDM_F16 1
#define FDM_F32 2
#define FDM_E4M3 3
#define FDM_S8 4

/* GEMM epilogue activation (applied after scales + bias, before the output rounding) */
#define FDM_ACT_NONE 0
#define FDM_ACT_GELU_TANH 1 /* F.gelu(approximate="tanh"): fastdm/layer/activations.py:38-41 */
#define FDM_ACT_GELU_ERF 2  /* F.gelu exact:              fastdm/model/flux.py:61             */

const char* fdm_last_error(void);
/* "fastdm_b200 <version> sm_100a" */
const char* fdm_version(void);
/* 0 if device `dev` can run the kernels (cc 10.x), FDM_ERR_ARCH otherwise. */
int fdm_check_device(int dev);

/* ---------------------------------------------------------------------------------------------
 * Family 3: memory-bound ops (reference: csrc/elmwise_ops.cu; semantics follow the torch backend,
 * fastdm/kernel/torch/*.py, which is the parity oracle -- see SURVEY.md section 8(a)).
 * ------------------------------------------------------------------------------------------- */

/* Per-token dynamic FP8 (e4m3fn) quantisation.
 * Replaces fp8_quant_(out, input, scale, None): csrc/elmwise_ops.cu:524-547, ops.h:12-14.
 * Semantics: fastdm/kernel/torch/quantize.py:45-67 (bit-exact codes and scales).
 *   in   [rows, cols] in_dtype (BF16|F16|F32), row stride in_row_stride elements, unit column stride
 *   out  [rows, cols] e4m3, contiguous
 *   scale[rows] fp32 */
int fdm_quant_fp8(const void* in, void* out, float* scale, int64_t rows, int64_t cols,
                  int64_t in_row_stride, int in_dtype, void* stream);

/* Per-token dynamic INT8 quantisation, symmetric (azp == NULL) or asymmetric (azp != NULL).
 * Replaces int8_quant_(out, input, scales, azp): csrc/elmwise_ops.cu:402-431, ops.h:9-11.
 * Semantics: fastdm/kernel/torch/quantize.py:7-43 (bit-exact codes, scales and zero points). */
int fdm_quant_int8(const void* in, int8_t* out, float* scale, int32_t* azp, int64_t rows,
                   int64_t cols, int64_t in_row_stride, int in_dtype, void* stream);

/* RMSNorm over the last dimension: out = dtype(x * rsqrt(mean(x^2) + eps)) * weight.
 * Replaces rms_norm_(out, input, weight, eps): csrc/elmwise_ops.cu:433-449, ops.h:15-18.
 * Semantics: fastdm/kernel/torch/norm.py:5-27. weight may be NULL (no scaling).
 *   in/out [rows, cols], row strides in elements (out may alias in). */
int fdm_rms_norm(const void* in, void* out, const void* weight, int64_t rows, int64_t cols,
                 int64_t in_row_stride, int64_t out_row_stride, float eps, int dtype, void* stream);

/* In-place rotary embedding on q and k with positions = arange(seq) (the only use on the hot
 * path: fastdm/kernel/cuda/rotemb.py:36).
 * Replaces rotary_emb_(positions, query, key, head_size, cos_sin_cache, is_neox):
 * csrc/elmwise_ops.cu:451-522, ops.h:20-32. Semantics: fastdm/kernel/torch/rotemb.py:5-64
 * (every product / sum rounded to `dtype`).
 *   q [batch, seq, q_heads*head_size] with strides (q_batch_stride, q_token_stride, 1)
 *   k [batch, seq, k_heads*head_size] likewise (k may be NULL)
 *   cos_sin [>=seq, head_size] `dtype`, row stride cs_row_stride: cos(head_size/2) || sin(head_size/2) */
int fdm_rope(void* q, void* k, const void* cos_sin, int64_t batch, int64_t seq, int q_heads,
             int k_heads, int head_size, int64_t q_batch_stride, int64_t q_token_stride,
             int64_t k_batch_stride, int64_t k_token_stride, int64_t cs_row_stride, int is_neox,
             int dtype, void* stream);

/* Fused q/k RMSNorm + RoPE, in place on a fused qkv projection buffer -- the adjacent ops of
 * Attention.forward (fastdm/layer/transformer.py:275-298: slice, .contiguous(), rms_norm x2, rope)
 * and WanAttention.forward (:490-499) in one pass over q and k. Bit-identical to running
 * fdm_rms_norm then fdm_rope.
 *   buf      [tokens, ...] rows `token_stride` elements apart; q heads start at column q_offset,
 *            k heads at column k_offset (v, wherever it is, is untouched)
 *   across_heads = 0: each head is normalised on its own, wq/wk are [head_size]   (FLUX/SD3/Qwen)
 *   across_heads = 1: one norm over all q (k) heads of a token, wq/wk are [heads*head_size] (Wan)
 *   wq / wk  NULL => that tensor is not normalised;  cos_sin NULL => no rotation
 *   pos0     cache row of token 0 (text tokens precede image tokens in a joint sequence)
 *   rotation is the interleaved (is_neox = 0) form, the only one on the hot path */
int fdm_qk_norm_rope(void* buf, const void* wq, const void* wk, const void* cos_sin, int64_t tokens,
                     int q_heads, int k_heads, int head_size, int64_t token_stride, int64_t q_offset,
                     int64_t k_offset, int64_t pos0, int64_t cs_row_stride, float eps,
                     int across_heads, int dtype, void* stream);

/* LayerNorm (no affine, eps) * mul + add, fused with the per-token quantisation of the quantised
 * linear that consumes it -- the AdaLN "modulate" in front of every qkv / ff1 / proj_mlp GEMM
 * (SURVEY.md 8(f) item 1).
 *   round_steps = 1 (FLUX/SD3/Qwen, bf16 tensor ops): y = T(T(T(LN(x)) * mul) + add), mul = T(1+scale)
 *       fastdm/layer/normalization.py:191-199,228-234; fastdm/model/flux.py:156-158,170-171
 *   round_steps = 0 (Wan, fp32 chain):              y = T(LN(x) * mul + add)
 *       fastdm/model/wan.py:95,108 (mul = 1+scale, add = shift) and :101 (mul = weight, add = bias)
 *   mul / add  [batches, cols] or NULL, both of mod_dtype: FDM_F32, or FDM_BF16 (round_steps = 1 only: the
 *              reference evaluates (1 + scale) and shift in the tensor dtype, so bf16 loses nothing and the
 *              chain runs as native packed bf16 instructions); row r uses batch r / rows_per_batch
 *   out_dtype  FDM_E4M3 -> out + scale[rows]; FDM_S8 -> out + scale + azp (asymmetric);
 *              anything else -> no quantised output (y_out required)
 *   y_out      optional bf16 copy of y (row stride y_row_stride), NULL to skip
 * The quantised codes equal fdm_quant_*(y) bit for bit. */
int fdm_layernorm_modulate_quant(const void* in, const void* mul, const void* add, void* out,
                                 float* scale, int32_t* azp, void* y_out, int64_t rows, int64_t cols,
                                 int64_t in_row_stride, int64_t y_row_stride, int64_t rows_per_batch,
                                 float eps, int round_steps, int in_dtype, int out_dtype, int mod_dtype,
                                 void* stream);

/* out[r, :d] = x[r, :d] * gelu_erf(x[r, d:2d])   (second half gated).
 * The reference has no CUDA kernel for this op (fastdm/kernel/cuda/gelumul.py:17 raises; the
 * dispatcher forces Triton, fastdm/kernel/operators_set.py:54). Semantics:
 * fastdm/kernel/torch/gelumul.py:4-16. */
int fdm_gelu_and_mul(const void* in, void* out, int64_t rows, int64_t d, int64_t in_row_stride,
                     int64_t out_row_stride, int dtype, void* stream);

/* Elementwise GELU (act = FDM_ACT_GELU_TANH | FDM_ACT_GELU_ERF) fused with the per-token FP8/INT8
 * quantisation of the NEXT QLinear (the un-fused F.gelu + quant passes).
 * out_dtype FDM_E4M3: scale[rows]; FDM_S8: scale[rows] + azp[rows] (asymmetric).
 * Equals quant(gelu(x).to(dtype)) of the two reference ops bit for bit. */
int fdm_gelu_quant(const void* in, void* out, float* scale, int32_t* azp, int64_t rows,
                   int64_t cols, int64_t in_row_stride, int act, int in_dtype, int out_dtype,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Family 1: W8A8 GEMMs, per-token (row) activation scale x per-channel (column) weight scale.
 * ------------------------------------------------------------------------------------------- */

/* D[m,n] = out_dtype( act( sA[m]*sB[n]*sum_k A[m,k]*B[k,n] + bias[n] ) ), fp32 accumulation.
 * Replaces fp8_scaled_mm_(a, b, scales_a, scales_b, out_dtype, bias): csrc/torch_bindings.cpp:24-84.
 * Semantics: fastdm/kernel/torch/matrixmul.py:7-35.
 *   a  [M,K] e4m3 row-major, row stride lda (bytes == elements), 16-byte aligned rows
 *   b  [K,N] e4m3 COLUMN-major: element (k,n) at b[n*ldb + k] (reference: torch_bindings.cpp:36)
 *   scale_a[M], scale_b[N] fp32; bias[N] in out_dtype or NULL; d [M,N] row-major, row stride ldd
 *   out_dtype: FDM_BF16 | FDM_F16.   Requires K % 16 == 0, N % 8 == 0. */
int fdm_gemm_fp8(const void* a, const void* b, const float* scale_a, const float* scale_b,
                 const void* bias, void* d, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldd, int out_dtype, int act, void* stream);

/* D[m,n] = out_dtype( (acc_i32[m,n] - azp[m]*azp_adj[n]) * sA[m]*sB[n] ) + bias[n]
 * (rounded to out_dtype BEFORE the bias add, as the oracle does; azp/azp_adj NULL => symmetric).
 * Replaces int8_scaled_mm_(a, b, scales_a, scales_b, out_dtype, azp_adj, azp, bias):
 * csrc/torch_bindings.cpp:86-160. Semantics: fastdm/kernel/torch/matrixmul.py:37-74.
 * Layouts as fdm_gemm_fp8 with int8 operands; azp_adj[N] int32 (weight column sums,
 * fastdm/layer/qlinear.py:49), azp[M] int32. */
int fdm_gemm_int8(const void* a, const void* b, const float* scale_a, const float* scale_b,
                  const int32_t* azp_adj, const int32_t* azp, const void* bias, void* d, int64_t M,
                  int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd, int out_dtype,
                  int act, void* stream);

/* The same GEMMs with the block's gate / residual chain folded into the epilogue (SURVEY.md 8(f)
 * item 2): d = T( residual + G ), G = gate[row / rows_per_batch, n] * T(linear) (rounded to T when
 * round_steps = 1, i.e. bf16 tensor ops: fastdm/model/flux.py:153-154,161-163,69-72; kept in fp32
 * when round_steps = 0: fastdm/model/wan.py:97,112). gate fp32 [batches, N] or NULL (plain residual
 * add, wan.py:105); residual [M, N] out_dtype with row stride ldr or NULL; d may alias residual. */
int fdm_gemm_fp8_residual(const void* a, const void* b, const float* scale_a, const float* scale_b,
                          const void* bias, void* d, int64_t M, int64_t N, int64_t K, int64_t lda,
                          int64_t ldb, int64_t ldd, int out_dtype, int act, const float* gate,
                          const void* residual, int64_t ldr, int64_t rows_per_batch, int round_steps,
                          void* stream);
int fdm_gemm_int8_residual(const void* a, const void* b, const float* scale_a, const float* scale_b,
                           const int32_t* azp_adj, const int32_t* azp, const void* bias, void* d,
                           int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd,
                           int out_dtype, int act, const float* gate, const void* residual,
                           int64_t ldr, int64_t rows_per_batch, int round_steps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Family 2: attention. Non-causal multi-head attention, token-major ("NHD") layouts.
 * ------------------------------------------------------------------------------------------- */

/* O = softmax(scale * Q K^T [+ block mask]) V.
 * Replaces the `sdpa` / `sdpa_sparse` cuda-backend routes (fastdm/kernel/cuda/attention.py:149-261,
 * i.e. cuDNN SDPA / sageattention / spas_sage_attn) and flash_attention_fp8_fwd_
 * (csrc/torch_bindings.cpp:162-189). Semantics: fastdm/kernel/torch/attention.py:7-43 and the fp32
 * reference tests/test_attention.py:23-63.
 *   q [B, Sq, H, hd], k/v [B, Sk, H, hd]: element (b,s,h,d) at base + b*batch_stride + s*token_stride + h*hd + d
 *   o [B, Sq, H, hd] with strides (o_batch_stride, o_token_stride), unit-stride head/dim: may be a
 *     column slice of a wider buffer (FLUX single block: cat([attn, mlp]) without the cat);
 *     always BF16 (F16 when qkv_dtype is F16)
 *   qkv_dtype: FDM_BF16 | FDM_F16 | FDM_E4M3 (fp8: per-tensor descale 1.0, P quantised to e4m3
 *              unscaled -- the reference's only fp8 semantics, csrc/attention/interface.cu:262-270;
 *              head_dim 128 only, output bf16)
 *   hd in {64, 128}
 *   block_mask: NULL (dense) or int8 [B, H, ceil(Sq/mask_bq), ceil(Sk/mask_bk)], 1 = compute,
 *               0 = skip (excluded from the softmax); mask_bq in {64,128}, mask_bk in {64,128}
 *               (reference geometry: fastdm/kernel/cuda/attention.py:89-103). */
int fdm_attn_fwd(const void* q, const void* k, const void* v, void* o, const int8_t* block_mask,
                 int64_t B, int64_t Sq, int64_t Sk, int H, int hd, int64_t q_batch_stride,
                 int64_t q_token_stride, int64_t k_batch_stride, int64_t k_token_stride,
                 int64_t v_batch_stride, int64_t v_token_stride, int64_t o_batch_stride,
                 int64_t o_token_stride, int mask_bq, int mask_bk, float scale, int qkv_dtype,
                 void* stream);

/* Debug aid (not part of the reference API): register a device buffer of 4*8*64 int64 that CTA
 * (0,0,0) of every following fdm_attn_fwd launch fills with clock64() stamps of its pipeline events
 * (tools/attn_trace.py prints the timeline); NULL disables it. */
/* fdm_attn_fwd whose epilogue SCATTERS the output rows to their owners (Ulysses sequence parallelism, no reference
 * counterpart): query row r is written to o_peers[r / rows_per_peer], row r % rows_per_peer, token stride o_ts -- each
 * pointer is the (peer-mapped, e.g. CUDA VMM / torch symmetric memory) address of that rank's [rows_per_peer, >= H*hd]
 * output buffer, already offset to this rank's head columns. The post-attention all-to-all and its unpack pass
 * disappear: the transfer is the kernel's own stores over NVLink, overlapped with the other CTAs' math. The caller
 * orders the peers' reads after every rank's launch (a cross-rank barrier on the stream). o_peers is a HOST array of
 * n_peers (<= 8) device pointers. Batch 1, head_dim 128, bf16; dense or block-sparse. */
int fdm_attn_fwd_scatter(const void* q, const void* k, const void* v, void* const* o_peers, int n_peers,
                         int64_t rows_per_peer, const int8_t* block_mask, int64_t Sq, int64_t Sk, int H, int hd,
                         int64_t q_ts, int64_t k_ts, int64_t v_ts, int64_t o_ts, int mask_bq, int mask_bk,
                         float scale, int qkv_dtype, void* stream);

int fdm_debug_set_attn_trace(void* device_buffer);

/* ---------------------------------------------------------------------------------------------
 * Family 4 helpers: Ulysses sequence-parallel layout kernels (no reference counterpart; the
 * exchange itself is NCCL all-to-all issued from the host side, fastdm_b200/ulysses.py).
 * ------------------------------------------------------------------------------------------- */

/* Pack the local token shard of n_seg head-major tensors (e.g. the q|k|v segments of a fused qkv
 * projection: n_seg = 3, segments src_seg_stride elements apart, rows src_token_stride apart) into
 * the send buffer of the pre-attention all-to-all: dst [P, S_local, n_seg, H/P, hd] contiguous,
 * chunk p = head group p of every segment. After the exchange the receive buffer IS
 * [P*S_local tokens, n_seg, H/P, hd] -- the full sequence for this rank's heads, ready for
 * fdm_attn_fwd with token stride n_seg*(H/P)*hd. elem_size in bytes (1, 2, 4); H % P == 0. */
int fdm_ulysses_pack_heads(const void* src, void* dst, int64_t S_local, int H, int hd, int P,
                           int n_seg, int64_t src_token_stride, int64_t src_seg_stride,
                           int elem_size, void* stream);

/* Inverse, for the post-attention all-to-all: src [P, S_local, n_seg, H/P, hd] (chunk p = head
 * group p, received from rank p) -> dst token-major [S_local, n_seg x (H, hd)]. */
int fdm_ulysses_unpack_heads(const void* src, void* dst, int64_t S_local, int H, int hd, int P,
                             int n_seg, int64_t dst_token_stride, int64_t dst_seg_stride,
                             int elem_size, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Step-cache indicator: the relative L1 distance TeaCache / FBCache / DiCache threshold on,
 *   (a - b).abs().mean() / b.abs().mean()        fastdm/caching/xcaching.py:214-215, 361-362, 479-480
 * as ONE pass over a and b instead of five full-size torch kernels: out2[0] = sum |T(a - b)| (T = rounding to
 * the tensor dtype, as the reference's bf16 subtraction does), out2[1] = sum |b|, both fp32, device memory
 * (zeroed by the call). n elements, contiguous, 16-byte aligned; dtype FDM_BF16 or FDM_F16.
 * ------------------------------------------------------------------------------------------- */
int fdm_rel_l1_distance(const void* a, const void* b, int64_t n, int dtype, float* out2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTDM_B200_H_ */
