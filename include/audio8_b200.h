/* audio8_b200 — C ABI of the B200-native wav2vec2 training hot path (libaudio8_b200.so).
 *
 * The reference (mead-ml/audio8) exposes no FFI: its hot path sits behind Python classes in
 * audio8/wav2vec2.py and audio8/ctc.py whose arithmetic is dispatched by PyTorch to ATen / cuDNN / cuBLAS
 * kernels.  This header is the boundary a drop-in replacement binds instead: each entry point names the
 * reference call site (file:line under /root/reference/audio8) whose library kernels it replaces.
 *
 * Conventions (all entry points):
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - no allocation, no synchronisation and no host<->device copies inside: the caller allocates outputs and
 *     workspaces and keeps them alive until the stream work completes;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns immediately;
 *   - return value 0 on success, negative on error; a8_last_error() then returns a thread-local message;
 *   - bf16 tensors are channels-last ([rows, channels], unit stride on channels); fp32 where stated.
 */
#ifndef AUDIO8_B200_H
#define AUDIO8_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A8_ABI_VERSION 3

int a8_version(void);
const char* a8_last_error(void);
/* number of kernels launched by this library in this process since load (bench.py reports the delta) */
int64_t a8_launch_count(void);
/* account for kernels launched by replaying a captured CUDA graph (the host-side counter cannot see them) */
int a8_launch_count_add(int64_t n);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 / TMEM / TMA GEMM core.
 * Replaces: every nn.Linear / Dense (`wav2vec2.py:932,950,951,762`, eight_mile Dense inside
 * TransformerEncoderStack `wav2vec2.py:613-622`), conv layers 1-6 of the feature encoder as implicit GEMM
 * (`wav2vec2.py:426-428`), the grouped positional conv (`wav2vec2.py:600-609,634`), the attention
 * score/context products (eight_mile SeqScaledDotProductAttention via `wav2vec2.py:644`), and all of their
 * data-gradient / weight-gradient products in backward.
 *
 *   C[hi][lo][m][n] = epilogue( alpha * sum_k A(hi,lo)[m][k] * B(hi,lo)[n][k] )
 *
 * Operands are bf16 tensors described as 4-D strided views (dims[0] has unit stride); a tile of an operand is
 * fetched by TMA at coordinates that are affine in the loop indices, so convolution windows, time shifts,
 * head/group offsets and batch indices need no materialised im2col:
 *   coord[d] = base[d] + ck[d]*kin + cb[d]*kbatch + cr[d]*r + cl[d]*lo + ch[d]*hi
 * with k-block index kb in [0,k_blocks), kin = kb % k_inner, kbatch = kb / k_inner, and
 *   major == A8_MAJOR_K : one box {64 k-elements, 128 (A) or block_n (B) rows} per k-block, r = first row (m0 / n0)
 *   major == A8_MAJOR_MN: boxes {64 mn-elements, 64 k-rows}, one per 64 rows of the tile, r = (row0/64 + i)
 * Out-of-range coordinates read as zero (TMA fill), which implements 'same' padding and ragged edges.
 * ---------------------------------------------------------------------------------------------- */
enum { A8_MAJOR_K = 0, A8_MAJOR_MN = 1 };
enum { A8_OUT_BF16 = 0, A8_OUT_F32 = 1, A8_OUT_F32_ATOMIC = 2 };
enum { A8_ACT_NONE = 0, A8_ACT_GELU = 1, A8_ACT_GELU_DZ = 2 };
enum { A8_AUX_NONE = 0, A8_AUX_ADD = 1, A8_AUX_MUL_GELU_GRAD = 2, A8_AUX_MUL = 3 };
enum { A8_GEMM_PAIR = 2, A8_GEMM_TAP_WINDOW = 3 };

typedef struct {
  const void* ptr;    /* bf16 */
  int64_t dims[4];    /* extents in elements; dims[0] is contiguous */
  int64_t strides[3]; /* element strides of dims 1..3 (positive multiples of 8) */
  int32_t major;      /* A8_MAJOR_* */
  int32_t base[4], ck[4], cb[4], cr[4], cl[4], ch[4];
} a8_operand_t;

typedef struct {
  a8_operand_t a, b;
  int32_t M, N;               /* valid rows / columns of each (hi, lo) output block */
  int32_t lo_count, hi_count; /* >= 1 */
  int32_t k_blocks, k_inner;  /* 64-wide k-blocks per output tile; k-blocks per k-batch */
  int32_t split_k;            /* > 1 requires A8_OUT_F32_ATOMIC into a zeroed C */
  int32_t block_n;            /* 0 = choose, else 64 / 128 / 256 */
  void* c;                    /* output; element offset = hi*c_stride_hi + lo*c_stride_lo + m*ldc + n */
  int32_t c_dtype;            /* A8_OUT_* */
  int32_t act;                /* A8_ACT_* applied after bias */
  int64_t ldc, c_stride_lo, c_stride_hi; /* multiples of 8 elements; rows are padded to 8 columns */
  void* z_out;                /* optional 2-byte side output addressed like C: the pre-activation (alpha*acc + bias) in bf16,
                                 or, with act = A8_ACT_GELU_DZ, gelu'(pre-activation) in IEEE fp16 (values in [-0.13, 1.13]:
                                 11 mantissa bits) - what the backward pass multiplies by (A8_AUX_MUL), so that its epilogue
                                 is one multiply instead of an erf + exp evaluation */
  const void* aux;            /* optional bf16 tensor addressed like C */
  int32_t aux_mode;           /* A8_AUX_ADD: out = act(..) + aux;  A8_AUX_MUL_GELU_GRAD: out = (..) * gelu'(aux);
                                 A8_AUX_MUL: out = (..) * aux with aux in IEEE fp16 (the factors A8_ACT_GELU_DZ wrote) */
  int32_t bias_stride_lo;
  const float* bias;          /* optional fp32, index lo*bias_stride_lo + n */
  float alpha;
  int32_t reserved;           /* launch-shape request, low byte: 0 = plain, A8_GEMM_PAIR = CTA pairs compute 256 x block_n
                                 tiles (cta_group::2), A8_GEMM_TAP_WINDOW = tap-window kernel for K-major operands whose A
                                 tile moves by exactly one row of dims[1] per k-block (cb = {0, +-1, 0, 0}, k_inner = 1,
                                 k_blocks a multiple of 8 in [16, 128], N <= 64, bf16 output: the grouped positional conv and
                                 its data gradient).  Same result as the plain kernel; the rows the taps share are staged
                                 once per 8 k-blocks instead of once per k-block.  Bits 8..15 (window only): how many of the
                                 four 16-element k-steps of a k-block hold non-zero B columns (0 = all), the rest is skipped. */
  float* colsum;              /* optional fp32 [N], accumulated into (zeroed by the caller): column sums over the M rows of
                                 the bf16 output as stored, i.e. the bias gradient of the layer whose output gradient this
                                 GEMM produces (`db1 = sum_rows(dY W2 * gelu')`).  Only with A8_AUX_MUL, bf16 output, no
                                 split-K, lo_count = hi_count = 1, plain / pair kernels. */
} a8_gemm_t;

int a8_gemm(const a8_gemm_t* p, void* stream);

/* One persistent launch over n (<= 48) GEMM problems that share operand majors, coordinate maps, block_n, cluster
 * request, split_k, k_inner, alpha and output type, each with its own operands, output, M, N and k_blocks (no bias / aux /
 * z_out / activation, lo_count = hi_count = 1).  Instantiated for (MN-major, MN-major) operands, i.e. weight gradients
 * dW = dY^T X: the transformer stack defers the 4 weight-gradient GEMMs of every layer (what autograd runs as 48 separate
 * addmm calls behind `/root/reference/audio8/wav2vec2.py:644`) to the end of its backward and issues them as one kernel. */
int a8_gemm_group(const a8_gemm_t* problems, int32_t n, void* stream);
/* The same in two steps, for callers that launch the same group repeatedly (fixed operand addresses): prepare encodes
 * the 2n tensor maps and the tile table into a caller-owned HOST buffer of a8_gemm_group_blob_bytes() bytes (~100 us of
 * host time for 48 problems), launch only enqueues the kernel. */
size_t a8_gemm_group_blob_bytes(void);
int a8_gemm_group_prepare(const a8_gemm_t* problems, int32_t n, void* blob);
int a8_gemm_group_launch(const void* blob, void* stream);
/* debug aid: later a8_gemm launches stamp clock64() timelines of their first CTAs into `buf` (device memory,
 * 4*3*8*4 int64); NULL turns it off.  Not used by the product path. */
void a8_gemm_set_trace(void* buf);

/* ------------------------------------------------------------------------------------------------
 * CTC loss (csrc/ctc_loss.cu).  Replaces `torch.nn.functional.ctc_loss` as called at `ctc.py:197-205`
 * (ATen ctc_loss_log_alpha / log_beta / collect kernels; cuDNN disabled by the reference) and, with from_logits != 0,
 * the `log_softmax` in front of it as well (`wav2vec2.py:770`): x then holds the classifier's logits, rows are
 * normalised on the fly and the backward pass returns d loss / d logits = (softmax - occupancy) * scale directly.
 * x: fp32 [T,B,V] with arbitrary element strides (the reference passes a transposed view, `train.py:39`).
 * targets: int32 concatenated labels (ctc.py:193-194 stripping is done by a8_ctc_prep), tgt_offsets[b] = start of
 * utterance b.  in_lengths: B ints followed by ONE int that is zero on entry (a8_ctc_prep writes it).
 * alpha, beta: fp32 scratch of a8_ctc_scratch_floats(T, B, max_S) floats each (log2 domain; alpha includes the emission
 * at t, beta holds the sum over successors without it).
 * a8_ctc_forward runs both sweeps concurrently (one CTA per utterance), fills alpha, beta, nll[b] (+inf if infeasible)
 * and, if loss != NULL, the reduced loss: sum_b nll_b, or mean_b(nll_b / max(S_b,1)) when reduction_mean; infinite rows
 * count as 0 when zero_infinity.  One launch.
 * a8_ctc_backward writes grad[t*grad_stride_t + b*grad_stride_b + v] = (exp(logp) - occupancy) * scale_b for t < len,
 * else 0 (PyTorch's convention; with from_logits this IS the gradient w.r.t. the logits), scale_b =
 * grad_out[b*grad_out_stride] (/ (max(S_b,1)*B) when reduction_mean); infeasible rows get 0.  One launch.
 * ---------------------------------------------------------------------------------------------- */
size_t a8_ctc_scratch_floats(int32_t T, int32_t B, int32_t max_S);
/* ctc.py:193-194 on the device, sync-free: flat[] = row-major compaction of targets[B,S] (int64, strided)
 * without PAD/EOS; tgt_offsets = exclusive cumsum(target_lengths); lengths converted to int32.
 * flat has room for B*S entries, row_start is B ints of scratch, in_lengths has room for B + 1 ints (the last one is
 * set to 0: the completion ticket a8_ctc_forward's loss reduction counts on). */
/* Greedy best-path decode, the reference's only alignment (`ctc.py:161-162`: argmax(-1).unique_consecutive(), blank
 * dropped): lp fp32 [B,T,V] with element strides, in_len int32 [B] (or NULL) -> out int32 [B,T] (decoded ids, then -1)
 * and out_len int32 [B].  Integer result, bit-exact against the reference's ops on the same log-probs. */
int a8_ctc_greedy(const float* lp, int64_t stride_b, int64_t stride_t, int64_t stride_v, int32_t B, int32_t T, int32_t V,
                  const int32_t* in_len, int32_t blank, int32_t* out, int32_t* out_len, void* stream);
int a8_ctc_prep(const int64_t* targets, int64_t stride_b, int64_t stride_s, int32_t B, int32_t S, int32_t pad,
                int32_t eos, const int64_t* target_lengths, const int64_t* input_lengths, int32_t* flat,
                int32_t* row_start, int32_t* tgt_offsets, int32_t* tgt_lengths, int32_t* in_lengths, void* stream);
int a8_ctc_forward(const float* x, int64_t stride_t, int64_t stride_b, int64_t stride_v, int32_t T, int32_t B, int32_t V,
                   int32_t from_logits, const int32_t* targets, const int32_t* tgt_offsets, const int32_t* tgt_lengths,
                   const int32_t* in_lengths, int32_t max_S, int32_t blank, int32_t reduction_mean,
                   int32_t zero_infinity, float* alpha, float* beta, float* nll, float* loss, void* stream);
int a8_ctc_backward(const float* x, int64_t stride_t, int64_t stride_b, int64_t stride_v, int32_t T, int32_t B, int32_t V,
                    int32_t from_logits, const int32_t* targets, const int32_t* tgt_offsets, const int32_t* tgt_lengths,
                    const int32_t* in_lengths, int32_t max_S, int32_t blank, const float* alpha, const float* beta,
                    const float* nll, const float* grad_out, int64_t grad_out_stride, int32_t reduction_mean, int32_t zero_infinity,
                    float* grad, int64_t grad_stride_t, int64_t grad_stride_b, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (+ residual add, + dropout), bf16 rows of C <= 1024 channels (C % 8 == 0), fp32 statistics.
 * Replaces nn.LayerNorm + residual `+` + nn.Dropout around it: `wav2vec2.py:904,930` (post-extractor LN),
 * `:623,638-640` (encoder LN after the positional conv) and ln1/ln2 of every transformer layer.
 *   fwd:  s = x + drop_h(h);  y = drop_y(LN(s)*gamma + beta);  h, s_out, y_f32 optional (NULL)
 *   bwd:  g = drop_y(dy (+ dy_f32));  ds = dLN(g);  dh = drop_h(ds) (optional);
 *         dgamma/dbeta/dbias_h (= column sum of dh, or of ds when dh is NULL) are ACCUMULATED (zero them first)
 * Dropout masks are regenerated from (seed, element index): Philox-4x32, nothing is stored.  The effective seed of
 * every dropout-capable launch is `seed argument + *seed_source`: a8_set_seed_source() names a 64-bit word in DEVICE
 * memory (or NULL = 0) that subsequent launches of the calling host thread read at run time, so a captured CUDA graph
 * draws fresh masks on each replay when the caller refreshes that word (torch's CUDA generator does, graph-safely).
 * ---------------------------------------------------------------------------------------------- */
int a8_set_seed_source(const void* dev_u64);
int a8_layernorm_fwd(const void* x, const void* h, float p_h, uint64_t seed_h, void* s_out, const float* gamma,
                     const float* beta, float eps, void* y, float* y_f32, float p_y, uint64_t seed_y, float* mean,
                     float* rstd, int32_t R, int32_t C, void* stream);
int a8_layernorm_bwd(const void* dy, const float* dy_f32, float p_y, uint64_t seed_y, const void* s,
                     const float* mean, const float* rstd, const float* gamma, void* ds, void* dh, float p_h,
                     uint64_t seed_h, float* dgamma, float* dbeta, float* dbias_h, int32_t R, int32_t C,
                     void* stream);

/* Fused scaled-dot-product attention, d_k = 64 (tcgen05 / TMEM; scores and probabilities never reach HBM).
 * Replaces eight_mile's SeqScaledDotProductAttention as called through `wav2vec2.py:644` and its autograd:
 *   P = dropout(softmax(scale * Q K^T + key mask));  ctx = P V
 * qkv bf16 [B,T,3*H*64] = the fused projection output (Q | K | V, head h at columns h*64 of each third);
 * key_keep uint8 [B,T] or NULL (0 = padded key: the reference's masked_fill(-1e9)); ctx bf16 [B,T,H*64];
 * lse fp32 [B,H,T] = log2-sum-exp2 of the scaled scores (saved for backward); delta fp32 [B,H,T] scratch;
 * dqkv bf16 [B,T,3*H*64] receives dQ | dK | dV.  Dropout keep decisions are a function of (seed + *seed_source,
 * b, h, q, k) with keep probability 1 - floor(pdrop*2^32)/2^32; a8_attn_dropmask writes them as bytes
 * [B,H,T,T] (tests only). */
int a8_attn_fwd(const void* qkv, const uint8_t* key_keep, void* ctx, float* lse, int32_t B, int32_t H, int32_t T,
                float scale, float pdrop, uint64_t seed, void* stream);
int a8_attn_bwd(const void* qkv, const uint8_t* key_keep, const void* ctx, const void* dctx, const float* lse,
                float* delta, void* dqkv, float* dbias, int32_t B, int32_t H, int32_t T, float scale, float pdrop,
                uint64_t seed, void* stream); /* dbias: optional fp32 [3D], accumulated into (zeroed by the caller):
                                                 column sums of dqkv as stored = the gradient of the fused QKV bias */
int a8_attn_dropmask(uint8_t* keep_out, int32_t B, int32_t H, int32_t T, float pdrop, uint64_t seed, void* stream);

/* out[c] += sum_r x[r][c]  (bias gradients; x bf16 [R, ld], out fp32 zeroed by the caller) */
int a8_colsum(const void* x, int64_t ld, int32_t R, int32_t C, float* out, void* stream);
/* element-wise dropout (nn.Dropout at `wav2vec2.py:934-935,713`), dtype 0 = fp32, 1 = bf16; its own backward */
int a8_dropout(const void* x, void* out, int32_t dtype, int64_t n, float p, uint64_t seed, void* stream);
/* dz = dy * gelu'(z) (bf16): backward of the GELU that a GEMM epilogue applied (`wav2vec2.py:422,428,607`) */
int a8_gelu_bwd(const void* dy, const void* z, void* dz, int64_t n, void* stream);
/* GELU backward with the stored derivative: out = dy * g elementwise, dy / out bf16, g = gelu'(z) in IEEE fp16 as written
 * by A8_ACT_GELU_DZ; n % 8 == 0 */
int a8_mul_dgelu(const void* dy, const void* g, void* out, int64_t n, void* stream);
/* F.log_softmax(-1) of `wav2vec2.py:770`: x fp32 [R,V] -> y fp32; bwd consumes a strided fp32 gradient
 * (element (row, c) at dy[(row / rows_inner)*stride_outer + (row % rows_inner)*stride_row + c*stride_v], so the
 * [T,B,V] CTC gradient needs no transpose) and writes dx bf16 [R,V] */
int a8_log_softmax_fwd(const float* x, float* y, int32_t R, int32_t V, void* stream);
int a8_log_softmax_bwd(const float* dy, int64_t stride_outer, int64_t stride_row, int64_t stride_v,
                       int32_t rows_inner, const float* y, void* dx, int32_t R, int32_t V, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Feature-encoder layer 0: Conv1d(1->C,k,stride, no bias) + GroupNorm(C,C) + GELU, fused (`wav2vec2.py:419-422`).
 * x fp32 [B,L]; w fp32 [C,k]; y / da bf16 [B,L0,C] channels-last, L0 = (L-k)/stride + 1.
 *   stats: per-(b,c) mean / rstd over time from the window moments of x (moments: 65*B doubles, kept for bwd)
 *   fwd:   y = gelu(((conv(x) - mean) * rstd) * gamma + beta)
 *   bwd:   given da = dL/dy: dw [C,k], dgamma, dbeta are WRITTEN.  One pass over da accumulates sum_t dy x_j,
 *          sum_t dy, sum_t dy xhat per (b,c) (acc: 12*B*C floats of scratch); the GroupNorm backward's mean
 *          corrections of dw follow from those and the forward's window moments.
 * ---------------------------------------------------------------------------------------------- */
int a8_conv0_stats(const float* x, int32_t B, int64_t L, const float* w, int32_t C, int32_t k, int32_t stride,
                   float eps, double* moments, float* mean, float* rstd, void* stream);
int a8_conv0_fwd(const float* x, int32_t B, int64_t L, const float* w, const float* gamma, const float* beta,
                 const float* mean, const float* rstd, int32_t C, int32_t k, int32_t stride, void* y, void* stream);
int a8_conv0_bwd(const float* x, int32_t B, int64_t L, const float* w, const float* gamma, const float* beta,
                 const float* mean, const float* rstd, const double* moments, int32_t C, int32_t k, int32_t stride,
                 const void* da, float* acc, float* dw, float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Time-mask plumbing (`wav2vec2.py:939,946,381,632,717,721`): index driven, no nonzero / host sync.
 * dtype codes: 0 = fp32, 1 = bf16.
 * ---------------------------------------------------------------------------------------------- */
int a8_rows_copy(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, const int32_t* idx, int32_t n,
                 int32_t C, int32_t scatter, void* stream); /* gather dst[i]=src[idx[i]] / scatter dst[idx[i]]=src[i];
                                                               idx[i] < 0 = padding entry: zero row / skipped (also in
                                                               a8_rows_set / a8_rows_set_bwd) */
int a8_rows_set(void* x, const int32_t* idx, int32_t n, int32_t C, const float* vec, void* stream);
int a8_rows_set_bwd(void* dx, const int32_t* idx, int32_t n, int32_t C, float* dvec, void* stream);
int a8_mask_apply(void* x, const uint8_t* row_keep, const uint8_t* chan_zero, int32_t B, int32_t T, int32_t C,
                  void* stream);
int a8_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream);
/* bf16x3 split of an fp32 matrix [R,C] -> bf16 [R,3C] so one bf16 GEMM reproduces an fp32-accurate product */
int a8_split3(const float* src, void* dst, int32_t R, int32_t C, int32_t b_side, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gumbel vector quantizer rows (`wav2vec2.py:547-576`; closed forms in SURVEY D.2).
 * z fp32 [R,G*V] logits; noise fp32 [R*G,V] = -log(Exp(1)) drawn by the caller exactly as F.gumbel_softmax does,
 * or NULL in eval mode; vars fp32 [G*V,vd].  fwd: kidx[R*G] = argmax, q[R,G*vd] = selected codewords (+ bf16
 * copy), avg_sums[V] = sum of softmax(z) rows pooled over groups, ppl = exp(-sum q log(q+1e-7)).
 * bwd: a_dot fp32 [R,G*V] = dq . vars^T per group (from a8_gemm); writes dz bf16 [R,G*V]; dvars ACCUMULATED.
 * ---------------------------------------------------------------------------------------------- */
/* n_valid (device int32 scalar, may be NULL = all rows): the row lists of a step are padded to their worst-case length R
 * so that every shape is static (CUDA-graph replay); only rows [0, *n_valid) exist.  Padding rows get a code / zero
 * gradient but stay out of the pooled statistics (fwd) and contribute nothing (bwd). */
int a8_vq_fwd(const float* z, const float* noise, float tau, const float* vars, int32_t R, int32_t G, int32_t V,
              int32_t vd, const int32_t* n_valid, float* q, void* q_bf16, int32_t* kidx, float* avg_sums, float* ppl,
              void* stream);
int a8_vq_bwd(const float* z, const float* noise, float tau, int32_t R, int32_t G, int32_t V, int32_t vd,
              const int32_t* n_valid, const float* a_dot, const float* dq, const int32_t* kidx, const float* avg_sums, const float* ppl,
              const float* dppl, void* dz, float* dvars, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Contrastive loss (`wav2vec2.py:377-392`, Sampler :955-976; closed forms in SURVEY D.3).
 * x, y fp32 [R,C] (context outputs at the masked steps, projected quantized targets), idx int32 [R,K] rows of y
 * (the reference's numpy-drawn negatives, already offset by b*Tm).  cos-sim logits over [positive | K negatives],
 * ce = mean_i CE(logits_i, 0), loss = xe_w*ce + div_w*(n_vars - *ppl)/n_vars (ppl may be NULL).
 * Scratch kept for backward: xn, yn [R]; cosv, prob [R,K+1]; row_loss [R].
 * bwd: dce = d loss / d ce (device scalar) -> dx [R,C], dy [R,C].
 * n_valid (device int32 scalar or NULL): rows [*n_valid, R) are padding - excluded from the mean, zero gradient.
 * ---------------------------------------------------------------------------------------------- */
int a8_contrastive_fwd(const float* x, const float* y, const int32_t* idx, int32_t R, int32_t C, int32_t K,
                       const int32_t* n_valid, const float* ppl, float n_vars, float xe_w, float div_w, float* xn, float* yn, float* cosv,
                       float* prob, float* row_loss, float* ce, float* loss, void* stream);
int a8_contrastive_bwd(const float* x, const float* y, const int32_t* idx, int32_t R, int32_t C, int32_t K,
                       const int32_t* n_valid, const float* xn, const float* yn, const float* cosv, const float* prob, const float* dce,
                       float* dx, float* dy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side draws of the pre-training step (opt-in, `wav2vec2.set_device_draws`): the span mask of
 * `create_mask` (`wav2vec2.py:189-216`) and the negative indices of `Sampler.negatives` (`wav2vec2.py:955-976`) from
 * Philox4x32-10 on the GPU instead of numpy's global generator on the host.  Same distributions as the reference
 * (uniform subset of span starts per row, every row cut down to the batch-minimum count by a uniform subset, negatives
 * uniform over the OTHER masked steps of the same utterance), not the same numbers: the bit-exact path stays the host
 * one.  Effective seed = seed + *seed_dev (seed_dev: device uint64 or NULL), so a CUDA-graph replay draws afresh.
 * a8_span_mask_draw: rows int32 [R_max + 1] = flat indices b*T + t of the masked frames in row-major order, -1
 *   padding, rows[R_max] = their number; mask uint8 [B,T].  R_max >= B * min(T, int(p_start*T/mask_length + 1) *
 *   mask_length).  Fails (like np.random.choice in the reference) when the spans cannot start at distinct frames.
 * a8_negatives_draw: out int32 [R_max, K] candidate rows of the flattened latents for every masked step r <
 *   *n_valid (n_valid = rows + R_max), never r itself, always inside r's utterance (steps per utterance = *n_valid / B);
 *   0 for the padding rows.
 * ---------------------------------------------------------------------------------------------- */
int a8_span_mask_draw(uint64_t seed, const void* seed_dev, int32_t B, int32_t T, double p_start, int32_t mask_length,
                      int32_t R_max, int32_t* rows, uint8_t* mask, void* stream);
int a8_negatives_draw(uint64_t seed, const void* seed_dev, const int32_t* n_valid, int32_t B, int32_t K, int32_t R_max,
                      int32_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Parameter re-layout, once per step: fp32 masters in PyTorch layouts -> bf16 GEMM operands (csrc/wprep.cu).
 * a8_cast_multi: table of n entries {const float* src; void* dst; int64 numel; int64 dst_is_f32} in DEVICE memory;
 *   one launch casts / copies them all (e.g. w_Q|w_K|w_V into one fused [3D,D] bf16 operand without a concat).
 * a8_conv_pack: Conv1d weight [Cout,Cin,k] (`wav2vec2.py:426`) -> wk [Cout,k*Cin] and, per stride phase p < 2,
 *   wt_p [Cin, ntaps_p*Cout] for the data-gradient GEMMs (wt0/wt1 may be NULL).  a8_conv_unpack: the inverse for dwk.
 * a8_posconv_pack: weight_norm(dim=2) (`wav2vec2.py:609`): w = g[j]*v/||v[:,:,j]|| -> packed bf16 [D,k*64] (rows =
 *   output channels) and its per-group transpose (rows = input channels); norm2[k] is kept for backward and must be
 *   followed by a8_posconv_norm_scratch_floats(D, cg, k) floats of scratch (ordered partial sums: the result is
 *   bit-reproducible, no atomics).
 * a8_posconv_wn_bwd: from the GEMM's dwp fp32 [groups,k*64,64] to dv [D,cg,k], dg [k]; t_scratch: k floats.
 * ---------------------------------------------------------------------------------------------- */
int a8_cast_multi(const void* table, int32_t n_entries, void* stream);
int a8_conv_pack(const float* w, int32_t Cout, int32_t Cin, int32_t k, int32_t s, void* wk, void* wt0, void* wt1,
                 void* stream);
int a8_conv_unpack(const float* dwk, int32_t Cout, int32_t Cin, int32_t k, float* dw, void* stream);
int64_t a8_posconv_norm_scratch_floats(int32_t D, int32_t cg, int32_t k);
int a8_posconv_pack(const float* g, const float* v, int32_t D, int32_t cg, int32_t k, float* norm2, void* wp,
                    void* wpt, void* stream);
int a8_posconv_wn_bwd(const float* dwp, const float* g, const float* v, const float* norm2, int32_t D, int32_t cg,
                      int32_t k, float* t_scratch, float* dv, float* dg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer side of the step (csrc/optim.cu): replaces `torch.nn.utils.clip_grad_norm_(model.parameters(),
 * args.clip)` + `optimizer.step()` (torch.optim.AdamW behind eight_mile's OptimizerManager) + `optimizer.scale_grads(s)`
 * at /root/reference/audio8/pretrain.py:182-184 and train.py:323-325 with two multi-tensor launches.
 * table: n_tensors rows of 6 int64 in DEVICE memory {float* p, const float* g (0 = no gradient this step: skipped),
 *   float* exp_avg, float* exp_avg_sq, int64 numel, bf16* operand copy of p or 0}.  The work is cut into chunks of
 *   `chunk` elements (multiple of 4): chunk_tensor[c] = table row, chunk_off[c] = element offset inside that tensor.
 * a8_optim_grad_sqnorm: partials[c] = sum of g^2 over chunk c (no atomics, nothing to zero).
 * a8_optim_adamw: total = sqrt(sum partials) * |grad_scale|; coef = min(1, max_norm / (total + 1e-6)) when max_norm > 0;
 *   every gradient is read as g * grad_scale * coef (never written back) and the parameter, exp_avg, exp_avg_sq (and the
 *   bf16 copy) are updated with torch.optim.AdamW's arithmetic (decoupled weight decay, bias_correction1 = 1 - beta1^t,
 *   bias_correction2_sqrt = sqrt(1 - beta2^t), both computed by the caller; the hyper-parameters arrive as doubles and
 *   1 - lr*wd, 1 - beta1, 1 - beta2, lr / bias_correction1 are formed in double and rounded once, as torch does).  partials may be NULL when max_norm <= 0.
 *   scale_grads_only != 0: clip_grad_norm_ semantics instead - the gradients are scaled in place, nothing else changes.
 *   total_norm_out (optional): the pre-clip gradient norm, what clip_grad_norm_ returns.
 * ---------------------------------------------------------------------------------------------- */
int a8_optim_grad_sqnorm(const void* table, const int32_t* chunk_tensor, const int64_t* chunk_off, int32_t n_chunks,
                         int32_t chunk, float* partials, void* stream);
int a8_optim_adamw(const void* table, const int32_t* chunk_tensor, const int64_t* chunk_off, int32_t n_chunks,
                   int32_t chunk, const float* partials, float max_norm, float grad_scale, double lr, double beta1,
                   double beta2, double eps, double weight_decay, double bias_correction1, double bias_correction2_sqrt,
                   int32_t scale_grads_only, float* total_norm_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange through the NVSwitch (replaces DistributedDataParallel's bucketed NCCL all-reduce,
 * /root/reference/audio8/pretrain.py:153, train.py:285): in-place all-reduce of fp32 elements [begin, end) of a buffer
 * that every rank has mapped into one MULTICAST window (multicast_base = this rank's multicast address of element 0;
 * audio8_b200/parallel.py gets it from torch's symmetric memory).  Rank r reduces its 1/world slice with
 * multimem.ld_reduce, multiplies by `scale` (1/world: DDP's average) and multimem.st's it to every rank.
 * begin / end are multiples of 4 elements.  ctas <= 0 picks the default.  The caller orders the launch between two
 * cross-rank barriers on the same stream.  Fails (-1) without a multicast mapping.
 * ---------------------------------------------------------------------------------------------- */
int a8_allreduce_mc(void* multicast_base, int64_t begin, int64_t end, int32_t rank, int32_t world, float scale,
                    int32_t ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIO8_B200_H */
