/* mhimk.h -- C ABI of libmhimk.so: B200 (sm_100a) kernels for MHIM-MIL's per-bag aggregation path.
 *
 * The reference (DearCaat/MHIM-MIL @ 9d0c91a) is pure PyTorch and has NO FFI for this path; each entry
 * point below names the reference code it replaces (file:line under /root/reference).  INTEGRATION.md
 * shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns all memory incl. workspaces; kernels never allocate; outputs are fully overwritten;
 *   - all work is enqueued on `stream` (a cudaStream_t); no call synchronises the device;
 *   - return 0 on success, <0 for an argument error, >0 = cudaError_t; mil_last_error() describes the
 *     last failure on the calling thread;
 *   - batch is always one bag (as everywhere in the reference); fp32 in / fp32 out; indices int64;
 *   - calls that share a workspace (`ws`: weight images + the fused pass's finalisation counter) must be ordered on ONE stream; calls with
 *     different workspaces are independent.
 */
#ifndef MHIMK_H_
#define MHIMK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIL_ABI_VERSION 2

/* activation codes used by every entry point */
enum { MIL_ACT_NONE = 0, MIL_ACT_RELU = 1, MIL_ACT_GELU = 2, MIL_ACT_TANH = 3, MIL_ACT_SIGMOID = 4 };

/* arithmetic of the tensor-core contractions in the fused pass */
enum {
  MIL_PREC_BF16X3 = 0, /* x = hi + lo (bf16), 3 tcgen05 products, fp32 accumulate: fp32-class (parity mode) */
  MIL_PREC_FP16   = 1, /* single fp16 product, fp32 accumulate: TF32-class (the reference's own GPU numerics,
                          main.py:435 enables TF32) */
  MIL_PREC_BF16   = 2, /* single bf16 product (fastest, ~2^-9 operand rounding) */
  MIL_PREC_FP16X3 = 3  /* x = hi + lo (fp16), 3 tcgen05 products, fp32 accumulate: unit roundoff ~2^-22 (bf16x3: 2^-17) at the same
                          cost -- ReLU gates flip ~30x less often against an fp32 reference.  Operands must stay inside fp16's range:
                          |x| <= 65504 (larger values saturate); weight images are built from 16 W (exact), undone in the epilogue. */
};

/* pipeline of the fused pass, OR-ed into `precision` as (pipeline << 8); 0 = library default (the pair pipeline) */
enum {
  MIL_PIPE_DEFAULT = 0,
  MIL_PIPE_SINGLE  = 1, /* one CTA per 128-row tile (cta_group::1, M = 128): the accumulator fills TMEM, the epilogue of a tile
                           runs after its GEMM1 */
  MIL_PIPE_PAIR    = 2  /* two CTAs of a cluster share every MMA (cta_group::2, M = 128 -> 64 rows per CTA): the accumulator is
                           256 TMEM columns, double-buffered, so the epilogue of tile t overlaps GEMM1 of tile t+1 */
};

typedef void* mil_stream_t; /* cudaStream_t */

/* Dropout on the embedding h inside the kernels (nn.Dropout after the feature layer: modules/mhim.py:193-194,331; abmil.py:188-189).
 * A HOST struct passed by pointer (NULL or mode 0 = no dropout).  Kept values are scaled by 1/(1-p) like torch.nn.functional.dropout.
 *   mode 1: keep_bits = device uint32[rows][ncols/32], bit i of word (r, c) = keep flag of column 32c+i (parity tests feed the
 *           reference's own Bernoulli mask this way);
 *   mode 2: in-kernel Philox4x32-10: counter = (row, column/8, offset_lo, offset_hi), key = (seed_lo, seed_hi); each call yields
 *           eight 16-bit uniforms u (low half-word first), keep iff u < round((1-p)*65536).  Stateless in (row, column): the same
 *           (seed, offset) regenerates the same mask in the backward pass or on the host (tests/philox_ref.py).
 *   mode 3: as mode 2, but (seed, offset) are read by the kernel from DEVICE memory: keep_bits points to uint64[2] = {seed, offset}.
 *           For CUDA-graph capture: the host values of mode 2 would be baked into the graph and every replay would repeat one mask;
 *           here a captured generator kernel refreshes the two words before every replay (mhimk.engines.GraphedStep). */
enum { MIL_DROP_NONE = 0, MIL_DROP_BITS = 1, MIL_DROP_PHILOX = 2, MIL_DROP_PHILOX_DEV = 3 };
typedef struct {
  int             mode;
  float           p;
  uint64_t        seed, offset;
  const uint32_t* keep_bits;
} mil_dropout_t;
/* keep_bits_out[rows][ncols/32] = the mask mode 2 would apply (ncols % 32 == 0). */
int mil_dropout_bits(int64_t rows, int ncols, const mil_dropout_t* drop, uint32_t* keep_bits_out, mil_stream_t stream);

int         mil_abi_version(void);
const char* mil_last_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM/TMA available), else 0. */
int         mil_device_supported(void);

/* ---------------------------------------------------------------------------------------------
 * Fused ABMIL forward pass: ONE streaming pass over the bag.
 *   h_n = drop(act(W1 x_n + b1));  s_n = wc . att_act(Wa h_n + ba) + bc   (att_act: tanh, or relu / gelu for MHIM's da_act)
 *   per-CTA online softmax over its rows -> partial (m, l, P[H]) -> merged (m, l, pooled = P/l)
 * Replaces: modules/abmil.py:213-234 (DAttention.forward), :121-139 (AttentionGated.forward),
 *           modules/mhim.py:193 + modules/mhim_modules/baseline.py:31-41,97-110 (feature + DAttention),
 *           and with `keep` the gather of modules/mhim_modules/masking.py:91-110.
 * X [N,D] row-major, 16-byte aligned, D % 32 == 0, H == 512, Da == 128; Wb / bb (the gated branch) must be NULL: the gated heads
 * (abmil.py:111-143, Da = 384) are NOT fused -- their two 128 x 384 GEMM2 accumulators do not fit the 512 TMEM columns next to h
 *  -- and run as mil_linear_act_tc_f32 + mil_softmax_pool_fwd_f32 launches; passing Wb returns an argument error.
 * keep      nullable uint8[N]: rows with keep[n]==0 are skipped (masked instances).
 * s_out     nullable float[N]: raw attention logits (-inf for skipped rows).
 * t_out     nullable float[N,C]: t_{n,c} = h_n . Wp_c (needs Wp [C,H], C <= 4) -- input of mil_cam_score_f32.
 * h_out     nullable float[N,H]: materialised embedding (training / return_act).
 * drop      nullable: dropout applied to h right after the activation (everything downstream -- attention logits, pooling,
 *           t_out, h_out -- sees the dropped embedding, as in the reference).
 * part      float[n_part,(2+H)] scratch for the per-CTA partials, n_part = mil_fused_num_partials() (one record more than
 *           CTAs are launched: the bulk copies of the in-kernel merge round up to 16 bytes).
 * stats     float[2] = (m, l); pooled float[H]: written by the last CTA to finish (in-kernel log-sum-exp merge of the partials).
 * rec_out   nullable float[2+H] = (m, l, P[H] = l * pooled): the record an instance-sharded bag exchanges (one all-gather, then
 *           mil_shard_merge_cls_f32), written by the same tail -- no extra launch.
 * logits    nullable float[n_cls] = Wcls pooled + bcls (classifier fused into the same tail; replaces abmil.py:238 /
 *           mhim.py:267); Wcls [n_cls, H], bcls nullable.
 * ws / ws_bytes: scratch of at least mil_fused_workspace_bytes(D, H, Da, gated) bytes; it holds the 16-bit hi/lo images of
 *           W1 and Wa, the finalisation counter and (pair pipeline) the exchange area of the tail split: when the last wave of
 *           128-row tiles fills at most half of the CTA pairs, each of its tiles is shared by K range between two pairs, one of
 *           which hands its partial accumulator (128 x 512 fp32) to the other through this area (~19 MB on a 148-SM part;
 *           MHIMK_NOSPLIT=1 switches the split off).  Deterministic: one partial, added in a fixed place.
 *           ws_ready = 0: the images are (re)built by this call; ws_ready = 1: the caller guarantees `ws` was
 *           filled by an earlier call with the same weights, precision AND pipeline (skips two small kernels per bag).
 * precision MIL_PREC_* | (MIL_PIPE_* << 8).
 */
int mil_abmil_fused_fwd_f32(const float* X, int64_t N, int D, int H,
                            const float* W1, const float* b1, int act,
                            const float* Wa, const float* ba, const float* Wb, const float* bb, int Da, int att_act,
                            const float* wc, const float* bc,
                            const uint8_t* keep, const float* Wp, int C,
                            float* s_out, float* t_out, float* h_out,
                            float* part, float* stats, float* pooled, float* rec_out,
                            const float* Wcls, const float* bcls, int n_cls, float* logits, const mil_dropout_t* drop,
                            void* ws, size_t ws_bytes, int ws_ready, int precision, mil_stream_t stream);
int    mil_fused_num_partials(void);
/* Kernel-only timing of the fused pass for the roofline line of bench.py: while enabled, every fused launch is bracketed
 * by CUDA events on its stream; mil_profile_collect() synchronises them, returns how many launches were timed and writes
 * their summed duration in ms. */
void   mil_profile_enable(int on);
int    mil_profile_collect(double* total_ms);
size_t mil_fused_workspace_bytes(int D, int H, int Da, int gated);

/* ---------------------------------------------------------------------------------------------
 * C[M,N] = act( sum_k A(m,k) B(n,k) + bias[n] ), fp32 FFMA tiles (exact fp32; the parity reference on the GPU
 * and the workhorse of the backward pass).  A(m,k) = A[rowA(m)*sAm + k*sAk] with rowA(m) = row_ids ? row_ids[m] : m,
 * B(n,k) = B[n*sBn + k*sBk].  One of (sAm,sAk) and one of (sBn,sBk) must be 1.  pre_out (nullable, ld = ldc)
 * receives the pre-activation.  splitk > 1 splits K over `splitk` slices reduced deterministically through ws
 * (ws_bytes >= splitk*M*N*4); use for weight gradients where K = #instances.
 * Replaces: every nn.Linear on the path (abmil.py:181,194-196; mhim.py:69,97; baseline.py:15,27; dsmil.py:62-70,133)
 * and its autograd (weight grad = TN form, input grad = NN form).
 */
int mil_sgemm_f32(const float* A, int64_t sAm, int64_t sAk, const int64_t* row_ids,
                  const float* B, int64_t sBn, int64_t sBk, const float* bias,
                  float* C, int64_t ldc, float* pre_out,
                  int64_t M, int64_t N, int64_t K, int act, int splitk, void* ws, size_t ws_bytes, mil_stream_t stream);

/* Batched fp32 GEMM: C_b[M,N] = sum_k A_b(m,k) B_b(n,k) for b < batch, operand b at base + b * bA / bB / bC elements (strides as in
 * mil_sgemm_f32, no bias / activation / split-K).  The 8 x (256 x 256 x 256) products of the Moore-Penrose iteration
 * (nystrom_attention.py:12-27) and the small landmark products of the Nystrom layers. */
int mil_sgemm_batched_f32(const float* A, int64_t sAm, int64_t sAk, int64_t bA, const float* B, int64_t sBn, int64_t sBk, int64_t bB,
                          float* C, int64_t ldc, int64_t bC, int64_t M, int64_t N, int64_t K, int batch, mil_stream_t stream);

/* Tensor-core variant of the NT form: Y[M,N] = act(X[M,K] W[N,K]^T + bias) through the fused pass's TMA -> 16-bit split ->
 * tcgen05 pipeline (fp32-class results in MIL_PREC_BF16X3).  K % 32 == 0; N in {64,128,192,256,512}; pre_out nullable
 * (pre-activation, ld = N).  ws >= mil_linear_tc_workspace_bytes(N, K) holds the weight image; ws_ready as above.
 * Replaces the forward of the N-row projections (mhim.py:69, dsmil.py:62-70, nystrom_attention.py:52-57, merge.py:35-41). */
int    mil_linear_act_tc_f32(const float* X, int64_t M, int K, const float* W, const float* bias, int N, int act, float* pre_out,
                             float* Y, const mil_dropout_t* drop /* nullable: dropout on Y (N % 32 == 0) */,
                             void* ws, size_t ws_bytes, int ws_ready, int precision, mil_stream_t stream);
/* The same on operands that are column blocks of wider row-major buffers: X has leading dimension ldx (>= K), Y and pre_out have ldy (>= N);
 * both multiples of 4.  (Per-head slices of the qkv buffer, head-wise outputs of the Nystrom aggregation.) */
int    mil_linear_act_tc_ld_f32(const float* X, int64_t ldx, int64_t M, int K, const float* W, const float* bias, int N, int act, float* pre_out,
                                float* Y, int64_t ldy, const mil_dropout_t* drop, void* ws, size_t ws_bytes, int ws_ready, int precision,
                                mil_stream_t stream);
size_t mil_linear_tc_workspace_bytes(int N, int K);

/* Skinny Linear layers (GEMV-shaped), forward and backward, exact fp32 streaming kernels: Y[M,N] = act(X[M,K] W[N,K]^T + b) with
 * N <= 8 outputs over many rows (attention logit 128 -> 1: abmil.py:196, baseline.py:27; DSMIL instance classifier / critical-instance
 * logits: dsmil.py:62,93) or M <= 8 rows (classifier / predictor on the pooled vector: abmil.py:238, mhim.py:267; Merge's to_q / to_out:
 * merge.py:35-41; q(h_crit): dsmil.py:92).  K <= 1536.  mil_skinny_supported: 0 = no, 1 = "thin" (N <= 8), 2 = "short" (M <= 8).
 * X has leading dimension ldx; pre_out nullable (pre-activation, for the gelu backward).
 * Backward: G = dL/d(pre-activation) [M,N]; any of dW [N,K], db [N] (only together with dW), dX [M,K] may be NULL.
 * ws >= mil_skinny_workspace_bytes(M, N, K) (row-slice partials of the thin weight gradient, reduced in a fixed order). */
int    mil_skinny_supported(int64_t M, int N, int K);
int    mil_skinny_fwd_f32(const float* X, int64_t ldx, int64_t M, int K, const float* W, const float* b, int N, int act,
                          float* pre_out, float* Y, mil_stream_t stream);
int    mil_skinny_bwd_f32(const float* G, const float* X, int64_t ldx, const float* W, int64_t M, int N, int K,
                          float* dW, float* db, float* dX, void* ws, size_t ws_bytes, mil_stream_t stream);
size_t mil_skinny_workspace_bytes(int64_t M, int N, int K);

/* Weight (and bias) gradient of an N-row Linear on the tensor cores: dW[n,k] = sum_m G[m,n] X[m,k], db[n] = sum_m G[m,n]
 * (db nullable).  G [M, Nn] (leading dimension ldg), X [M, K] (ldx), fp32 row-major, 16-byte aligned; Nn % 128 == 0, K % 256 == 0.
 * bf16 hi+lo split, 3 tcgen05 products, fp32 accumulation (fp32-class); the instance range is cut into slices whose partial
 * sums are added in a fixed order (deterministic).  ws >= mil_wgrad_tc_workspace_bytes(M, Nn, K).
 * Replaces the autograd weight gradient of every N-row nn.Linear on the path (abmil.py:213, mhim.py:193/335, baseline.py:15,27,
 * dsmil.py:62-70, nystrom_attention.py:52-57, merge.py:35-41) -- the dW1 = sum_n g_pre^T x reduction of SURVEY 9.2. */
int    mil_wgrad_tc_f32(const float* G, int64_t ldg, const float* X, int64_t ldx, int64_t M, int Nn, int K, float* dW, float* db,
                        void* ws, size_t ws_bytes, mil_stream_t stream);
size_t mil_wgrad_tc_workspace_bytes(int64_t M, int Nn, int K);

/* g_pre = g_y * act'(.) elementwise; `y_or_pre` is the activation OUTPUT for relu/tanh/sigmoid and the
 * PRE-activation for gelu.  n elements.  (autograd of nn.ReLU/GELU/Tanh/Sigmoid on the path) */
int mil_act_bwd_f32(const float* g_y, const float* y_or_pre, int64_t n, int act, float* g_pre, mil_stream_t stream);
/* The same through a dropout that followed the activation: g_pre = g_y * keep/(1-p) * act'(.) over a [rows, ncols] tensor (ncols % 32 == 0);
 * the mask is regenerated from `drop` (mode 1 or 2).  `y_or_pre`: for relu the activation output (dropped or not: y_dropped > 0 <=> kept
 * and pre > 0); for gelu / tanh / sigmoid the PRE-activation. */
int mil_act_bwd_drop_f32(const float* g_y, const float* y_or_pre, int64_t rows, int ncols, int act, const mil_dropout_t* drop, float* g_pre,
                         mil_stream_t stream);

/* out[n] = sum_m A[m,n]  (bias gradients), deterministic. */
int mil_colsum_f32(const float* A, int64_t M, int64_t N, float* out, void* ws, size_t ws_bytes, mil_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * softmax over instances + weighted sum:  a = softmax_L(s); pooled = a @ h.
 * Replaces: abmil.py:231-234, baseline.py:33-36, and (per class column) dsmil.py:94-96 / baseline.py:144-147.
 * s [L] (stride s_stride floats), h [L,H]; keep nullable uint8[L]; part: float[n_part,(2+H)] with
 * n_part = mil_pool_num_partials(L); stats float[2] = (m, l); pooled float[H]; attn_out nullable float[L].
 */
int mil_softmax_pool_fwd_f32(const float* s, int64_t s_stride, const float* h, int64_t L, int H, const uint8_t* keep,
                             float* part, float* stats, float* pooled, float* attn_out, mil_stream_t stream);
int mil_pool_num_partials(int64_t L);
/* backward: g_s[n] = a_n (h_n.g_p - pooled.g_p) (+ g_attn handling is done by the caller);
 * g_h[n,:] (+)= a_n g_p   (g_h nullable; accumulate_gh != 0 adds into g_h). */
int mil_softmax_pool_bwd_f32(const float* s, int64_t s_stride, const float* h, int64_t L, int H, const uint8_t* keep,
                             const float* stats, const float* pooled, const float* g_p,
                             float* g_s, int64_t gs_stride, float* g_h, int accumulate_gh, mil_stream_t stream);

/* Merge n_part partials (m_i, l_i, P_i[H]) laid out as float[n_part,(2+H)] into stats=(m,l), pooled=P/l.
 * Used for the per-CTA partials of one GPU and for the per-rank partials of an instance-sharded bag
 * (SURVEY.md §9.3); entries with l_i == 0 are ignored. */
int mil_pool_merge_f32(const float* part, int n_part, int H, float* stats, float* pooled, mil_stream_t stream);

/* Instance-sharded bag (BASELINE config 5): merge the n_rec gathered records (m_g, l_g, P_g[H]) of the ranks AND apply the classifier
 * in ONE launch: stats = (m, l), pooled = sum_g P_g e^{m_g - m} / l, logits = Wcls pooled + bcls (nullable Wcls -> no logits).
 * Records with l_g == 0 (ranks without rows) are ignored; n_rec <= 1024; fixed summation order -> identical on every rank. */
int mil_shard_merge_cls_f32(const float* rec, int n_rec, int H, const float* Wcls, const float* bcls, int n_cls,
                            float* stats, float* pooled, float* logits, mil_stream_t stream);

/* Teacher attention -> score.  Replaces modules/mhim_modules/scoring.py:37-58 (get_pseudo_score):
 * score_n = max_c softmax_c( a_n * t_{n,c} + bias0 ), a_n = exp(s_n - m)/l, t = h Wp^T given as float[L,C]. */
int mil_cam_score_f32(const float* s, const float* t, int64_t L, int C, const float* stats, float bias0,
                      float* score, mil_stream_t stream);
/* Same, with bias0 = bias_dev[0] read on the device (the predictor's bias parameter, scoring.py:54 `classifier bias[0]`):
 * the caller needs no device->host read of the parameter, so the teacher pass enqueues without a host sync. */
int mil_cam_score_dev_f32(const float* s, const float* t, int64_t L, int C, const float* stats, const float* bias_dev,
                          float* score, mil_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Masked hard-instance selection.  Replaces torch.topk + the python-set complement of
 * modules/mhim_modules/masking.py:61-63,77-86 (select_mask_fn).
 * mil_topk_f32: idx_out[k] = indices of the k largest (largest!=0) / smallest scores, ordered by value
 *   (descending for largest) with ties broken lowest-index-first (a total order: deterministic).
 * mil_mask_from_indices: mask_ids[N] = [ids NOT in idx, ascending || idx in the given order], keep[N] = 1 for
 *   kept rows, 0 for masked; *len_keep_out (device int64) = N - #distinct idx.  idx entries must be distinct.
 * ws: at least mil_topk_workspace_bytes(N) bytes.
 */
int    mil_topk_f32(const float* score, int64_t N, int64_t k, int largest, int64_t* idx_out,
                    void* ws, size_t ws_bytes, mil_stream_t stream);
int    mil_mask_from_indices(const int64_t* idx, int64_t k, int64_t N, int64_t* mask_ids, uint8_t* keep,
                             int64_t* len_keep_out, void* ws, size_t ws_bytes, mil_stream_t stream);
size_t mil_topk_workspace_bytes(int64_t N);
/* Critical instance per class: idx_out[c] = argmax_m A[m,c] over the M instances (lowest index among equal maxima), val_out[c]
 * (nullable) = that maximum.  Replaces the full `torch.sort(c, 0, descending=True)` the reference runs only to read row 0
 * (modules/dsmil.py:91-92, mhim_modules/baseline.py:137-138) and the max-pooling of the instance logits (dsmil.py:160). */
int    mil_col_argmax_f32(const float* A, int64_t M, int C, int64_t* idx_out, float* val_out, mil_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row selection by a permutation, forward and backward.  Replaces the gathers of modules/mhim_modules/masking.py:91-110 (mask_fn:
 * x[mask_ids[:len_keep]]) and modules/mhim_modules/merge.py:171-174 (keep / drop split by argsort(rand(L))) and their autograd
 * (torch: index_put_ with accumulate, which sorts the indices).  `perm` is a permutation of the rows of x (or a prefix of one).
 *   take:    out[i,:] = x[perm[i],:], i < n_out.
 *   scatter: gx[perm[i],:] = i < n_a ? ga[i,:] : (gb ? gb[i-n_a,:] : 0), i < n_rows -- every row of gx written exactly once.
 * cols % 4 == 0, 16-byte aligned. */
int mil_take_rows_f32(const float* x, const int64_t* perm, int64_t n_out, int cols, float* out, mil_stream_t stream);
int mil_scatter_rows_f32(const float* ga, const float* gb, const int64_t* perm, int64_t n_a, int64_t n_rows, int cols, float* gx,
                         mil_stream_t stream);

/* Merge's cross-attention (modules/mhim_modules/merge.py:52-65): kq <= 8 query tokens over L instances, `heads` x dh (dh <= 64).
 * q [kq, heads*dh]; kv [L, 2*heads*dh] = [K | V] (the to_kv output).  P [heads, kq, L] receives softmax_L(scale q.K) (kept for the
 * backward); pmask (nullable, same shape) is the attention dropout mask already scaled by 1/(1-p) (merge.py:60);
 * out [kq, heads*dh] = (P * pmask) V.  Backward: dq [kq, heads*dh], dkv [L, 2*heads*dh]; dS_scratch [heads, kq, L].
 * ws >= mil_mca_workspace_bytes(L, kq, heads, dh).  Deterministic (chunk partials summed in a fixed order). */
int    mil_mca_fwd_f32(const float* q, const float* kv, int64_t L, int kq, int heads, int dh, float scale, const float* pmask, float* P,
                       float* out, void* ws, size_t ws_bytes, mil_stream_t stream);
int    mil_mca_bwd_f32(const float* g_out, const float* q, const float* kv, const float* P, const float* pmask, int64_t L, int kq, int heads,
                       int dh, float scale, float* dS_scratch, float* dq, float* dkv, void* ws, size_t ws_bytes, mil_stream_t stream);
size_t mil_mca_workspace_bytes(int64_t L, int kq, int heads, int dh);

/* ---------------------------------------------------------------------------------------------
 * EMA teacher update as ONE launch over all parameters.  Replaces the per-parameter loop
 * `param_k.data.mul_(mm).add_(param_q.data, alpha=1 - mm)` of engines/base_engine.py:166-167 (and :488-489).
 * segs_dev: device array of n_seg records; the caller cuts every parameter into segments of a few 10^4 elements (one CTA
 * each).  dst <- fma(one_minus_mm, src, dst * mm), i.e. the same two roundings as the reference's mul_ + add_(alpha).
 * mm outside [0, 1] is an argument error (the reference asserts the same, base_engine.py:164). */
typedef struct { float* dst; const float* src; int64_t n; } mil_ema_seg_t;
int mil_ema_update_f32(const mil_ema_seg_t* segs_dev, int n_seg, float mm, float one_minus_mm, mil_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * CUDA-core pieces of the Nystrom / TransMIL path (modules/nystrom_attention.py:65-152, transmil.py:23-64, emb_position.py:85-120).
 *   mil_layernorm_fwd_f32     y = LayerNorm(x) over the last axis (transmil.py:31, baseline.py:205).
 *   mil_segment_mean_f32      landmarks: out[h][j][d] = scale * mean over segment j (seg_len rows) of x[row, col0 + h*dh + d]
 *                             (nystrom_attention.py:93-109; x = the qkv buffer, leading dimension ld).
 *   mil_row_softmax_f32       in-place softmax over the last axis of S [rows, cols <= 1024] (attn1, attn2: :127).
 *   mil_colsoftmax_pool_f32   out[j,:] = sum_n softmax_n(S[n,j]) V[n,:] for S [n, m <= 256] = (k q_l^T), V [n, dh <= 64] (leading dimension ldv):
 *                             attn3 @ v without materialising attn3 (SURVEY 9.7 pass A); also returns the column maxima / sums.
 *   mil_expdot_rows_f32       out[r] = sum_j w[j] exp(S[r,j] - M[j])  -- the cls-row attention over the keys (:143-150).
 *   mil_dwconv_tokens_f32     out[r, c] (+)= sum_t w[head(c)][t] v[r + t - taps/2, c]: the depth-wise residual conv over tokens (:135-136).
 *   mil_ppeg_f32              y = depth-wise 7x7 conv (zero padded) of the [H*W, C] token grid with an effective kernel w49 [49,C] (tap-major; = 7x7 + padded
 *                             5x5 + padded 3x3 + identity) + bias (transmil.py:50-64, emb_position.py:85-120). */
int    mil_layernorm_fwd_f32(const float* x, int64_t rows, int cols, const float* w, const float* b, float eps, float* y, mil_stream_t stream);
int    mil_segment_mean_f32(const float* x, int64_t ld, int m, int seg_len, int col0, int heads, int dh, float scale, float* out, mil_stream_t stream);
int    mil_row_softmax_f32(float* S, int64_t rows, int cols, mil_stream_t stream);
int    mil_colsoftmax_pool_f32(const float* S, const float* V, int64_t ldv, int64_t n, int m, int dh, float* out, float* colmax, float* colsum,
                               void* ws, size_t ws_bytes, mil_stream_t stream);
size_t mil_colsoftmax_pool_workspace_bytes(int64_t n, int m);
int    mil_expdot_rows_f32(const float* S, int64_t rows, int m, const float* M, const float* w, float* out, mil_stream_t stream);
int    mil_dwconv_tokens_f32(const float* v, int64_t ldv, int64_t rows, int heads, int dh, const float* w, int taps, float* out, int64_t ldo,
                             int accumulate, mil_stream_t stream);
int    mil_ppeg_f32(const float* x, int H, int W, int C, const float* w49, const float* bias, float* y, mil_stream_t stream);

/* Adam / AdamW optimiser step as ONE launch over all parameters (the reference builds torch.optim.Adam / AdamW,
 * train_utils.py:55-65, and calls optimizer.step() once per bag, engines/base_engine.py:110-120).  segs_dev: device array of records
 * (parameter, gradient, exp_avg, exp_avg_sq, n) cut into segments of a few 10^4 elements (one CTA each).  decoupled = 0: Adam (L2
 * weight decay added to the gradient), 1: AdamW.  bias_correction1 = 1 - beta1^t, bias_correction2_sqrt = sqrt(1 - beta2^t) for the
 * step count t of THIS step (betas are doubles so that 1 - beta is rounded once, as torch's python scalars are); step_dev (nullable) = device float holding t instead (capturable: CUDA-graph replays). */
typedef struct { float* p; const float* g; float* m; float* v; int64_t n; } mil_adam_seg_t;
int mil_adam_step_f32(const mil_adam_seg_t* segs_dev, int n_seg, float lr, double beta1, double beta2, float eps, float weight_decay,
                      int decoupled, float bias_correction1, float bias_correction2_sqrt, const float* step_dev, mil_stream_t stream);

/* Self-test hook for the tcgen05/TMA plumbing: C[M,N] = A[M,K] B[N,K]^T with the fused pass's operand pipeline
 * (fp32 in HBM -> TMA -> bf16/fp16 split in shared memory -> tcgen05.mma -> TMEM -> registers).  M % 128 == 0,
 * N in {64,128,256,512}, K % 32 == 0.  Used by tests/ only. */
int mil_umma_selftest_f32(const float* A, const float* B, float* C, int M, int N, int K, int precision,
                          void* ws, size_t ws_bytes, mil_stream_t stream);

/* Test hook for the host-side plan of the fused pass's pair pipeline (no device work, no GPU needed): the i-th work item of CTA pair
 * `pair` for a bag of N rows x D features.  out9[0..4] = 128-row tile, first and end pipeline stage of its K loop, kind (0 whole tile,
 * 1 owner / 2 helper of a tail-split tile), index of the exchanged partial; out9[5..8] = number of CTA pairs, parts per split tile,
 * whole-tile waves, tiles of the split wave.  Returns the number of items of that pair (out9[0..4] untouched when i is out of range),
 * < 0 on a bad argument.  The SM count is the current device's (148 without one). */
int mil_pair_plan_item(int64_t N, int D, int precision, int pair, int i, int64_t* out9);

#ifdef __cplusplus
}
#endif
#endif /* MHIMK_H_ */
