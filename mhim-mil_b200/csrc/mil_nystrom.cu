// CUDA-core pieces of the Nystrom / TransMIL path (nystrom_attention.py:65-152, transmil.py:23-64, emb_position.py:85-120) that sit
// between the tensor-core projections: LayerNorm, landmark segment means, softmax over the landmark axis, the streaming
// softmax-over-N aggregation kv = softmax_N(q_l k^T) v of SURVEY 9.7 (the [m x N] similarity never needs a transpose or a second
// materialisation), the 33-tap depth-wise residual convolution over the token axis, the cls-row attention read-out, the PPEG
// depth-wise convolution and a batched fp32 GEMM for the 256 x 256 pseudo-inverse iteration.  All bandwidth- or FFMA-bound
// streaming kernels; the big contractions (qkv, q k_l^T, softmax . Z, to_out) run on tcgen05 (mil_linear_act_tc_ld_f32).
#include "mil_common.cuh"

namespace mil {
namespace nys {

// y[r, :] = (x[r, :] - mean) * rstd * w + b       one warp per row
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t rows, int cols, const float* __restrict__ w,
                                                        const float* __restrict__ b, float eps, float* __restrict__ y) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + r * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c];
  const float mean = warp_sum(s) / cols;
  float v = 0.f;
  for (int c = lane; c < cols; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(warp_sum(v) / cols + eps);
  for (int c = lane; c < cols; c += 32) y[r * cols + c] = (xr[c] - mean) * rstd * w[c] + (b ? b[c] : 0.f);
}

// out[h][j][d] = scale * mean over the rows of segment j of x[row, col0 + h * dh + d]      grid = m segments, block = heads * dh threads
__global__ void segment_mean_kernel(const float* __restrict__ x, int64_t ld, int seg_len, int col0, int heads, int dh, float scale, int m,
                                    float* __restrict__ out) {
  const int j = blockIdx.x, c = threadIdx.x;
  if (c >= heads * dh) return;
  const float* p = x + (int64_t)j * seg_len * ld + col0 + c;
  float s = 0.f;
  for (int r = 0; r < seg_len; ++r) s += p[(int64_t)r * ld];
  const int h = c / dh, d = c % dh;
  out[((int64_t)h * m + j) * dh + d] = s * (scale / seg_len);
}

// in-place softmax over the last axis of S [rows, cols] (cols <= 1024)      one warp per row
__global__ void __launch_bounds__(256) row_softmax_kernel(float* __restrict__ S, int64_t rows, int cols) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float* s = S + r * cols;
  float v[32];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < cols ? s[c] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v[i] = (lane + 32 * i) < cols ? expf(v[i] - m) : 0.f;
    l += v[i];
  }
  const float inv = 1.f / warp_sum(l);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < cols) s[c] = v[i] * inv;
  }
}

// ---- kv[j, :] = sum_n softmax_n(S[n, j]) v[n, :]  with S = (k q_l^T) given [n, m] (the transpose of sim3), m = blockDim = 256 ----
// pass 1: pmax[chunk][j] = max over the chunk's rows
__global__ void __launch_bounds__(256) colmax_kernel(const float* __restrict__ S, int64_t n, int m, int64_t rows_per_chunk, float* __restrict__ pmax) {
  const int j = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  int64_t r1 = r0 + rows_per_chunk;
  if (r1 > n) r1 = n;
  float mx = -INFINITY;
  if (j < m)
    for (int64_t r = r0; r < r1; ++r) mx = fmaxf(mx, S[r * m + j]);
  if (j < m) pmax[(int64_t)blockIdx.x * m + j] = mx;
}
// pass 2: per chunk, thread j: e = exp(S[r, j] - M[j]); pl[chunk][j] = sum e; pacc[chunk][j][d] = sum e * v[r, d]   (dh <= 64)
__global__ void __launch_bounds__(256, 2) colsoftmax_pool_kernel(const float* __restrict__ S, const float* __restrict__ V, int64_t ldv, int64_t n, int m, int dh,
                                                              int64_t rows_per_chunk, const float* __restrict__ pmax, int chunks,
                                                              float* __restrict__ colmax_out, float* __restrict__ pl, float* __restrict__ pacc) {
  __shared__ __align__(16) float sv[32][64];
  const int j = threadIdx.x;
  float M = -INFINITY;
  if (j < m)
    for (int c = 0; c < chunks; ++c) M = fmaxf(M, pmax[(int64_t)c * m + j]);
  if (blockIdx.x == 0 && j < m) colmax_out[j] = M;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  int64_t r1 = r0 + rows_per_chunk;
  if (r1 > n) r1 = n;
  float acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  float l = 0.f;
  for (int64_t rb = r0; rb < r1; rb += 32) {
    const int nr = (int)((r1 - rb) < 32 ? (r1 - rb) : 32);
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 64; i += 256) {
      const int rr = i >> 6, d = i & 63;
      sv[rr][d] = (rr < nr && d < dh) ? V[(rb + rr) * ldv + d] : 0.f;
    }
    __syncthreads();
    if (j < m) {
      // the 32 similarity values of this thread's column first (32 independent coalesced loads in flight), then the FMA block: with
      // the load inside the row loop the kernel sat at 25 % FMA-pipe utilisation waiting on one load per 64 FMAs (ncu, round 2)
#pragma unroll 1
      for (int hb = 0; hb < 32; hb += 16) {                     // two batches of 16 rows: 16 loads in flight, <= 128 registers (2 CTAs per SM)
        float sv_j[16];
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) sv_j[rr] = (hb + rr) < nr ? S[(rb + hb + rr) * m + j] : -INFINITY;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
          const float e = expf(sv_j[rr] - M);                   // rows past the chunk: exp(-inf) = 0
          l += e;
#pragma unroll
          for (int d = 0; d < 64; d += 4) {                     // broadcast LDS.128: one shared-memory instruction per four FMAs
            const float4 v4 = *reinterpret_cast<const float4*>(&sv[hb + rr][d]);
            acc[d] = fmaf(e, v4.x, acc[d]); acc[d + 1] = fmaf(e, v4.y, acc[d + 1]);
            acc[d + 2] = fmaf(e, v4.z, acc[d + 2]); acc[d + 3] = fmaf(e, v4.w, acc[d + 3]);
          }
        }
      }
    }
  }
  if (j < m) {
    pl[(int64_t)blockIdx.x * m + j] = l;
    float* o = pacc + ((int64_t)blockIdx.x * m + j) * 64;
#pragma unroll
    for (int d = 0; d < 64; d += 4) *reinterpret_cast<float4*>(o + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
  }
}
// pass 3: out[j][d] = sum_chunks pacc / L[j];  L[j] = sum_chunks pl
__global__ void colsoftmax_finish_kernel(const float* __restrict__ pl, const float* __restrict__ pacc, int chunks, int m, int dh, float* __restrict__ colsum_out,
                                         float* __restrict__ out) {
  const int j = blockIdx.x, d = threadIdx.x;
  float L = 0.f;
  for (int c = 0; c < chunks; ++c) L += pl[(int64_t)c * m + j];
  if (d == 0) colsum_out[j] = L;
  if (d < dh) {
    float a = 0.f;
    for (int c = 0; c < chunks; ++c) a += pacc[((int64_t)c * m + j) * 64 + d];
    out[(int64_t)j * dh + d] = a / L;
  }
}

// out[r] = sum_j w[j] * exp(S[r, j] - M[j])      (cls-row attention over the keys: w = r / L, nystrom_attention.py:143-150)    one warp per row
__global__ void __launch_bounds__(256) expdot_rows_kernel(const float* __restrict__ S, int64_t rows, int m, const float* __restrict__ M,
                                                          const float* __restrict__ w, float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float a = 0.f;
  for (int j = lane; j < m; j += 32) a = fmaf(w[j], expf(S[r * m + j] - M[j]), a);
  a = warp_sum(a);
  if (lane == 0) out[r] = a;
}

// out[r, h*dh + d] (+)= sum_t w[h][t] * v[r + t - taps/2, h*dh + d]  (zero padding).  One CTA = 64 rows x 64 columns: the rows and their
// halo (taps - 1 <= 32 extra rows) are staged in shared memory once and reused by every tap (the first version re-read v from L2 per tap:
// 33 x 300 MB per call at N = 50 000).  256 threads = 64 columns x 4 row groups of 16 rows.
constexpr int DW_ROWS = 64, DW_COLS = 64, DW_MAXTAPS = 33;      // the reference's residual_conv_kernel (nystrom_attention.py:39)
__global__ void __launch_bounds__(256) dwconv_tokens_kernel(const float* __restrict__ v, int64_t ldv, int64_t rows, int heads, int dh, const float* __restrict__ w,
                                                            int taps, float* __restrict__ out, int64_t ldo, int accumulate) {
  __shared__ float tile[DW_ROWS + DW_MAXTAPS - 1][DW_COLS];
  __shared__ float sw[DW_COLS][DW_MAXTAPS];
  const int C = heads * dh, half = taps / 2;
  const int c0 = blockIdx.y * DW_COLS;
  const int64_t r0 = (int64_t)blockIdx.x * DW_ROWS;
  for (int i = threadIdx.x; i < (DW_ROWS + taps - 1) * DW_COLS; i += 256) {
    const int rr = i / DW_COLS, cc = i % DW_COLS;
    const int64_t r = r0 + rr - half;
    tile[rr][cc] = (r >= 0 && r < rows && c0 + cc < C) ? v[r * ldv + c0 + cc] : 0.f;
  }
  for (int i = threadIdx.x; i < DW_COLS * taps; i += 256) {
    const int cc = i / taps, t = i % taps;
    sw[cc][t] = (c0 + cc < C) ? w[((c0 + cc) / dh) * taps + t] : 0.f;
  }
  __syncthreads();
  const int cc = threadIdx.x & 63, rg = threadIdx.x >> 6;
  if (c0 + cc >= C) return;
  for (int k = 0; k < 16; ++k) {
    const int rr = rg * 16 + k;
    const int64_t r = r0 + rr;
    if (r >= rows) break;
    float a = 0.f;
    for (int t = 0; t < taps; ++t) a = fmaf(sw[cc][t], tile[rr + t][cc], a);
    float* o = out + r * ldo + c0 + cc;
    *o = accumulate ? *o + a : a;
  }
}

// PPEG: y[(yy, xx), c] = sum_{dy, dx} W[dy, dx][c] * x[(yy + dy - 3, xx + dx - 3), c] + bias[c]   (7 x 7 effective depth-wise kernel, zero padding).
// One CTA = an 8 x 8 patch of positions x 32 channels: the 14 x 14 halo patch is staged in shared memory ([channel][position], conflict-free for
// a warp of 32 positions) and reused by all 49 taps -- the direct version re-read x from L2 once per tap (49 x 100 MB at N = 50 000: 1.5 ms).
constexpr int PP_T = 8, PP_H = PP_T + 6, PP_C = 32, PP_LD = PP_H * PP_H + 1;
__global__ void __launch_bounds__(256) ppeg_kernel(const float* __restrict__ x, int H, int W, int C, const float* __restrict__ w49,
                                                   const float* __restrict__ bias, float* __restrict__ y) {
  __shared__ float tile[PP_C][PP_LD];
  __shared__ float sw[49][PP_C];
  const int tiles_x = (W + PP_T - 1) / PP_T;
  const int ty0 = (blockIdx.x / tiles_x) * PP_T, tx0 = (blockIdx.x % tiles_x) * PP_T, c0 = blockIdx.y * PP_C;
  for (int i = threadIdx.x; i < PP_H * PP_H * PP_C; i += 256) {           // consecutive threads -> consecutive channels of one position
    const int c = i % PP_C, pos = i / PP_C, yy = ty0 + pos / PP_H - 3, xx = tx0 + pos % PP_H - 3;
    tile[c][pos] = (yy >= 0 && yy < H && xx >= 0 && xx < W && c0 + c < C) ? x[((int64_t)yy * W + xx) * C + c0 + c] : 0.f;
  }
  for (int i = threadIdx.x; i < 49 * PP_C; i += 256) sw[i / PP_C][i % PP_C] = (c0 + i % PP_C < C) ? w49[(int64_t)(i / PP_C) * C + c0 + i % PP_C] : 0.f;
  __syncthreads();
  const int p = threadIdx.x & 63, cg = threadIdx.x >> 6;                    // 64 positions x 4 groups of 8 channels
  const int py = p / PP_T, px = p % PP_T, yy = ty0 + py, xx = tx0 + px;
  if (yy >= H || xx >= W) return;
#pragma unroll 1
  for (int cc = 0; cc < 8; ++cc) {
    const int c = cg * 8 + cc;
    if (c0 + c >= C) break;
    float a = bias ? bias[c0 + c] : 0.f;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy)
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) a = fmaf(sw[dy * 7 + dx][c], tile[c][(py + dy) * PP_H + px + dx], a);
    y[((int64_t)yy * W + xx) * C + c0 + c] = a;
  }
}

static int pool_chunks(int64_t n) {
  // whole waves: 2 CTAs of 256 threads are resident per SM, so 2 x #SM chunks run as ONE wave (197 chunks on 148 SMs ran as 1.33)
  const int wave = 2 * num_sms();
  int64_t c = (n + 63) / 64;                               // at least 64 rows per chunk
  if (c >= wave) c = wave;
  else if (c > num_sms()) c = num_sms();
  return (int)(c < 1 ? 1 : c);
}

}  // namespace nys
}  // namespace mil

using namespace mil;

extern "C" int mil_layernorm_fwd_f32(const float* x, int64_t rows, int cols, const float* w, const float* b, float eps, float* y, mil_stream_t stream) {
  MIL_CHECK_ARG(x && w && y && rows >= 0 && cols > 0, "mil_layernorm_fwd_f32: bad arguments");
  if (rows == 0) return 0;
  nys::layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, w, b, eps, y);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_segment_mean_f32(const float* x, int64_t ld, int m, int seg_len, int col0, int heads, int dh, float scale, float* out, mil_stream_t stream) {
  MIL_CHECK_ARG(x && out && m > 0 && seg_len > 0 && heads * dh > 0 && heads * dh <= 1024, "mil_segment_mean_f32: bad arguments (heads * dh <= 1024)");
  nys::segment_mean_kernel<<<m, heads * dh, 0, (cudaStream_t)stream>>>(x, ld, seg_len, col0, heads, dh, scale, m, out);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_row_softmax_f32(float* S, int64_t rows, int cols, mil_stream_t stream) {
  MIL_CHECK_ARG(S && rows >= 0 && cols > 0 && cols <= 1024, "mil_row_softmax_f32: bad arguments (cols <= 1024)");
  if (rows == 0) return 0;
  nys::row_softmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, rows, cols);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mil_colsoftmax_pool_workspace_bytes(int64_t n, int m) { return (size_t)nys::pool_chunks(n) * m * (2 + 64) * sizeof(float) + 64; }

extern "C" int mil_colsoftmax_pool_f32(const float* S, const float* V, int64_t ldv, int64_t n, int m, int dh, float* out, float* colmax, float* colsum,
                                       void* ws, size_t ws_bytes, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(S && V && out && colmax && colsum && ws && n > 0 && m > 0 && m <= 256 && dh > 0 && dh <= 64 && ldv >= dh,
                "mil_colsoftmax_pool_f32: bad arguments (m <= 256, dh <= 64)");
  MIL_CHECK_ARG(ws_bytes >= mil_colsoftmax_pool_workspace_bytes(n, m), "mil_colsoftmax_pool_f32: workspace too small");
  const int chunks = nys::pool_chunks(n);
  const int64_t per = (n + chunks - 1) / chunks;
  float* pmax = (float*)ws;
  float* pl = pmax + (size_t)chunks * m;
  float* pacc = pl + (size_t)chunks * m;
  nys::colmax_kernel<<<chunks, 256, 0, stream>>>(S, n, m, per, pmax);
  MIL_LAUNCH_CHECK();
  nys::colsoftmax_pool_kernel<<<chunks, 256, 0, stream>>>(S, V, ldv, n, m, dh, per, pmax, chunks, colmax, pl, pacc);
  MIL_LAUNCH_CHECK();
  nys::colsoftmax_finish_kernel<<<m, 64, 0, stream>>>(pl, pacc, chunks, m, dh, colsum, out);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_expdot_rows_f32(const float* S, int64_t rows, int m, const float* M, const float* w, float* out, mil_stream_t stream) {
  MIL_CHECK_ARG(S && M && w && out && rows >= 0 && m > 0, "mil_expdot_rows_f32: bad arguments");
  if (rows == 0) return 0;
  nys::expdot_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, rows, m, M, w, out);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_dwconv_tokens_f32(const float* v, int64_t ldv, int64_t rows, int heads, int dh, const float* w, int taps, float* out, int64_t ldo,
                                     int accumulate, mil_stream_t stream) {
  MIL_CHECK_ARG(v && w && out && rows > 0 && rows < (1ll << 31) && heads * dh > 0 && taps > 0 && (taps & 1) && taps <= nys::DW_MAXTAPS,
                "mil_dwconv_tokens_f32: bad arguments (odd taps <= %d)", nys::DW_MAXTAPS);
  const int C = heads * dh;
  nys::dwconv_tokens_kernel<<<dim3((unsigned)((rows + nys::DW_ROWS - 1) / nys::DW_ROWS), (C + nys::DW_COLS - 1) / nys::DW_COLS), 256, 0, (cudaStream_t)stream>>>(
      v, ldv, rows, heads, dh, w, taps, out, ldo, accumulate);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_ppeg_f32(const float* x, int H, int W, int C, const float* w49, const float* bias, float* y, mil_stream_t stream) {
  MIL_CHECK_ARG(x && w49 && y && H > 0 && W > 0 && C > 0, "mil_ppeg_f32: bad arguments");
  const int tiles = ((H + nys::PP_T - 1) / nys::PP_T) * ((W + nys::PP_T - 1) / nys::PP_T);
  nys::ppeg_kernel<<<dim3((unsigned)tiles, (C + nys::PP_C - 1) / nys::PP_C), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, w49, bias, y);
  MIL_LAUNCH_CHECK();
  return 0;
}
