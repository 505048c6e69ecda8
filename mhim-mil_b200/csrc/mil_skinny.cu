// Skinny Linear layers of the path, forward and backward, as streaming CUDA-core kernels (exact fp32).
//
// Every pooling head ends in a Linear with ONE to a few outputs over all N instances (attention logit 128 -> 1: abmil.py:196,
// baseline.py:27; DSMIL instance classifier 512 -> C and its critical-instance logits: dsmil.py:62,93), and everything after the pooling
// works on ONE to a few rows (classifier / predictor on the pooled [1, 512]: abmil.py:238, mhim.py:267; Merge's to_q / to_out on k = 5
// tokens: merge.py:35-41; q(h_crit) on C rows: dsmil.py:92).  Round 1 sent all of them through the 128 x 128-tile FFMA GEMM: 20-56 us
// per call for a few KFLOP (profiles/round2_train_step_*.txt).  They are GEMV-shaped, i.e. bandwidth/latency bound:
//
//   "thin"  : N <= 8 outputs, M rows large.   fwd  Y[M,N] = act(X W^T + b)       one warp per row, W in shared memory
//                                              dW   [N,K] = sum_m G[m,:]^T X[m,:] row slices per CTA -> partials -> fixed-order reduce
//                                              dX   [M,K] = G W                   elementwise, W in shared memory
//   "short" : M <= 8 rows, N outputs large.   fwd  one warp per output column, the M input rows in shared memory
//                                              dW   [N,K] = sum_i G[i,:]^T X[i,:] elementwise (M outer products)
//                                              dX   [M,K] = G W                   one thread per (column k), loop over N
#include "mil_common.cuh"

namespace mil {
namespace skinny {

constexpr int MAXS = 8;          // the skinny dimension

// ---------------------------------------------------------------- thin: N <= 8
__global__ void __launch_bounds__(256) thin_fwd_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int K, const float* __restrict__ W,
                                                       const float* __restrict__ b, int N, int act, float* __restrict__ pre_out,
                                                       float* __restrict__ Y) {
  extern __shared__ float sw[];                        // [N][K]
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) sw[i] = W[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t m = (int64_t)blockIdx.x * 8 + warp; m < M; m += (int64_t)gridDim.x * 8) {
    const float* x = X + m * ldx;
    float acc[MAXS];
#pragma unroll
    for (int j = 0; j < MAXS; ++j) acc[j] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float xv = x[k];
#pragma unroll
      for (int j = 0; j < MAXS; ++j)
        if (j < N) acc[j] = fmaf(xv, sw[j * K + k], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < MAXS; ++j)
      if (j < N) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
      for (int j = 0; j < N; ++j) {
        const float v = acc[j] + (b ? b[j] : 0.f);
        if (pre_out) pre_out[m * N + j] = v;
        Y[m * N + j] = act_apply(v, act);
      }
    }
  }
}

// partial[blk][j][k] = sum over the CTA's rows of G[m,j] X[m,k];  partial_b[blk][j] = sum G[m,j].
// 256 threads = 8 row groups x 32 lanes; lane l owns columns k = l + 32 c; a thread keeps NV = N * ceil(K / 32) accumulators
// (a = j * kv + c), walks the rows rg, rg + 8, ... of the CTA's slice, and the 8 row groups are summed through shared memory in a
// fixed order.  Rows are independent loads, so the row loop pipelines (the first version walked 256 rows per thread serially).
template <int NV>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float* __restrict__ G, const float* __restrict__ X, int64_t ldx, int64_t M, int K, int N,
                                                         int64_t rows_per_cta, float* __restrict__ partial, float* __restrict__ partial_b) {
  extern __shared__ float red[];                        // [8][N * kv * 32 + N]
  const int rg = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv = (K + 31) / 32;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t m1 = m0 + rows_per_cta;
  if (m1 > M) m1 = M;
  float acc[NV];
#pragma unroll
  for (int a = 0; a < NV; ++a) acc[a] = 0.f;
  float accb[MAXS];
#pragma unroll
  for (int j = 0; j < MAXS; ++j) accb[j] = 0.f;
  for (int64_t m = m0 + rg; m < m1; m += 8) {
    const float* x = X + m * ldx;
    float g[MAXS];
#pragma unroll
    for (int j = 0; j < MAXS; ++j) g[j] = j < N ? G[m * N + j] : 0.f;
#pragma unroll
    for (int j = 0; j < MAXS; ++j) accb[j] += g[j];
#pragma unroll
    for (int a = 0; a < NV; ++a) {
      const int j = a / kv, c = a - j * kv, k = lane + 32 * c;
      if (a < N * kv && k < K) acc[a] = fmaf(g[j], x[k], acc[a]);
    }
  }
  const int per = N * kv * 32 + N;
  float* mine = red + rg * per;
#pragma unroll
  for (int a = 0; a < NV; ++a)
    if (a < N * kv) mine[a * 32 + lane] = acc[a];
  if (lane == 0)
    for (int j = 0; j < N; ++j) mine[N * kv * 32 + j] = accb[j];
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * N * K;
  for (int i = threadIdx.x; i < N * kv * 32; i += 256) {
    const int a = i >> 5, l = i & 31, j = a / kv, c = a - j * kv, k = l + 32 * c;
    if (k < K) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) v += red[r * per + i];
      out[j * K + k] = v;
    }
  }
  if (partial_b && (int)threadIdx.x < N) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) v += red[r * per + N * kv * 32 + threadIdx.x];
    partial_b[(int64_t)blockIdx.x * N + threadIdx.x] = v;
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n_part, int64_t n, float* __restrict__ out,
                                       const float* __restrict__ partial_b, int nb, float* __restrict__ out_b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float a = 0.f;
    for (int p = 0; p < n_part; ++p) a += partial[(int64_t)p * n + i];
    out[i] = a;
  }
  if (out_b && i < nb) {
    float a = 0.f;
    for (int p = 0; p < n_part; ++p) a += partial_b[(int64_t)p * nb + i];
    out_b[i] = a;
  }
}

// dX[m,k] = sum_j G[m,j] W[j,k]
__global__ void __launch_bounds__(256) thin_dx_kernel(const float* __restrict__ G, const float* __restrict__ W, int64_t M, int K, int N,
                                                      float* __restrict__ dX) {
  extern __shared__ float sw[];                        // [N][K]
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) sw[i] = W[i];
  __syncthreads();
  for (int64_t m = blockIdx.x; m < M; m += gridDim.x) {
    float g[MAXS];
#pragma unroll
    for (int j = 0; j < MAXS; ++j) g[j] = j < N ? G[m * N + j] : 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < MAXS; ++j)
        if (j < N) a = fmaf(g[j], sw[j * K + k], a);
      dX[m * K + k] = a;
    }
  }
}

// ---------------------------------------------------------------- short: M <= 8
__global__ void __launch_bounds__(256) short_fwd_kernel(const float* __restrict__ X, int64_t ldx, int M, int K, const float* __restrict__ W,
                                                        const float* __restrict__ b, int N, int act, float* __restrict__ pre_out,
                                                        float* __restrict__ Y) {
  extern __shared__ float sx[];                        // [M][K]
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) sx[i] = X[(int64_t)(i / K) * ldx + i % K];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = blockIdx.x * 8 + warp; n < N; n += gridDim.x * 8) {
    const float* w = W + (int64_t)n * K;
    float acc[MAXS];
#pragma unroll
    for (int i = 0; i < MAXS; ++i) acc[i] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = w[k];
#pragma unroll
      for (int i = 0; i < MAXS; ++i)
        if (i < M) acc[i] = fmaf(sx[i * K + k], wv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < MAXS; ++i)
      if (i < M) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
      const float bv = b ? b[n] : 0.f;
      for (int i = 0; i < M; ++i) {
        const float v = acc[i] + bv;
        if (pre_out) pre_out[(int64_t)i * N + n] = v;
        Y[(int64_t)i * N + n] = act_apply(v, act);
      }
    }
  }
}

// dW[n,k] = sum_i G[i,n] X[i,k];  db[n] = sum_i G[i,n]
__global__ void short_wgrad_kernel(const float* __restrict__ G, const float* __restrict__ X, int64_t ldx, int M, int K, int N,
                                   float* __restrict__ dW, float* __restrict__ db) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (int64_t)N * K) {
    const int n = (int)(idx / K), k = (int)(idx % K);
    float a = 0.f;
    for (int i = 0; i < M; ++i) a = fmaf(G[(int64_t)i * N + n], X[(int64_t)i * ldx + k], a);
    dW[idx] = a;
  }
  if (db && idx < N) {
    float a = 0.f;
    for (int i = 0; i < M; ++i) a += G[(int64_t)i * N + idx];
    db[idx] = a;
  }
}

// dX[i,k] = sum_n G[i,n] W[n,k].  One CTA per 32 columns k; its 8 warps split the N rows of W (warp w: n = w, w + 8, ...; lanes =
// 32 consecutive k, coalesced), then a fixed-order sum over the 8 warps through shared memory.
__global__ void __launch_bounds__(256) short_dx_kernel(const float* __restrict__ G, const float* __restrict__ W, int M, int K, int N,
                                                       float* __restrict__ dX) {
  extern __shared__ float sgm[];                       // [M][N] | [8][MAXS][32]
  float* red = sgm + M * N;
  for (int i = threadIdx.x; i < M * N; i += blockDim.x) sgm[i] = G[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 32 + lane;
  float acc[MAXS];
#pragma unroll
  for (int i = 0; i < MAXS; ++i) acc[i] = 0.f;
  if (k < K) {
#pragma unroll 4
    for (int n = warp; n < N; n += 8) {
      const float wv = W[(int64_t)n * K + k];
#pragma unroll
      for (int i = 0; i < MAXS; ++i)
        if (i < M) acc[i] = fmaf(sgm[i * N + n], wv, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < MAXS; ++i) red[(warp * MAXS + i) * 32 + lane] = acc[i];
  __syncthreads();
  if (warp < M && k < K) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[(w * MAXS + warp) * 32 + lane];
    dX[(int64_t)warp * K + k] = v;
  }
}

static int thin_ctas(int64_t M) {
  int64_t c = (M + 63) / 64;                           // >= 64 rows per CTA (8 per thread)
  const int cap = 2 * num_sms();
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  return (int)c;
}

}  // namespace skinny
}  // namespace mil

using namespace mil;

extern "C" int mil_skinny_supported(int64_t M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0 || K > 1536) return 0;
  // thin: W fits shared memory and the weight gradient fits 64 accumulators per thread (N * ceil(K / 32) <= 64)
  if (N <= skinny::MAXS) return ((size_t)N * K * 4 <= 48 * 1024 && N * ((K + 31) / 32) <= 64) ? 1 : 0;
  if (M <= skinny::MAXS) return ((size_t)M * K * 4 <= 48 * 1024 && (size_t)M * N * 4 + 8 * skinny::MAXS * 32 * 4 <= 48 * 1024) ? 2 : 0;
  return 0;
}

extern "C" size_t mil_skinny_workspace_bytes(int64_t M, int N, int K) {
  if (mil_skinny_supported(M, N, K) != 1) return 16;
  return (size_t)skinny::thin_ctas(M) * ((size_t)N * K + N) * sizeof(float) + 16;
}

extern "C" int mil_skinny_fwd_f32(const float* X, int64_t ldx, int64_t M, int K, const float* W, const float* b, int N, int act,
                                  float* pre_out, float* Y, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int kind = mil_skinny_supported(M, N, K);
  MIL_CHECK_ARG(X && W && Y && ldx >= K, "mil_skinny_fwd_f32: bad argument");
  MIL_CHECK_ARG(kind != 0, "mil_skinny_fwd_f32: needs N <= 8 or M <= 8 and K <= 1536 (got M=%lld N=%d K=%d)", (long long)M, N, K);
  if (kind == 1) {
    int64_t blocks = (M + 7) / 8;
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    skinny::thin_fwd_kernel<<<(unsigned)blocks, 256, (size_t)N * K * 4, stream>>>(X, ldx, M, K, W, b, N, act, pre_out, Y);
  } else {
    int blocks = (N + 7) / 8;
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    skinny::short_fwd_kernel<<<blocks, 256, (size_t)M * K * 4, stream>>>(X, ldx, (int)M, K, W, b, N, act, pre_out, Y);
  }
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_skinny_bwd_f32(const float* G, const float* X, int64_t ldx, const float* W, int64_t M, int N, int K, float* dW, float* db,
                                  float* dX, void* ws, size_t ws_bytes, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int kind = mil_skinny_supported(M, N, K);
  MIL_CHECK_ARG(G && kind != 0, "mil_skinny_bwd_f32: needs N <= 8 or M <= 8 and K <= 1536 (got M=%lld N=%d K=%d)", (long long)M, N, K);
  MIL_CHECK_ARG((!dW && !db) || X, "mil_skinny_bwd_f32: the weight gradient needs X");
  MIL_CHECK_ARG(!dX || W, "mil_skinny_bwd_f32: the input gradient needs W");
  if (kind == 1) {
    if (dW || db) {
      MIL_CHECK_ARG(dW, "mil_skinny_bwd_f32: db comes with dW");
      MIL_CHECK_ARG(ws && ws_bytes >= mil_skinny_workspace_bytes(M, N, K), "mil_skinny_bwd_f32: workspace needs %zu bytes", mil_skinny_workspace_bytes(M, N, K));
      const int ctas = skinny::thin_ctas(M);
      const int64_t rows = (M + ctas - 1) / ctas;
      float* partial = (float*)ws;
      float* partial_b = partial + (size_t)ctas * N * K;
      const int nv = N * ((K + 31) / 32);
      const size_t sm = (size_t)8 * (nv * 32 + N) * sizeof(float);
      if (nv <= 4) skinny::thin_wgrad_kernel<4><<<ctas, 256, sm, stream>>>(G, X, ldx, M, K, N, rows, partial, db ? partial_b : nullptr);
      else if (nv <= 16) skinny::thin_wgrad_kernel<16><<<ctas, 256, sm, stream>>>(G, X, ldx, M, K, N, rows, partial, db ? partial_b : nullptr);
      else if (nv <= 32) skinny::thin_wgrad_kernel<32><<<ctas, 256, sm, stream>>>(G, X, ldx, M, K, N, rows, partial, db ? partial_b : nullptr);
      else {
        MIL_CUDA(cudaFuncSetAttribute(skinny::thin_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        skinny::thin_wgrad_kernel<64><<<ctas, 256, sm, stream>>>(G, X, ldx, M, K, N, rows, partial, db ? partial_b : nullptr);
      }
      MIL_LAUNCH_CHECK();
      const int64_t n = (int64_t)N * K;
      skinny::reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(partial, ctas, n, dW, partial_b, N, db);
      MIL_LAUNCH_CHECK();
    }
    if (dX) {
      int64_t blocks = M < 4 * num_sms() ? M : 4 * num_sms();
      skinny::thin_dx_kernel<<<(unsigned)blocks, 256, (size_t)N * K * 4, stream>>>(G, W, M, K, N, dX);
      MIL_LAUNCH_CHECK();
    }
  } else {
    if (dW) {
      const int64_t n = (int64_t)N * K;
      skinny::short_wgrad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(G, X, ldx, (int)M, K, N, dW, db);
      MIL_LAUNCH_CHECK();
    }
    if (dX) {
      skinny::short_dx_kernel<<<(K + 31) / 32, 256, (size_t)M * N * 4 + 8 * skinny::MAXS * 32 * 4, stream>>>(G, W, (int)M, K, N, dX);
      MIL_LAUNCH_CHECK();
    }
  }
  return 0;
}
