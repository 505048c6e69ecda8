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

// partial[blk][j][k] = sum over the CTA's rows of G[m,j] X[m,k];  partial_b[blk][j] = sum G[m,j]
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float* __restrict__ G, const float* __restrict__ X, int64_t ldx, int64_t M, int K, int N,
                                                         int64_t rows_per_cta, float* __restrict__ partial, float* __restrict__ partial_b) {
  __shared__ float sg[64][MAXS];
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t m1 = m0 + rows_per_cta;
  if (m1 > M) m1 = M;
  constexpr int CPT = 6;                               // columns per thread: K <= 1536
  float acc[MAXS][CPT];
#pragma unroll
  for (int j = 0; j < MAXS; ++j)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[j][c] = 0.f;
  float accb = 0.f;                                    // thread j < N: bias partial
  for (int64_t mb = m0; mb < m1; mb += 64) {
    const int nr = (int)((m1 - mb) < 64 ? (m1 - mb) : 64);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * N; i += blockDim.x) sg[i / N][i % N] = G[(mb + i / N) * N + i % N];
    __syncthreads();
    if ((int)threadIdx.x < N)
      for (int r = 0; r < nr; ++r) accb += sg[r][threadIdx.x];
    for (int r = 0; r < nr; ++r) {
      const float* x = X + (mb + r) * ldx;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int k = threadIdx.x + c * 256;
        if (k < K) {
          const float xv = x[k];
#pragma unroll
          for (int j = 0; j < MAXS; ++j)
            if (j < N) acc[j][c] = fmaf(sg[r][j], xv, acc[j][c]);
        }
      }
    }
  }
  float* out = partial + (int64_t)blockIdx.x * N * K;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int k = threadIdx.x + c * 256;
    if (k < K)
#pragma unroll
      for (int j = 0; j < MAXS; ++j)
        if (j < N) out[j * K + k] = acc[j][c];
  }
  if (partial_b && (int)threadIdx.x < N) partial_b[(int64_t)blockIdx.x * N + threadIdx.x] = accb;
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

// dX[i,k] = sum_n G[i,n] W[n,k]
__global__ void __launch_bounds__(128) short_dx_kernel(const float* __restrict__ G, const float* __restrict__ W, int M, int K, int N,
                                                       float* __restrict__ dX) {
  extern __shared__ float sgm[];                       // [M][N]
  for (int i = threadIdx.x; i < M * N; i += blockDim.x) sgm[i] = G[i];
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc[MAXS];
#pragma unroll
  for (int i = 0; i < MAXS; ++i) acc[i] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float wv = W[(int64_t)n * K + k];
#pragma unroll
    for (int i = 0; i < MAXS; ++i)
      if (i < M) acc[i] = fmaf(sgm[i * N + n], wv, acc[i]);
  }
  for (int i = 0; i < M; ++i) dX[(int64_t)i * K + k] = acc[i];
}

static int thin_ctas(int64_t M) {
  int64_t c = (M + 255) / 256;                         // >= 256 rows per CTA
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
  if (N <= skinny::MAXS) return (size_t)N * K * 4 <= 48 * 1024 ? 1 : 0;
  if (M <= skinny::MAXS) return ((size_t)M * K * 4 <= 48 * 1024 && (size_t)M * N * 4 <= 48 * 1024) ? 2 : 0;
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
      skinny::thin_wgrad_kernel<<<ctas, 256, 0, stream>>>(G, X, ldx, M, K, N, rows, partial, db ? partial_b : nullptr);
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
      skinny::short_dx_kernel<<<(K + 127) / 128, 128, (size_t)M * N * 4, stream>>>(G, W, (int)M, K, N, dX);
      MIL_LAUNCH_CHECK();
    }
  }
  return 0;
}
