// Row selection and the Merge cross-attention (MCA) of the student pass as the library's own kernels, forward and backward.
//
//  * take / split rows by a permutation: mask_fn (masking.py:91-110) gathers the kept rows x[mask_ids[:len_keep]] and Merge's random
//    keep (merge.py:163-174) splits x by argsort(rand(L)).  In both cases the index vector is (a prefix of) a PERMUTATION of the rows,
//    so the backward is a plain scatter in which every source row is written exactly once -- torch's index_put_(accumulate=True)
//    sorts the indices first (~35 us per call at L = 10 000, profiles/round2_train_step_attn.txt).
//  * MCA (merge.py:43-65): k <= 8 query tokens attend over the L_m dropped instances, 8 heads x 64: scores [8, k, L_m], softmax over
//    L_m, weighted sum of V.  torch runs it as two cuBLAS batched GEMMs with a skinny dimension (62 us each) plus a softmax; here it is
//    a score kernel, a softmax-statistics kernel and an output kernel (and their backward), all bandwidth-bound on kv [L_m, 1024].
#include "mil_common.cuh"

namespace mil {
namespace rows {

// out[i, :] = x[perm[i], :] for i < n_out (rows of `cols` floats, cols % 4 == 0)
__global__ void take_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ perm, int64_t n_out, int cols4, float4* __restrict__ out) {
  const int64_t i = blockIdx.x;
  if (i >= n_out) return;
  const float4* src = reinterpret_cast<const float4*>(x) + perm[i] * cols4;
  for (int c = threadIdx.x; c < cols4; c += blockDim.x) out[i * cols4 + c] = src[c];
}
// gx[perm[i], :] = i < n_a ? ga[i, :] : (gb ? gb[i - n_a, :] : 0)   for every i < n_rows (perm is a permutation of 0..n_rows-1)
__global__ void scatter_rows_kernel(const float4* __restrict__ ga, const float4* __restrict__ gb, const int64_t* __restrict__ perm, int64_t n_a,
                                    int64_t n_rows, int cols4, float4* __restrict__ gx) {
  const int64_t i = blockIdx.x;
  if (i >= n_rows) return;
  float4* dst = gx + perm[i] * cols4;
  const float4* src = i < n_a ? (ga ? ga + i * cols4 : nullptr) : (gb ? gb + (i - n_a) * cols4 : nullptr);
  for (int c = threadIdx.x; c < cols4; c += blockDim.x) dst[c] = src ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------- MCA: heads x dh = inner; q [kq, inner]; kv [L, 2 * inner] = [K | V] ----------------
// scores[h, a, j] = scale * q[a, h, :] . K[j, h, :]          grid (ceil(L / 8), heads), block 256 = 8 rows x 32 lanes
__global__ void __launch_bounds__(256) mca_scores_kernel(const float* __restrict__ q, const float* __restrict__ kv, int64_t L, int kq, int heads, int dh,
                                                         float scale, float* __restrict__ S) {
  extern __shared__ float sq[];                         // [kq][dh]
  const int h = blockIdx.y, inner = heads * dh;
  for (int i = threadIdx.x; i < kq * dh; i += blockDim.x) sq[i] = q[(i / dh) * inner + h * dh + i % dh];
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= L) return;
  const float* kr = kv + j * 2 * inner + h * dh;
  float acc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) acc[a] = 0.f;
  for (int d = lane; d < dh; d += 32) {
    const float kvv = kr[d];
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a < kq) acc[a] = fmaf(sq[a * dh + d], kvv, acc[a]);
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
    if (a < kq) {
      const float v = warp_sum(acc[a]);
      if (lane == 0) S[((int64_t)h * kq + a) * L + j] = v * scale;
    }
}
// per (h, a): m = max_j S, l = sum_j exp(S - m); S <- exp(S - m) / l  (the attention weights, kept for the backward)
__global__ void __launch_bounds__(1024) mca_softmax_kernel(float* __restrict__ S, int64_t L) {
  __shared__ float red[32];
  float* s = S + (int64_t)blockIdx.x * L;
  float m = -INFINITY;
  for (int64_t j = threadIdx.x; j < L; j += blockDim.x) m = fmaxf(m, s[j]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float l = 0.f;
  for (int64_t j = threadIdx.x; j < L; j += blockDim.x) {
    const float e = expf(s[j] - m);
    s[j] = e;
    l += e;
  }
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  l = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) l += red[w];
  const float inv = 1.f / l;
  for (int64_t j = threadIdx.x; j < L; j += blockDim.x) s[j] *= inv;
}
// partial[c, h, a, d] = sum over the chunk's rows j of P[h, a, j] * V[j, h, d]     grid (chunks, heads), block 256 = 4 row groups x 64 d
__global__ void __launch_bounds__(256) mca_pv_kernel(const float* __restrict__ P, const float* __restrict__ pmask, const float* __restrict__ kv, int64_t L,
                                                     int kq, int heads, int dh, int64_t rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[4][8][64];
  const int h = blockIdx.y, inner = heads * dh, d = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int64_t j0 = (int64_t)blockIdx.x * rows_per_chunk;
  int64_t j1 = j0 + rows_per_chunk;
  if (j1 > L) j1 = L;
  float acc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) acc[a] = 0.f;
  if (d < dh)
    for (int64_t j = j0 + rg; j < j1; j += 4) {
      const float v = kv[j * 2 * inner + inner + h * dh + d];
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (a < kq) {
          const int64_t o = ((int64_t)h * kq + a) * L + j;
          acc[a] = fmaf(pmask ? P[o] * pmask[o] : P[o], v, acc[a]);
        }
    }
#pragma unroll
  for (int a = 0; a < 8; ++a) red[rg][a][d] = acc[a];
  __syncthreads();
  if (rg == 0 && d < dh)
    for (int a = 0; a < kq; ++a)
      partial[(((int64_t)blockIdx.x * heads + h) * kq + a) * dh + d] = (red[0][a][d] + red[1][a][d]) + (red[2][a][d] + red[3][a][d]);
}
// out[a, h * dh + d] = sum_c partial[c, h, a, d]
__global__ void mca_reduce_kernel(const float* __restrict__ partial, int chunks, int kq, int heads, int dh, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // (h, a, d)
  if (i >= heads * kq * dh) return;
  float v = 0.f;
  for (int c = 0; c < chunks; ++c) v += partial[(int64_t)c * heads * kq * dh + i];
  const int h = i / (kq * dh), a = (i / dh) % kq, d = i % dh;
  out[a * heads * dh + h * dh + d] = v;
}
// backward, pass 1: dP[h, a, j] = g[a, h, :] . V[j, h, :]   (written over a scratch [heads, kq, L]); rowdot[h, a] += sum_j P dP (atomic-free:
// computed by mca_bwd_rowdot_kernel afterwards)
__global__ void __launch_bounds__(256) mca_bwd_dp_kernel(const float* __restrict__ g, const float* __restrict__ pmask, const float* __restrict__ kv, int64_t L,
                                                         int kq, int heads, int dh, float* __restrict__ dP) {
  extern __shared__ float sg[];                         // [kq][dh]
  const int h = blockIdx.y, inner = heads * dh;
  for (int i = threadIdx.x; i < kq * dh; i += blockDim.x) sg[i] = g[(i / dh) * inner + h * dh + i % dh];
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= L) return;
  const float* vr = kv + j * 2 * inner + inner + h * dh;
  float acc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) acc[a] = 0.f;
  for (int d = lane; d < dh; d += 32) {
    const float vv = vr[d];
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a < kq) acc[a] = fmaf(sg[a * dh + d], vv, acc[a]);
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
    if (a < kq) {
      const float v = warp_sum(acc[a]);
      const int64_t o = ((int64_t)h * kq + a) * L + j;
      if (lane == 0) dP[o] = pmask ? v * pmask[o] : v;
    }
}
// per (h, a): r = sum_j P dP;  dS[j] = scale * P[j] (dP[j] - r)   (in place over dP)
__global__ void __launch_bounds__(1024) mca_bwd_ds_kernel(const float* __restrict__ P, float* __restrict__ dP, int64_t L, float scale) {
  __shared__ float red[32];
  const float* p = P + (int64_t)blockIdx.x * L;
  float* dp = dP + (int64_t)blockIdx.x * L;
  float r = 0.f;
  for (int64_t j = threadIdx.x; j < L; j += blockDim.x) r = fmaf(p[j], dp[j], r);
  r = warp_sum(r);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = r;
  __syncthreads();
  r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
  for (int64_t j = threadIdx.x; j < L; j += blockDim.x) dp[j] = scale * p[j] * (dp[j] - r);
}
// d(kv)[j, :] : dK[j, h, d] = sum_a dS[h, a, j] q[a, h, d];  dV[j, h, d] = sum_a P[h, a, j] g[a, h, d]     one CTA per row j, 2 * inner threads' worth
__global__ void __launch_bounds__(256) mca_bwd_dkv_kernel(const float* __restrict__ P, const float* __restrict__ pmask, const float* __restrict__ dS,
                                                          const float* __restrict__ q, const float* __restrict__ g, int64_t L, int kq, int heads, int dh,
                                                          float* __restrict__ dkv) {
  const int64_t j = blockIdx.x;
  const int inner = heads * dh;
  for (int c = threadIdx.x; c < 2 * inner; c += blockDim.x) {
    const bool isv = c >= inner;
    const int cc = isv ? c - inner : c, h = cc / dh;
    const float* w = isv ? P : dS;
    const float* src = isv ? g : q;
    float v = 0.f;
    for (int a = 0; a < kq; ++a) {
      const int64_t o = ((int64_t)h * kq + a) * L + j;
      v = fmaf((isv && pmask) ? w[o] * pmask[o] : w[o], src[a * inner + cc], v);
    }
    dkv[j * 2 * inner + c] = v;
  }
}
// dq partial[c, h, a, d] = sum over the chunk of dS[h, a, j] K[j, h, d]: the same kernel as mca_pv with K instead of V -> pass kv offset 0
__global__ void __launch_bounds__(256) mca_sk_kernel(const float* __restrict__ dS, const float* __restrict__ kv, int64_t L, int kq, int heads, int dh,
                                                     int64_t rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[4][8][64];
  const int h = blockIdx.y, inner = heads * dh, d = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int64_t j0 = (int64_t)blockIdx.x * rows_per_chunk;
  int64_t j1 = j0 + rows_per_chunk;
  if (j1 > L) j1 = L;
  float acc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) acc[a] = 0.f;
  if (d < dh)
    for (int64_t j = j0 + rg; j < j1; j += 4) {
      const float v = kv[j * 2 * inner + h * dh + d];
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (a < kq) acc[a] = fmaf(dS[((int64_t)h * kq + a) * L + j], v, acc[a]);
    }
#pragma unroll
  for (int a = 0; a < 8; ++a) red[rg][a][d] = acc[a];
  __syncthreads();
  if (rg == 0 && d < dh)
    for (int a = 0; a < kq; ++a)
      partial[(((int64_t)blockIdx.x * heads + h) * kq + a) * dh + d] = (red[0][a][d] + red[1][a][d]) + (red[2][a][d] + red[3][a][d]);
}

static int mca_chunks(int64_t L) {
  int64_t c = (L + 127) / 128;
  if (c > 64) c = 64;
  return (int)(c < 1 ? 1 : c);
}

}  // namespace rows
}  // namespace mil

using namespace mil;

extern "C" int mil_take_rows_f32(const float* x, const int64_t* perm, int64_t n_out, int cols, float* out, mil_stream_t stream) {
  MIL_CHECK_ARG(x && perm && out && n_out >= 0 && cols > 0 && cols % 4 == 0, "mil_take_rows_f32: bad arguments (cols %% 4 == 0)");
  MIL_CHECK_ARG((uintptr_t)x % 16 == 0 && (uintptr_t)out % 16 == 0, "mil_take_rows_f32: pointers must be 16-byte aligned");
  if (n_out == 0) return 0;
  rows::take_rows_kernel<<<(unsigned)n_out, 128, 0, (cudaStream_t)stream>>>(x, perm, n_out, cols / 4, reinterpret_cast<float4*>(out));
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_scatter_rows_f32(const float* ga, const float* gb, const int64_t* perm, int64_t n_a, int64_t n_rows, int cols, float* gx,
                                    mil_stream_t stream) {
  MIL_CHECK_ARG(perm && gx && n_a >= 0 && n_a <= n_rows && cols > 0 && cols % 4 == 0, "mil_scatter_rows_f32: bad arguments");
  MIL_CHECK_ARG((uintptr_t)gx % 16 == 0 && (!ga || (uintptr_t)ga % 16 == 0) && (!gb || (uintptr_t)gb % 16 == 0), "mil_scatter_rows_f32: pointers must be 16-byte aligned");
  if (n_rows == 0) return 0;
  rows::scatter_rows_kernel<<<(unsigned)n_rows, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(ga), reinterpret_cast<const float4*>(gb), perm, n_a,
                                                                                 n_rows, cols / 4, reinterpret_cast<float4*>(gx));
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mil_mca_workspace_bytes(int64_t L, int kq, int heads, int dh) {
  return (size_t)rows::mca_chunks(L) * heads * kq * dh * sizeof(float) + 64;
}

extern "C" int mil_mca_fwd_f32(const float* q, const float* kv, int64_t L, int kq, int heads, int dh, float scale, const float* pmask, float* P,
                               float* out, void* ws, size_t ws_bytes, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(q && kv && P && out && ws && L > 0 && kq >= 1 && kq <= 8 && heads >= 1 && dh >= 1 && dh <= 64, "mil_mca_fwd_f32: bad arguments (kq <= 8, dh <= 64)");
  MIL_CHECK_ARG(ws_bytes >= mil_mca_workspace_bytes(L, kq, heads, dh), "mil_mca_fwd_f32: workspace too small");
  rows::mca_scores_kernel<<<dim3((unsigned)((L + 7) / 8), heads), 256, (size_t)kq * dh * 4, stream>>>(q, kv, L, kq, heads, dh, scale, P);
  MIL_LAUNCH_CHECK();
  rows::mca_softmax_kernel<<<heads * kq, 1024, 0, stream>>>(P, L);
  MIL_LAUNCH_CHECK();
  const int chunks = rows::mca_chunks(L);
  const int64_t per = (L + chunks - 1) / chunks;
  rows::mca_pv_kernel<<<dim3(chunks, heads), 256, 0, stream>>>(P, pmask, kv, L, kq, heads, dh, per, (float*)ws);
  MIL_LAUNCH_CHECK();
  rows::mca_reduce_kernel<<<(heads * kq * dh + 255) / 256, 256, 0, stream>>>((const float*)ws, chunks, kq, heads, dh, out);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_mca_bwd_f32(const float* g_out, const float* q, const float* kv, const float* P, const float* pmask, int64_t L, int kq, int heads,
                               int dh, float scale, float* dS_scratch, float* dq, float* dkv, void* ws, size_t ws_bytes, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(g_out && q && kv && P && dS_scratch && dq && dkv && ws && L > 0 && kq >= 1 && kq <= 8 && dh >= 1 && dh <= 64, "mil_mca_bwd_f32: bad arguments");
  MIL_CHECK_ARG(ws_bytes >= mil_mca_workspace_bytes(L, kq, heads, dh), "mil_mca_bwd_f32: workspace too small");
  rows::mca_bwd_dp_kernel<<<dim3((unsigned)((L + 7) / 8), heads), 256, (size_t)kq * dh * 4, stream>>>(g_out, pmask, kv, L, kq, heads, dh, dS_scratch);
  MIL_LAUNCH_CHECK();
  rows::mca_bwd_ds_kernel<<<heads * kq, 1024, 0, stream>>>(P, dS_scratch, L, scale);
  MIL_LAUNCH_CHECK();
  rows::mca_bwd_dkv_kernel<<<(unsigned)L, 256, 0, stream>>>(P, pmask, dS_scratch, q, g_out, L, kq, heads, dh, dkv);
  MIL_LAUNCH_CHECK();
  const int chunks = rows::mca_chunks(L);
  const int64_t per = (L + chunks - 1) / chunks;
  rows::mca_sk_kernel<<<dim3(chunks, heads), 256, 0, stream>>>(dS_scratch, kv, L, kq, heads, dh, per, (float*)ws);
  MIL_LAUNCH_CHECK();
  rows::mca_reduce_kernel<<<(heads * kq * dh + 255) / 256, 256, 0, stream>>>((const float*)ws, chunks, kq, heads, dh, dq);
  MIL_LAUNCH_CHECK();
  return 0;
}
