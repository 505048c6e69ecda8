// fp32 CUDA-core kernels: exact-fp32 GEMM (NT / TN / NN through strides), activation backward, column sums,
// softmax-over-instances pooling (forward, backward, partial merge) and the teacher CAM score.
// These are the parity reference on the GPU and the backward path; the headline forward is mil_fused_sm100.cu.
#include "mil_common.cuh"

namespace mil {

// ---------------------------------------------------------------------------------------------------------
// SGEMM: C[m,n] = act(sum_k A(m,k) B(n,k) + bias[n])
// ---------------------------------------------------------------------------------------------------------
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_THREADS = 256;

struct SgemmParams {
  const float* A; int64_t sAm, sAk; const int64_t* row_ids;
  const float* B; int64_t sBn, sBk;
  const float* bias; float* C; int64_t ldc; float* pre_out;
  int64_t M, N, K; int act; int64_t k_per_split; float* partial; int vecA, vecB;
  int batch; int64_t bA, bB, bC;      // batch > 1: blockIdx.z = batch index (element strides bA / bB / bC), no split-K
};

// Load a [128 x 8] operand tile into registers (4 floats per thread).
// KC (k contiguous): thread -> (row = t/2, k = (t%2)*4 .. +3).  MC (row contiguous): thread -> (k = t/32, row = (t%32)*4 .. +3).
template <bool KC>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t s_row, int64_t s_k, const int64_t* __restrict__ row_ids,
                                          int64_t row0, int64_t nrows, int64_t k0, int64_t kend, int vec, float (&v)[4]) {
  const int t = threadIdx.x;
  if (KC) {
    const int64_t r = row0 + (t >> 1);
    const int64_t k = k0 + (t & 1) * 4;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (r < nrows) {
      const int64_t rr = row_ids ? row_ids[r] : r;
      const float* p = P + rr * s_row + k;
      if (vec && k + 3 < kend) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (k + j < kend) v[j] = p[j];
      }
    }
  } else {
    const int64_t k = k0 + (t >> 5);
    const int64_t r = row0 + (t & 31) * 4;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (k < kend) {
      const float* p = P + k * s_k + r;
      if (vec && r + 3 < nrows) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (r + j < nrows) v[j] = p[j];
      }
    }
  }
}

template <bool KC>
__device__ __forceinline__ void store_tile(float (*S)[SG_BM], const float (&v)[4]) {
  const int t = threadIdx.x;
  if (KC) {
    const int r = t >> 1, k = (t & 1) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) S[k + j][r] = v[j];
  } else {
    const int k = t >> 5, r = (t & 31) * 4;
    *reinterpret_cast<float4*>(&S[k][r]) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(SG_THREADS, 2) sgemm_kernel(const SgemmParams p) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN];
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
  const bool batched = p.batch > 1;
  const float* __restrict__ Ap = batched ? p.A + (int64_t)blockIdx.z * p.bA : p.A;
  const float* __restrict__ Bp = batched ? p.B + (int64_t)blockIdx.z * p.bB : p.B;
  float* __restrict__ Cp = batched ? p.C + (int64_t)blockIdx.z * p.bC : p.C;
  const int64_t kbeg = batched ? 0 : (int64_t)blockIdx.z * p.k_per_split;
  const int64_t kend = batched ? p.K : min(p.K, kbeg + p.k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  load_tile<A_KC>(Ap, p.sAm, p.sAk, p.row_ids, m0, p.M, kbeg, kend, p.vecA, ra);
  load_tile<B_KC>(Bp, p.sBn, p.sBk, nullptr, n0, p.N, kbeg, kend, p.vecB, rb);
  store_tile<A_KC>(As[0], ra);
  store_tile<B_KC>(Bs[0], rb);
  __syncthreads();

  int buf = 0;
  for (int64_t k0 = kbeg; k0 < kend; k0 += SG_BK) {
    const bool more = k0 + SG_BK < kend;
    if (more) {
      load_tile<A_KC>(Ap, p.sAm, p.sAk, p.row_ids, m0, p.M, k0 + SG_BK, kend, p.vecA, ra);
      load_tile<B_KC>(Bp, p.sBn, p.sBk, nullptr, n0, p.N, k0 + SG_BK, kend, p.vecB, rb);
    }
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile<A_KC>(As[buf ^ 1], ra);
      store_tile<B_KC>(Bs[buf ^ 1], rb);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.partial) {
        p.partial[((int64_t)blockIdx.z * p.M + m) * p.N + n] = v;
      } else {
        if (p.bias) v += p.bias[n];
        if (p.pre_out) p.pre_out[m * p.ldc + n] = v;
        Cp[m * p.ldc + n] = act_apply(v, p.act);
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t M, int64_t N, const float* __restrict__ bias,
                                     float* __restrict__ C, int64_t ldc, float* __restrict__ pre_out, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += partial[(int64_t)z * M * N + i];   // fixed order: deterministic
  const int64_t m = i / N, n = i % N;
  if (bias) v += bias[n];
  if (pre_out) pre_out[m * ldc + n] = v;
  C[m * ldc + n] = act_apply(v, act);
}

// ---------------------------------------------------------------------------------------------------------
// elementwise activation backward, column sums
// ---------------------------------------------------------------------------------------------------------
__global__ void act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t n, int act, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = y[i];
  float d;
  switch (act) {
    case MIL_ACT_RELU: d = v > 0.f ? 1.f : 0.f; break;
    case MIL_ACT_TANH: d = 1.f - v * v; break;
    case MIL_ACT_SIGMOID: d = v * (1.f - v); break;
    case MIL_ACT_GELU:   // v is the PRE-activation: Phi(v) + v phi(v)
      d = 0.5f * (1.f + erff(v * 0.70710678118654752440f)) + v * 0.39894228040143267794f * expf(-0.5f * v * v);
      break;
    default: d = 1.f;
  }
  out[i] = g[i] * d;
}

// the same through a dropout that followed the activation: one thread per (row, 32-column chunk)
__global__ void act_bwd_drop_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t rows, int words_per_row, int act,
                                    int mode, const uint32_t* __restrict__ bits, uint32_t thresh, float scale, uint32_t s0, uint32_t s1,
                                    uint32_t o0, uint32_t o1, float* __restrict__ out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= rows * words_per_row) return;
  if (mode == 3) { s0 = bits[0]; s1 = bits[1]; o0 = bits[2]; o1 = bits[3]; }
  const uint32_t seed[2] = {s0, s1}, off[2] = {o0, o1};
  const uint32_t word = mode == 1 ? bits[w] : mil::philox_keep_word((uint32_t)(w / words_per_row), (uint32_t)(w % words_per_row), thresh, seed, off);
  const float4* g4 = reinterpret_cast<const float4*>(g + w * 32);
  const float4* y4 = reinterpret_cast<const float4*>(y + w * 32);
  float4* o4 = reinterpret_cast<float4*>(out + w * 32);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 gv = g4[q], yv = y4[q];
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, vv[4] = {yv.x, yv.y, yv.z, yv.w};
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = vv[e];
      float d;
      switch (act) {   // v: relu -> the (dropped or undropped) output; every other activation -> the PRE-activation
        case MIL_ACT_RELU: d = v > 0.f ? 1.f : 0.f; break;
        case MIL_ACT_TANH: { const float t = tanhf(v); d = 1.f - t * t; break; }
        case MIL_ACT_SIGMOID: { const float t = 1.f / (1.f + expf(-v)); d = t * (1.f - t); break; }
        case MIL_ACT_GELU: d = 0.5f * (1.f + erff(v * 0.70710678118654752440f)) + v * 0.39894228040143267794f * expf(-0.5f * v * v); break;
        default: d = 1.f;
      }
      r[e] = ((word >> (q * 4 + e)) & 1u) ? gg[e] * scale * d : 0.f;
    }
    o4[q] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

__global__ void colsum_partial_kernel(const float* __restrict__ A, int64_t M, int64_t N, int64_t rows_per_slice, float* __restrict__ ws) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice, r1 = min(M, r0 + rows_per_slice);
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += A[r * N + n];
  ws[(int64_t)blockIdx.y * N + n] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ ws, int slices, int64_t N, float* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int z = 0; z < slices; ++z) s += ws[(int64_t)z * N + n];
  out[n] = s;
}

// ---------------------------------------------------------------------------------------------------------
// softmax over instances + weighted pooling
// ---------------------------------------------------------------------------------------------------------
constexpr int POOL_THREADS = 256;   // 8 warps; warp w takes rows w, w+8, ...; lane takes float4 columns lane + 32 j

template <int NV>   // H = NV * 128
__global__ void __launch_bounds__(POOL_THREADS) softmax_pool_partial_kernel(const float* __restrict__ s, int64_t s_stride, const float* __restrict__ h,
                                                                            int64_t L, const uint8_t* __restrict__ keep, int64_t rows_per_block,
                                                                            float* __restrict__ part) {
  constexpr int H = NV * 128;
  __shared__ float red[8];
  __shared__ __align__(16) float accs[8][H];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(L, r0 + rows_per_block);

  float m = -INFINITY;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += POOL_THREADS)
    if (!keep || keep[r]) m = fmaxf(m, s[r * s_stride]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();

  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float l = 0.f;
  if (m > -INFINITY) {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      if (keep && !keep[r]) continue;
      const float e = expf(s[r * s_stride] - m);
      l += e;
      const float4* hr = reinterpret_cast<const float4*>(h + r * H);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 v = hr[lane + 32 * j];
        acc[j].x = fmaf(e, v.x, acc[j].x); acc[j].y = fmaf(e, v.y, acc[j].y);
        acc[j].z = fmaf(e, v.z, acc[j].z); acc[j].w = fmaf(e, v.w, acc[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) *reinterpret_cast<float4*>(&accs[warp][(lane + 32 * j) * 4]) = acc[j];
  if (lane == 0) red[warp] = l;
  __syncthreads();
  float* out = part + (int64_t)blockIdx.x * (2 + H);
  for (int c = threadIdx.x; c < H; c += POOL_THREADS) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += accs[w][c];
    out[2 + c] = v;
  }
  if (threadIdx.x == 0) {
    float lt = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) lt += red[w];
    out[0] = m;
    out[1] = lt;
  }
}

// Merge n_part partials (m_i, l_i, P_i[H]).  Grid: one block per 64 columns; block = 64 columns x 4 row groups.
// Every block recomputes the (tiny) global max / denominator so no second launch or grid sync is needed.
constexpr int MERGE_COLS = 64, MERGE_GROUPS = 4, MERGE_MAXP = 1024;
__global__ void __launch_bounds__(MERGE_COLS * MERGE_GROUPS) pool_merge_kernel(const float* __restrict__ part, int n_part, int H,
                                                                                float* __restrict__ stats, float* __restrict__ pooled) {
  __shared__ float wgt[MERGE_MAXP];              // exp(m_i - m) (0 for idle partials)
  __shared__ float red[8];
  __shared__ float acc[MERGE_GROUPS][MERGE_COLS];
  const int stride = 2 + H, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float m = -INFINITY;
  for (int i = tid; i < n_part; i += blockDim.x)
    if (part[(int64_t)i * stride + 1] > 0.f) m = fmaxf(m, part[(int64_t)i * stride]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float l = 0.f;
  for (int i = tid; i < n_part; i += blockDim.x) {
    const float li = part[(int64_t)i * stride + 1];
    const float w = li > 0.f ? expf(part[(int64_t)i * stride] - m) : 0.f;
    wgt[i] = w;
    l += li * w;
  }
  l = warp_sum(l);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) l += red[w];   // fixed order: identical in every block
  const int col = blockIdx.x * MERGE_COLS + (tid & (MERGE_COLS - 1)), grp = tid / MERGE_COLS;
  float v = 0.f;
  if (col < H) {
    const int per = (n_part + MERGE_GROUPS - 1) / MERGE_GROUPS;
    const int i0 = grp * per, i1 = min(n_part, i0 + per);
#pragma unroll 4
    for (int i = i0; i < i1; ++i) v = fmaf(part[(int64_t)i * stride + 2 + col], wgt[i], v);
  }
  acc[grp][tid & (MERGE_COLS - 1)] = v;
  __syncthreads();
  if (grp == 0 && col < H) pooled[col] = (acc[0][tid] + acc[1][tid] + acc[2][tid] + acc[3][tid]) / l;
  if (blockIdx.x == 0 && tid == 0) { stats[0] = m; stats[1] = l; }
}

// One CTA: log-sum-exp merge of the per-rank records + classifier (the tail of an instance-sharded forward).
__global__ void __launch_bounds__(512) shard_merge_cls_kernel(const float* __restrict__ rec, int n_rec, int H, const float* __restrict__ Wcls,
                                                              const float* __restrict__ bcls, int n_cls, float* __restrict__ stats,
                                                              float* __restrict__ pooled, float* __restrict__ logits) {
  extern __shared__ float sm[];                  // [n_rec] weights | [H] pooled
  float* wgt = sm;
  float* pl = sm + n_rec;
  __shared__ float ml[2];
  const int stride = 2 + H, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    float m = -INFINITY;
    for (int i = lane; i < n_rec; i += 32)
      if (rec[(int64_t)i * stride + 1] > 0.f) m = fmaxf(m, rec[(int64_t)i * stride]);
    m = warp_max(m);
    float l = 0.f;
    for (int i0 = 0; i0 < n_rec; i0 += 32) {     // fixed order
      const int i = i0 + lane;
      float li = 0.f, w = 0.f;
      if (i < n_rec) {
        li = rec[(int64_t)i * stride + 1];
        w = li > 0.f ? expf(rec[(int64_t)i * stride] - m) : 0.f;
        wgt[i] = w;
      }
      l += warp_sum(li * w);
    }
    if (lane == 0) { ml[0] = m; ml[1] = l; stats[0] = m; stats[1] = l; }
  }
  __syncthreads();
  const float l = ml[1];
  for (int c = tid; c < H; c += blockDim.x) {
    float v = 0.f;
    for (int i = 0; i < n_rec; ++i) v = fmaf(rec[(int64_t)i * stride + 2 + c], wgt[i], v);
    v /= l;
    pl[c] = v;
    pooled[c] = v;
  }
  __syncthreads();
  if (logits)
    for (int k = warp; k < n_cls; k += (int)(blockDim.x >> 5)) {
      float a = 0.f;
      for (int c = lane; c < H; c += 32) a = fmaf(pl[c], Wcls[(int64_t)k * H + c], a);
      a = warp_sum(a);
      if (lane == 0) logits[k] = a + (bcls ? bcls[k] : 0.f);
    }
}

__global__ void attn_norm_kernel(const float* __restrict__ s, int64_t s_stride, int64_t L, const uint8_t* __restrict__ keep,
                                 const float* __restrict__ stats, float* __restrict__ attn) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  attn[i] = (keep && !keep[i]) ? 0.f : expf(s[i * s_stride] - stats[0]) / stats[1];
}

// one warp per row: g_s[n] = a_n (h_n.g_p - pooled.g_p); g_h[n,:] (+)= a_n g_p
__global__ void __launch_bounds__(256) softmax_pool_bwd_kernel(const float* __restrict__ s, int64_t s_stride, const float* __restrict__ h, int64_t L, int H,
                                                               const uint8_t* __restrict__ keep, const float* __restrict__ stats, const float* __restrict__ pooled,
                                                               const float* __restrict__ g_p, float* __restrict__ g_s, int64_t gs_stride,
                                                               float* __restrict__ g_h, int accumulate) {
  __shared__ float kappa_s;
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float kp = 0.f;
  for (int c = threadIdx.x; c < H; c += blockDim.x) kp = fmaf(pooled[c], g_p[c], kp);
  kp = warp_sum(kp);
  if (lane == 0) red[warp] = kp;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    kappa_s = t;
  }
  __syncthreads();
  const float kappa = kappa_s, m = stats[0], inv_l = 1.f / stats[1];
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < L; r += (int64_t)gridDim.x * 8) {
    const bool on = !keep || keep[r];
    const float a = on ? expf(s[r * s_stride] - m) * inv_l : 0.f;
    float dot = 0.f;
    if (on)
      for (int c = lane; c < H; c += 32) dot = fmaf(h[r * H + c], g_p[c], dot);
    dot = warp_sum(dot);
    if (lane == 0) g_s[r * gs_stride] = a * (dot - kappa);
    if (g_h) {
      for (int c = lane; c < H; c += 32) {
        const float v = a * g_p[c];
        g_h[r * H + c] = accumulate ? g_h[r * H + c] + v : v;
      }
    }
  }
}

__global__ void cam_score_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t L, int C, const float* __restrict__ stats,
                                 float bias0, const float* __restrict__ bias_dev, float* __restrict__ score) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  if (bias_dev) bias0 = bias_dev[0];
  const float a = expf(s[i] - stats[0]) / stats[1];
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, fmaf(a, t[i * C + c], bias0));
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(fmaf(a, t[i * C + c], bias0) - mx);
  score[i] = 1.f / den;     // max_c softmax_c = exp(0) / den
}

// Multi-tensor EMA: block b owns segment b (<= a few 10^4 elements of one parameter).  k = fma(1 - mm, q, k * mm) -- the two roundings of
// `param_k.mul_(mm).add_(param_q, alpha=1 - mm)`: the product k * mm is rounded on its own, alpha * q + that is fused.
__global__ void ema_update_kernel(const mil_ema_seg_t* __restrict__ segs, float mm, float alpha) {
  const mil_ema_seg_t sg = segs[blockIdx.x];
  float* __restrict__ k = sg.dst;
  const float* __restrict__ q = sg.src;
  const int64_t n = sg.n;
  if ((((uintptr_t)k | (uintptr_t)q) & 15) == 0) {
    const int64_t n4 = n >> 2;
    float4* k4 = reinterpret_cast<float4*>(k);
    const float4* q4 = reinterpret_cast<const float4*>(q);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 a = k4[i];
      const float4 b = q4[i];
      a.x = fmaf(alpha, b.x, __fmul_rn(a.x, mm));
      a.y = fmaf(alpha, b.y, __fmul_rn(a.y, mm));
      a.z = fmaf(alpha, b.z, __fmul_rn(a.z, mm));
      a.w = fmaf(alpha, b.w, __fmul_rn(a.w, mm));
      k4[i] = a;
    }
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) k[i] = fmaf(alpha, q[i], __fmul_rn(k[i], mm));
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) k[i] = fmaf(alpha, q[i], __fmul_rn(k[i], mm));
  }
}

}  // namespace mil

using namespace mil;

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" int mil_sgemm_f32(const float* A, int64_t sAm, int64_t sAk, const int64_t* row_ids, const float* B, int64_t sBn, int64_t sBk,
                             const float* bias, float* C, int64_t ldc, float* pre_out, int64_t M, int64_t N, int64_t K, int act, int splitk,
                             void* ws, size_t ws_bytes, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(A && B && C, "mil_sgemm_f32: null operand");
  MIL_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "mil_sgemm_f32: bad sizes M=%lld N=%lld K=%lld ldc=%lld", (long long)M, (long long)N, (long long)K, (long long)ldc);
  MIL_CHECK_ARG(sAk == 1 || sAm == 1, "mil_sgemm_f32: A needs a unit stride");
  MIL_CHECK_ARG(sBk == 1 || sBn == 1, "mil_sgemm_f32: B needs a unit stride");
  MIL_CHECK_ARG(!(row_ids && sAk != 1), "mil_sgemm_f32: row_ids needs k-contiguous A");
  if (M == 0 || N == 0) return 0;
  if (splitk < 1) splitk = 1;
  const bool a_kc = (sAk == 1), b_kc = (sBk == 1);
  SgemmParams p;
  p.A = A; p.sAm = sAm; p.sAk = sAk; p.row_ids = row_ids; p.B = B; p.sBn = sBn; p.sBk = sBk;
  p.bias = bias; p.C = C; p.ldc = ldc; p.pre_out = pre_out; p.M = M; p.N = N; p.K = K; p.act = act;
  int64_t kps = (K + splitk - 1) / splitk;
  kps = (kps + SG_BK - 1) / SG_BK * SG_BK;
  if (kps <= 0) kps = SG_BK;
  const int splits = K > 0 ? (int)((K + kps - 1) / kps) : 1;
  p.k_per_split = kps;
  p.partial = nullptr;
  p.batch = 1; p.bA = p.bB = p.bC = 0;
  if (splits > 1) {
    MIL_CHECK_ARG(ws && ws_bytes >= (size_t)splits * M * N * sizeof(float), "mil_sgemm_f32: workspace too small for splitk");
    p.partial = (float*)ws;
  }
  p.vecA = ((uintptr_t)A % 16 == 0) && ((a_kc ? sAm : sAk) % 4 == 0);
  p.vecB = ((uintptr_t)B % 16 == 0) && ((b_kc ? sBn : sBk) % 4 == 0);
  dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM), (unsigned)splits);
  if (a_kc && b_kc) sgemm_kernel<true, true><<<grid, SG_THREADS, 0, stream>>>(p);
  else if (a_kc) sgemm_kernel<true, false><<<grid, SG_THREADS, 0, stream>>>(p);
  else if (b_kc) sgemm_kernel<false, true><<<grid, SG_THREADS, 0, stream>>>(p);
  else sgemm_kernel<false, false><<<grid, SG_THREADS, 0, stream>>>(p);
  MIL_LAUNCH_CHECK();
  if (splits > 1) {
    const int64_t tot = M * N;
    splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(p.partial, splits, M, N, bias, C, ldc, pre_out, act);
    MIL_LAUNCH_CHECK();
  }
  return 0;
}

// 64 x 64 x 16 tiles, 256 threads x (4 x 4): the 256 x 256 products of the pseudo-inverse iteration give 16 tiles per matrix x 8 heads = 128
// CTAs (the 128 x 128 tiling above gave 32 CTAs on 148 SMs: 52 us per product).
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) sgemm64_kernel(const SgemmParams p) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const float* __restrict__ Ap = p.A + (int64_t)blockIdx.z * p.bA;
  const float* __restrict__ Bp = p.B + (int64_t)blockIdx.z * p.bB;
  float* __restrict__ Cp = p.C + (int64_t)blockIdx.z * p.bC;
  const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // register prefetch of the next k tile while the current one is consumed from shared memory (the products are latency-bound:
  // one CTA per SM, 16 k tiles)
  float ra[4], rb[4];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      const int r = A_KC ? i >> 4 : i & 63, k = A_KC ? i & 15 : i >> 6;
      const int64_t m = m0 + r, kk = k0 + k;
      ra[q] = (m < p.M && kk < p.K) ? Ap[m * p.sAm + kk * p.sAk] : 0.f;
      const int r2 = B_KC ? i >> 4 : i & 63, k2 = B_KC ? i & 15 : i >> 6;
      const int64_t n = n0 + r2, kk2 = k0 + k2;
      rb[q] = (n < p.N && kk2 < p.K) ? Bp[n * p.sBn + kk2 * p.sBk] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      As[A_KC ? i & 15 : i >> 6][A_KC ? i >> 4 : i & 63] = ra[q];
      Bs[B_KC ? i & 15 : i >> 6][B_KC ? i >> 4 : i & 63] = rb[q];
    }
  };
  fetch(0);
  for (int64_t k0 = 0; k0 < p.K; k0 += 16) {
    stash();
    __syncthreads();
    if (k0 + 16 < p.K) fetch(k0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n < p.N) Cp[m * p.ldc + n] = acc[i][j];
    }
  }
}

extern "C" int mil_sgemm_batched_f32(const float* A, int64_t sAm, int64_t sAk, int64_t bA, const float* B, int64_t sBn, int64_t sBk, int64_t bB, float* C,
                                     int64_t ldc, int64_t bC, int64_t M, int64_t N, int64_t K, int batch, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N && batch >= 1 && batch <= 65535, "mil_sgemm_batched_f32: bad arguments");
  MIL_CHECK_ARG((sAk == 1 || sAm == 1) && (sBk == 1 || sBn == 1), "mil_sgemm_batched_f32: operands need a unit stride");
  const bool a_kc = (sAk == 1), b_kc = (sBk == 1);
  SgemmParams p;
  p.A = A; p.sAm = sAm; p.sAk = sAk; p.row_ids = nullptr; p.B = B; p.sBn = sBn; p.sBk = sBk;
  p.bias = nullptr; p.C = C; p.ldc = ldc; p.pre_out = nullptr; p.M = M; p.N = N; p.K = K; p.act = MIL_ACT_NONE;
  p.k_per_split = K; p.partial = nullptr;
  p.batch = batch < 2 ? 2 : batch;                       // batch == 1 still takes the batched addressing (blockIdx.z = 0)
  p.bA = bA; p.bB = bB; p.bC = bC;
  p.vecA = ((uintptr_t)A % 16 == 0) && ((a_kc ? sAm : sAk) % 4 == 0) && (bA % 4 == 0);
  p.vecB = ((uintptr_t)B % 16 == 0) && ((b_kc ? sBn : sBk) % 4 == 0) && (bB % 4 == 0);
  const int64_t big_tiles = ((N + SG_BN - 1) / SG_BN) * ((M + SG_BM - 1) / SG_BM) * batch;
  if (big_tiles < num_sms()) {                              // too few 128 x 128 tiles to fill the GPU: 64 x 64 tiles
    dim3 grid64((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)batch);
    if (a_kc && b_kc) sgemm64_kernel<true, true><<<grid64, 256, 0, stream>>>(p);
    else if (a_kc) sgemm64_kernel<true, false><<<grid64, 256, 0, stream>>>(p);
    else if (b_kc) sgemm64_kernel<false, true><<<grid64, 256, 0, stream>>>(p);
    else sgemm64_kernel<false, false><<<grid64, 256, 0, stream>>>(p);
    MIL_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM), (unsigned)batch);
  if (a_kc && b_kc) sgemm_kernel<true, true><<<grid, SG_THREADS, 0, stream>>>(p);
  else if (a_kc) sgemm_kernel<true, false><<<grid, SG_THREADS, 0, stream>>>(p);
  else if (b_kc) sgemm_kernel<false, true><<<grid, SG_THREADS, 0, stream>>>(p);
  else sgemm_kernel<false, false><<<grid, SG_THREADS, 0, stream>>>(p);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_act_bwd_f32(const float* g_y, const float* y_or_pre, int64_t n, int act, float* g_pre, mil_stream_t stream) {
  MIL_CHECK_ARG(g_y && y_or_pre && g_pre && n >= 0, "mil_act_bwd_f32: bad arguments");
  if (n == 0) return 0;
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_y, y_or_pre, n, act, g_pre);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_act_bwd_drop_f32(const float* g_y, const float* y_or_pre, int64_t rows, int ncols, int act, const mil_dropout_t* drop,
                                    float* g_pre, mil_stream_t stream) {
  MIL_CHECK_ARG(g_y && y_or_pre && g_pre && rows >= 0 && rows < (1ll << 31) && ncols > 0 && ncols % 32 == 0, "mil_act_bwd_drop_f32: bad arguments");
  MIL_CHECK_ARG(drop && drop->mode >= MIL_DROP_BITS && drop->mode <= MIL_DROP_PHILOX_DEV && drop->p > 0.f && drop->p < 1.f,
                "mil_act_bwd_drop_f32: needs a dropout description (mode 1, 2 or 3, 0 < p < 1)");
  MIL_CHECK_ARG(drop->mode == MIL_DROP_PHILOX || drop->keep_bits, "mil_act_bwd_drop_f32: modes 1 and 3 need keep_bits");
  MIL_CHECK_ARG((uintptr_t)g_y % 16 == 0 && (uintptr_t)y_or_pre % 16 == 0 && (uintptr_t)g_pre % 16 == 0, "mil_act_bwd_drop_f32: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  const int64_t words = rows * (ncols / 32);
  const uint32_t thresh = (uint32_t)lrint((1.0 - (double)drop->p) * 65536.0);
  act_bwd_drop_kernel<<<(unsigned)((words + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      g_y, y_or_pre, rows, ncols / 32, act, drop->mode, drop->keep_bits, thresh, 1.f / (1.f - drop->p), (uint32_t)drop->seed,
      (uint32_t)(drop->seed >> 32), (uint32_t)drop->offset, (uint32_t)(drop->offset >> 32), g_pre);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_colsum_f32(const float* A, int64_t M, int64_t N, float* out, void* ws, size_t ws_bytes, mil_stream_t stream) {
  MIL_CHECK_ARG(A && out && M >= 0 && N > 0, "mil_colsum_f32: bad arguments");
  int slices = (int)min((int64_t)64, max((int64_t)1, (M + 255) / 256));
  MIL_CHECK_ARG(ws && ws_bytes >= (size_t)slices * N * sizeof(float), "mil_colsum_f32: workspace needs %zu bytes", (size_t)slices * N * sizeof(float));
  const int64_t rps = (M + slices - 1) / slices;
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)slices);
  colsum_partial_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(A, M, N, rps, (float*)ws);
  MIL_LAUNCH_CHECK();
  colsum_final_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const float*)ws, slices, N, out);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_pool_num_partials(int64_t L) {
  const int64_t want = (L + 63) / 64;
  const int64_t cap = 2 * (int64_t)num_sms();
  return (int)max((int64_t)1, min(want, cap));
}

extern "C" int mil_pool_merge_f32(const float* part, int n_part, int H, float* stats, float* pooled, mil_stream_t stream) {
  MIL_CHECK_ARG(part && stats && pooled && n_part > 0 && n_part <= MERGE_MAXP && H > 0, "mil_pool_merge_f32: bad arguments (n_part <= %d)", MERGE_MAXP);
  pool_merge_kernel<<<(H + MERGE_COLS - 1) / MERGE_COLS, MERGE_COLS * MERGE_GROUPS, 0, (cudaStream_t)stream>>>(part, n_part, H, stats, pooled);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_shard_merge_cls_f32(const float* rec, int n_rec, int H, const float* Wcls, const float* bcls, int n_cls, float* stats,
                                       float* pooled, float* logits, mil_stream_t stream) {
  MIL_CHECK_ARG(rec && stats && pooled && n_rec > 0 && n_rec <= 1024 && H > 0 && H <= 8192, "mil_shard_merge_cls_f32: bad arguments (n_rec <= 1024, H <= 8192)");
  MIL_CHECK_ARG(!logits || (Wcls && n_cls > 0), "mil_shard_merge_cls_f32: logits needs Wcls and n_cls > 0");
  shard_merge_cls_kernel<<<1, 512, (size_t)(n_rec + H) * sizeof(float), (cudaStream_t)stream>>>(rec, n_rec, H, Wcls, bcls, n_cls, stats, pooled, logits);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_softmax_pool_fwd_f32(const float* s, int64_t s_stride, const float* h, int64_t L, int H, const uint8_t* keep, float* part,
                                        float* stats, float* pooled, float* attn_out, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(s && h && part && stats && pooled && L > 0, "mil_softmax_pool_fwd_f32: bad arguments");
  MIL_CHECK_ARG(H % 128 == 0 && H >= 128 && H <= 1024, "mil_softmax_pool_fwd_f32: H=%d must be a multiple of 128 in [128,1024]", H);
  MIL_CHECK_ARG((uintptr_t)h % 16 == 0, "mil_softmax_pool_fwd_f32: h must be 16-byte aligned");
  const int np = mil_pool_num_partials(L);
  const int64_t rpb = (L + np - 1) / np;
  switch (H / 128) {
#define CASE(NV) case NV: softmax_pool_partial_kernel<NV><<<np, POOL_THREADS, 0, stream>>>(s, s_stride, h, L, keep, rpb, part); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  MIL_LAUNCH_CHECK();
  pool_merge_kernel<<<(H + MERGE_COLS - 1) / MERGE_COLS, MERGE_COLS * MERGE_GROUPS, 0, stream>>>(part, np, H, stats, pooled);
  MIL_LAUNCH_CHECK();
  if (attn_out) {
    attn_norm_kernel<<<(unsigned)((L + 255) / 256), 256, 0, stream>>>(s, s_stride, L, keep, stats, attn_out);
    MIL_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int mil_softmax_pool_bwd_f32(const float* s, int64_t s_stride, const float* h, int64_t L, int H, const uint8_t* keep, const float* stats,
                                        const float* pooled, const float* g_p, float* g_s, int64_t gs_stride, float* g_h, int accumulate_gh,
                                        mil_stream_t stream) {
  MIL_CHECK_ARG(s && h && stats && pooled && g_p && g_s && L > 0 && H > 0, "mil_softmax_pool_bwd_f32: bad arguments");
  const int64_t blocks = min((int64_t)4 * num_sms(), (L + 7) / 8);
  softmax_pool_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(s, s_stride, h, L, H, keep, stats, pooled, g_p, g_s, gs_stride, g_h,
                                                                                accumulate_gh);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_cam_score_f32(const float* s, const float* t, int64_t L, int C, const float* stats, float bias0, float* score, mil_stream_t stream) {
  MIL_CHECK_ARG(s && t && stats && score && L > 0 && C > 0, "mil_cam_score_f32: bad arguments");
  cam_score_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s, t, L, C, stats, bias0, nullptr, score);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_cam_score_dev_f32(const float* s, const float* t, int64_t L, int C, const float* stats, const float* bias_dev, float* score,
                                     mil_stream_t stream) {
  MIL_CHECK_ARG(s && t && stats && score && bias_dev && L > 0 && C > 0, "mil_cam_score_dev_f32: bad arguments");
  cam_score_kernel<<<(unsigned)((L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s, t, L, C, stats, 0.f, bias_dev, score);
  MIL_LAUNCH_CHECK();
  return 0;
}

// Adam / AdamW step over a device table of parameter segments: ONE launch for the whole model (torch's foreach path issues ~12
// multi-tensor launches per step, its single-tensor path ~10 per parameter).  Same arithmetic as torch.optim.Adam(W)'s reference
// implementation (torch/optim/adam.py:_single_tensor_adam): L2 (Adam) or decoupled (AdamW) weight decay, lerp for the first moment,
// mul + addcmul for the second, denom = sqrt(v) / sqrt(bias_correction2) + eps, p -= lr / bias_correction1 * m / denom.
__global__ void adam_step_kernel(const mil_adam_seg_t* __restrict__ segs, float lr, float beta1, float beta2, float omb1, float omb2, float eps, float wd,
                                 int decoupled, float bc1, float bc2_sqrt, const float* __restrict__ step_dev) {
  const mil_adam_seg_t sg = segs[blockIdx.x];
  if (step_dev) {                                      // capturable: the step count lives on the device (CUDA-graph replays)
    const float t = step_dev[0];
    bc1 = 1.f - powf(beta1, t);
    bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  }
  const float step_size = lr / bc1;
  for (int64_t i = threadIdx.x; i < sg.n; i += blockDim.x) {
    float p = sg.p[i], g = sg.g[i], m = sg.m[i], v = sg.v[i];
    if (decoupled) p *= 1.f - lr * wd;
    else if (wd != 0.f) g = fmaf(wd, p, g);
    m = fmaf(omb1, g - m, m);                          // lerp_(grad, 1 - beta1); 1 - beta rounded from double like torch's python scalars
    v = fmaf(omb2, g * g, v * beta2);
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= step_size * (m / denom);
    sg.p[i] = p; sg.m[i] = m; sg.v[i] = v;
  }
}

extern "C" int mil_adam_step_f32(const mil_adam_seg_t* segs_dev, int n_seg, float lr, double beta1_d, double beta2_d, float eps, float weight_decay,
                                 int decoupled, float bias_correction1, float bias_correction2_sqrt, const float* step_dev, mil_stream_t stream) {
  const float beta1 = (float)beta1_d, beta2 = (float)beta2_d;
  MIL_CHECK_ARG(segs_dev && n_seg >= 0, "mil_adam_step_f32: bad arguments");
  MIL_CHECK_ARG(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && weight_decay >= 0.f,
                "mil_adam_step_f32: invalid hyper-parameters (same checks as torch.optim.Adam)");
  MIL_CHECK_ARG(step_dev || (bias_correction1 > 0.f && bias_correction2_sqrt > 0.f), "mil_adam_step_f32: bias corrections must be positive");
  if (n_seg == 0) return 0;
  adam_step_kernel<<<n_seg, 256, 0, (cudaStream_t)stream>>>(segs_dev, lr, beta1, beta2, (float)(1.0 - (double)beta1_d), (float)(1.0 - (double)beta2_d), eps,
                                                            weight_decay, decoupled, bias_correction1, bias_correction2_sqrt, step_dev);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_ema_update_f32(const mil_ema_seg_t* segs_dev, int n_seg, float mm, float one_minus_mm, mil_stream_t stream) {
  MIL_CHECK_ARG(n_seg >= 0 && (segs_dev || n_seg == 0), "mil_ema_update_f32: bad segment table");
  MIL_CHECK_ARG(mm >= 0.f && mm <= 1.f, "mil_ema_update_f32: momentum %f outside [0, 1]", (double)mm);
  if (n_seg == 0) return 0;
  ema_update_kernel<<<(unsigned)n_seg, 256, 0, (cudaStream_t)stream>>>(segs_dev, mm, one_minus_mm);
  MIL_LAUNCH_CHECK();
  return 0;
}
