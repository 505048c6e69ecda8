// Masked hard-instance selection on the device (no host round trip).
//   mil_topk_f32          : radix-select the k-th key, compact the winners in index order, stable LSD radix sort by value
//   mil_mask_from_indices : complement + concatenation -> mask_ids = [kept ascending || masked], keep flags, len_keep
// One 1024-thread CTA each: the score vector (4 B/instance; 200 KB at N=50k) is L2-resident right after the teacher
// pass, the work is a handful of passes over it, and a single CTA needs no grid-wide synchronisation.
// Total order used everywhere: value (descending for largest), then index ascending -> deterministic, no ties.
#include "mil_common.cuh"

namespace mil {

constexpr int TK_THREADS = 1024;

__device__ __forceinline__ uint32_t sortable_key(float f, int largest) {
  uint32_t u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // monotone: bigger float -> bigger key
  return largest ? u : ~u;                          // "bigger key" always means "selected first"
}

// Exclusive prefix over the per-warp counts held in sh[0..31]; returns (prefix for this warp, total).
__device__ __forceinline__ void warp_prefix(const int* sh, int warp, int& prefix, int& total) {
  int p = 0, t = 0;
#pragma unroll
  for (int w = 0; w < TK_THREADS / 32; ++w) {
    const int c = sh[w];
    if (w < warp) p += c;
    t += c;
  }
  prefix = p;
  total = t;
}

// Inclusive suffix sums over 256 histogram bins: suf[d] = sum_{d' >= d} hist[d'] for thread d < 256 (all threads must call).
// Replaces a 256-step single-thread scan (a dependent chain of shared-memory loads: ~10k cycles per radix pass).
__device__ __forceinline__ int64_t suffix_sum256(const int* hist, int64_t* wsum /*[8]*/) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t v = 0;
  if (tid < 256) {
    v = hist[255 - tid];                                   // reversed: prefix over the reversed array = suffix
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) wsum[warp] = v;
  }
  __syncthreads();
  if (tid < 256) {
    for (int w = 0; w < warp; ++w) v += wsum[w];
  }
  __syncthreads();
  return v;                                                // thread tid < 256 holds suf[255 - tid]
}

__global__ void __launch_bounds__(TK_THREADS) topk_kernel(const float* __restrict__ score, int64_t N, int64_t k, int largest, int64_t cap,
                                                          uint32_t* __restrict__ keyA, uint32_t* __restrict__ keyB,
                                                          int64_t* __restrict__ idxA, int64_t* __restrict__ idxB,
                                                          int64_t* __restrict__ idx_out) {
  __shared__ int hist[256];
  __shared__ int64_t base[256];
  __shared__ int cnt_gt[32], cnt_tie[32];
  // dynamic shared memory: phases 1-2 cache the sortable keys of the first `cap` instances (4 B each: N = 50 000 fits), phase 3
  // re-uses the same bytes for the per-warp digit counters of the stable scatter
  extern __shared__ __align__(16) uint32_t dyn[];
  uint32_t* skeys = dyn;
  int (*wcnt)[256] = reinterpret_cast<int (*)[256]>(dyn);
  __shared__ uint32_t s_prefix;
  __shared__ int64_t s_need;
  __shared__ int64_t wsum[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int64_t i = tid; i < min(N, cap); i += TK_THREADS) skeys[i] = sortable_key(score[i], largest);
  __syncthreads();
  auto key_at = [&](int64_t i) -> uint32_t { return i < cap ? skeys[i] : sortable_key(score[i], largest); };

  // ---- 1. radix select: T = key of the k-th element in descending key order; need = how many keys == T to take
  uint32_t prefix = 0, mask = 0;
  int64_t need = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < N; c0 += TK_THREADS) {          // warp-uniform trip count (match_any needs the full warp)
      const int64_t i = c0 + tid;
      int d = 256 + lane;                                      // lanes without a candidate never match anybody
      if (i < N) {
        const uint32_t key = key_at(i);
        if ((key & mask) == prefix) d = (int)((key >> shift) & 255u);
      }
      // the CAM scores sit on a few hundred values around 0.5, i.e. in one or two bins: aggregate per warp before the atomic
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (d < 256 && (peers & lt_mask) == 0) atomicAdd(&hist[d], __popc(peers));
    }
    __syncthreads();
    {
      // the selected digit d is the largest one whose suffix count reaches `need`: suf[d] >= need > suf[d+1]
      const int64_t suf = suffix_sum256(hist, wsum);         // thread t < 256: suf of digit 255 - t
      if (tid < 256) {
        const int d = 255 - tid;
        const int64_t above = suf - hist[d];                 // keys with a larger digit
        if (suf >= need && above < need) { s_need = need - above; s_prefix = prefix | ((uint32_t)d << shift); }
      }
    }
    __syncthreads();
    need = s_need;
    prefix = s_prefix;
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t T = prefix;
  const int64_t r_ties = need;   // 1 <= r_ties <= #{key == T}

  // ---- 2. compact winners in ascending index order (ties at T: lowest index first)
  int64_t run_gt = 0, run_tie = 0;
  for (int64_t c0 = 0; c0 < N; c0 += TK_THREADS) {
    const int64_t i = c0 + tid;
    uint32_t key = 0;
    bool gt = false, tie = false;
    if (i < N) {
      key = key_at(i);
      gt = key > T;
      tie = key == T;
    }
    const uint32_t bg = __ballot_sync(0xffffffffu, gt), bt = __ballot_sync(0xffffffffu, tie);
    if (lane == 0) { cnt_gt[warp] = __popc(bg); cnt_tie[warp] = __popc(bt); }
    __syncthreads();
    int pg, tg, pt, tt;
    warp_prefix(cnt_gt, warp, pg, tg);
    warp_prefix(cnt_tie, warp, pt, tt);
    const int64_t gt_before = run_gt + pg + __popc(bg & lt_mask);
    const int64_t tie_before = run_tie + pt + __popc(bt & lt_mask);
    if (gt || (tie && tie_before < r_ties)) {
      const int64_t pos = gt_before + min(tie_before, r_ties);
      keyA[pos] = key;
      idxA[pos] = i;
    }
    run_gt += tg;
    run_tie += tt;
    __syncthreads();
  }

  __syncthreads();                       // the key cache is dead from here on: its bytes become wcnt
  // ---- 3. stable LSD radix sort of the k winners by key, descending (stability keeps index-ascending among equals)
  uint32_t* kin = keyA; uint32_t* kout = keyB;
  int64_t* iin = idxA; int64_t* iout = idxB;
  for (int shift = 0; shift < 32; shift += 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < k; c0 += TK_THREADS) {
      const int64_t i = c0 + tid;
      const int d = i < k ? (int)((kin[i] >> shift) & 255u) : 256 + lane;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (d < 256 && (peers & lt_mask) == 0) atomicAdd(&hist[d], __popc(peers));
    }
    __syncthreads();
    {
      const int64_t suf = suffix_sum256(hist, wsum);
      if (tid < 256) base[255 - tid] = suf - hist[255 - tid];  // descending order: bucket d starts after all larger digits
    }
    __syncthreads();
    for (int64_t c0 = 0; c0 < k; c0 += TK_THREADS) {
      for (int j = tid; j < 32 * 256; j += TK_THREADS) (&wcnt[0][0])[j] = 0;
      __syncthreads();
      const int64_t i = c0 + tid;
      const bool on = i < k;
      uint32_t key = 0;
      int64_t id = 0;
      int d = 256 + lane;                 // inactive lanes never match anybody
      if (on) { key = kin[i]; id = iin[i]; d = (key >> shift) & 255u; }
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const int rank = __popc(peers & lt_mask);
      if (on && rank == 0) wcnt[warp][d] = __popc(peers);
      __syncthreads();
      if (tid < 256) {                    // exclusive prefix over warps for digit tid; hist[] reused as the tile total
        int run = 0;
        for (int w = 0; w < 32; ++w) { const int c = wcnt[w][tid]; wcnt[w][tid] = run; run += c; }
        hist[tid] = run;
      }
      __syncthreads();
      if (on) {
        const int64_t pos = base[d] + wcnt[warp][d] + rank;
        kout[pos] = key;
        iout[pos] = id;
      }
      __syncthreads();
      if (tid < 256) base[tid] += hist[tid];
      __syncthreads();
    }
    uint32_t* tk = kin; kin = kout; kout = tk;
    int64_t* ti = iin; iin = iout; iout = ti;
  }
  for (int64_t i = tid; i < k; i += TK_THREADS) idx_out[i] = iin[i];
}

__global__ void __launch_bounds__(TK_THREADS) mask_from_indices_kernel(const int64_t* __restrict__ idx, int64_t k, int64_t N, int64_t cap,
                                                                       int64_t* __restrict__ mask_ids, uint8_t* __restrict__ keep,
                                                                       int64_t* __restrict__ len_keep_out) {
  __shared__ int cnt[32];
  extern __shared__ __align__(16) uint8_t sflag[];            // keep flags of the first `cap` instances (1 B each)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int64_t i = tid; i < min(N, cap); i += TK_THREADS) sflag[i] = 1;
  for (int64_t i = cap + tid; i < N; i += TK_THREADS) keep[i] = 1;
  __syncthreads();
  for (int64_t j = tid; j < k; j += TK_THREADS) {
    const int64_t v = idx[j];
    if (v >= 0 && v < N) { if (v < cap) sflag[v] = 0; else keep[v] = 0; }
  }
  __syncthreads();
  int64_t run = 0;
  for (int64_t c0 = 0; c0 < N; c0 += TK_THREADS) {
    const int64_t i = c0 + tid;
    const bool kp = i < N && (i < cap ? sflag[i] : keep[i]);
    const uint32_t b = __ballot_sync(0xffffffffu, kp);
    if (lane == 0) cnt[warp] = __popc(b);
    __syncthreads();
    int p, t;
    warp_prefix(cnt, warp, p, t);
    if (kp) mask_ids[run + p + __popc(b & lt_mask)] = i;
    if (i < cap && i < N) keep[i] = kp ? 1 : 0;
    run += t;
    __syncthreads();
  }
  for (int64_t j = tid; j < k; j += TK_THREADS)
    if (run + j < N) mask_ids[run + j] = idx[j];
  if (tid == 0) *len_keep_out = run;
}

}  // namespace mil

using namespace mil;

// idx[c] = argmax_m A[m, c] (lowest row index among equal maxima), val[c] = the maximum: one CTA per column.
__global__ void __launch_bounds__(1024) col_argmax_kernel(const float* __restrict__ A, int64_t M, int C, int64_t* __restrict__ idx, float* __restrict__ val) {
  __shared__ float sv[32];
  __shared__ long long si[32];
  const int c = blockIdx.x;
  float best = -INFINITY;
  long long bi = 0x7fffffffffffffffll;
  for (int64_t m = threadIdx.x; m < M; m += blockDim.x) {
    const float v = A[m * C + c];
    if (v > best || (v == best && m < bi)) { best = v; bi = m; }
  }
  auto better = [](float v, long long i, float bv, long long bj) { return v > bv || (v == bv && i < bj); };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffffffffffffll;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
      if (bi == 0x7fffffffffffffffll) bi = 0;        // all NaN / -inf: row 0, like an empty comparison
      idx[c] = bi;
      if (val) val[c] = best;
    }
  }
}

extern "C" int mil_col_argmax_f32(const float* A, int64_t M, int C, int64_t* idx_out, float* val_out, mil_stream_t stream) {
  MIL_CHECK_ARG(A && idx_out && M > 0 && C > 0 && C <= 65535, "mil_col_argmax_f32: bad arguments");
  col_argmax_kernel<<<C, 1024, 0, (cudaStream_t)stream>>>(A, M, C, idx_out, val_out);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mil_topk_workspace_bytes(int64_t N) { return (size_t)N * 24 + 256; }

extern "C" int mil_topk_f32(const float* score, int64_t N, int64_t k, int largest, int64_t* idx_out, void* ws, size_t ws_bytes, mil_stream_t stream) {
  MIL_CHECK_ARG(score && idx_out && N > 0, "mil_topk_f32: bad arguments");
  MIL_CHECK_ARG(k >= 0 && k <= N, "mil_topk_f32: k=%lld out of range for N=%lld", (long long)k, (long long)N);
  MIL_CHECK_ARG(ws && ws_bytes >= mil_topk_workspace_bytes(N), "mil_topk_f32: workspace needs %zu bytes", mil_topk_workspace_bytes(N));
  if (k == 0) return 0;
  char* w = (char*)(((uintptr_t)ws + 15) & ~(uintptr_t)15);
  int64_t* idxA = (int64_t*)w;
  int64_t* idxB = idxA + N;
  uint32_t* keyA = (uint32_t*)(idxB + N);
  uint32_t* keyB = keyA + N;
  // key cache: as many instances as fit next to the static buffers (>= the 32 KB the sort phase needs)
  const size_t dyn_max = 200 * 1024, dyn_min = 32 * 256 * sizeof(int);
  size_t dyn = (size_t)N * sizeof(uint32_t);
  dyn = dyn < dyn_min ? dyn_min : (dyn > dyn_max ? dyn_max : dyn);
  static bool attr_set_dev[64] = {false};        // the attribute is per device (context): one flag per device ordinal
  int attr_dev = 0;
  cudaGetDevice(&attr_dev);
  bool& attr_set = attr_set_dev[attr_dev & 63];
  if (!attr_set) {
    MIL_CUDA(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max));
    attr_set = true;
  }
  topk_kernel<<<1, TK_THREADS, dyn, (cudaStream_t)stream>>>(score, N, k, largest, (int64_t)(dyn / sizeof(uint32_t)), keyA, keyB, idxA, idxB, idx_out);
  MIL_LAUNCH_CHECK();
  return 0;
}

extern "C" int mil_mask_from_indices(const int64_t* idx, int64_t k, int64_t N, int64_t* mask_ids, uint8_t* keep, int64_t* len_keep_out, void* ws,
                                     size_t ws_bytes, mil_stream_t stream) {
  (void)ws; (void)ws_bytes;
  MIL_CHECK_ARG(mask_ids && keep && len_keep_out && N > 0 && k >= 0 && k <= N && (idx || k == 0), "mil_mask_from_indices: bad arguments");
  const size_t dyn_max = 200 * 1024;
  const size_t dyn = (size_t)N < dyn_max ? (((size_t)N + 15) & ~(size_t)15) : dyn_max;
  static bool attr_set_dev[64] = {false};        // the attribute is per device (context): one flag per device ordinal
  int attr_dev = 0;
  cudaGetDevice(&attr_dev);
  bool& attr_set = attr_set_dev[attr_dev & 63];
  if (!attr_set) {
    MIL_CUDA(cudaFuncSetAttribute(mask_from_indices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max));
    attr_set = true;
  }
  mask_from_indices_kernel<<<1, TK_THREADS, dyn, (cudaStream_t)stream>>>(idx, k, N, (int64_t)dyn, mask_ids, keep, len_keep_out);
  MIL_LAUNCH_CHECK();
  return 0;
}
