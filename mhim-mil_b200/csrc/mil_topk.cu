// Masked hard-instance selection on the device (no host round trip).
//   mil_topk_f32          : select the k-th key (two bits per counting pass), compact the winners in index order, stable LSD radix sort by value
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

// Lanes of the warp that hold the same 8-bit digit as this lane (0 for a lane that is not `on`): nine ballots.  __match_any_sync on
// the digit was the cost of the first version of this file -- ~3 000 cycles per call when most lanes differ (tools/time_topk.py:
// every radix pass cost ~10 k cycles whatever k, the select 1.1 us per 1024 elements and pass).
__device__ __forceinline__ uint32_t match_digit8(int d, bool on) {
  uint32_t m = __ballot_sync(0xffffffffu, on);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1;
    const uint32_t v = __ballot_sync(0xffffffffu, bit);
    m &= bit ? v : ~v;
  }
  return on ? m : 0u;
}

// Inclusive scan over the 32 lanes of a warp.
__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Phases (one 1024-thread CTA; the sortable keys of the first `cap` instances are cached in shared memory):
//   1. T = the k-th largest key, two bits per counting pass (per-thread counters over 16-byte shared-memory loads, packed shuffle sums).
//   2. compact the winners (key > T, then the first r_ties keys == T) in ascending index order: every thread owns a contiguous
//      index range, one block-wide scan of the per-thread counts.
//   3. order the k winners by key, descending, index-ascending among equal keys: k <= 1024 by counting ranks, else a stable LSD radix sort
//      whose passes are skipped when all winners share the digit.
__global__ void __launch_bounds__(TK_THREADS) topk_kernel(const float* __restrict__ score, int64_t N, int64_t k, int largest, int64_t cap,
                                                          uint32_t* __restrict__ keyA, uint32_t* __restrict__ keyB,
                                                          int64_t* __restrict__ idxA, int64_t* __restrict__ idxB,
                                                          int64_t* __restrict__ idx_out, long long* __restrict__ stamps) {
  __shared__ int hist[256];
  __shared__ int64_t base[256];
  __shared__ int s_cnt[32][3];
  __shared__ unsigned long long s_pk[32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_wg[32], s_wt[32];
  __shared__ int s_skip;
  // dynamic shared memory: phases 1-2 cache the sortable keys of the first `cap` instances (4 B each: N = 50 000 fits), phase 3
  // re-uses the same bytes for the per-(digit, warp) counters of the stable scatter
  extern __shared__ __align__(16) uint32_t dyn[];
  uint32_t* skeys = dyn;
  int* wc = reinterpret_cast<int*>(dyn);                     // [256 digits][32 warps]
  __shared__ int64_t wsum[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  constexpr uint32_t FULL = 0xffffffffu;

  auto stamp = [&](int j) { if (stamps && tid == 0) stamps[j] = clock64(); };   // phase timeline (tools/time_topk.py); 8 stores per launch
  stamp(0);
  for (int64_t i = tid; i < min(N, cap); i += TK_THREADS) skeys[i] = sortable_key(score[i], largest);
  __syncthreads();
  stamp(1);
  auto key_at = [&](int64_t i) -> uint32_t { return i < cap ? skeys[i] : sortable_key(score[i], largest); };

  // ---- 1. T = key of the k-th element in descending key order, two bits per pass from the top: count the keys >= each of the three
  // candidates prefix | (1, 2, 3) << shift.  Plain per-thread counters over 16-byte shared-memory loads, shuffle-tree sums -- no warp
  // collective in the element loop.  (Measured, tools/time_topk.py: an 8-bit radix pass with match_any / nine ballots / REDUX per element
  // group costs ~1 200 cycles per 1024 elements whichever way the histogram is kept -- the vote / match / redux path of the SM serves the 32
  // warps one after the other; this loop costs ~110.)
  uint32_t prefix = 0;                                       // invariant: #{key >= prefix} >= k
  const int n_cached = (int)min(N, cap), n4 = n_cached >> 2;
  const bool packed = N < (1ll << 21);
  const uint4* sk4 = reinterpret_cast<const uint4*>(skeys);
  for (int shift = 30; shift >= 0; shift -= 2) {
    const uint32_t c1 = prefix | (1u << shift), c2 = prefix | (2u << shift), c3 = prefix | (3u << shift);
    int n1 = 0, n2 = 0, n3 = 0;
    for (int q = tid; q < n4; q += TK_THREADS) {
      const uint4 v = sk4[q];
      n1 += (v.x >= c1) + (v.y >= c1) + (v.z >= c1) + (v.w >= c1);
      n2 += (v.x >= c2) + (v.y >= c2) + (v.z >= c2) + (v.w >= c2);
      n3 += (v.x >= c3) + (v.y >= c3) + (v.z >= c3) + (v.w >= c3);
    }
    for (int64_t i = (int64_t)n4 * 4 + tid; i < N; i += TK_THREADS) {      // the last < 4 cached keys and whatever did not fit the cache
      const uint32_t key = key_at(i);
      n1 += key >= c1; n2 += key >= c2; n3 += key >= c3;
    }
    // warp sums, then block sums by warp 0 alone: a shuffle costs the SM one issue slot per WARP (32 warps x 30 shuffles per pass were
    // ~1 000 of the ~1 700 cycles a pass cost), so the three counters travel packed in one 64-bit word (21 bits each; N < 2^21)
    if (packed) {
      unsigned long long pk = (unsigned long long)n1 | ((unsigned long long)n2 << 21) | ((unsigned long long)n3 << 42);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pk += __shfl_xor_sync(FULL, pk, o);
      if (lane == 0) s_pk[warp] = pk;
      __syncthreads();
      if (warp == 0) {
        unsigned long long tp = s_pk[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tp += __shfl_xor_sync(FULL, tp, o);
        const int64_t t1 = (int64_t)(tp & 0x1FFFFFull), t2 = (int64_t)((tp >> 21) & 0x1FFFFFull), t3 = (int64_t)(tp >> 42);
        if (lane == 0) s_prefix = t3 >= k ? c3 : (t2 >= k ? c2 : (t1 >= k ? c1 : prefix));
      }
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        n1 += __shfl_xor_sync(FULL, n1, o); n2 += __shfl_xor_sync(FULL, n2, o); n3 += __shfl_xor_sync(FULL, n3, o);
      }
      if (lane == 0) { s_cnt[warp][0] = n1; s_cnt[warp][1] = n2; s_cnt[warp][2] = n3; }
      __syncthreads();
      if (warp == 0) {
        int64_t t1 = s_cnt[lane][0], t2 = s_cnt[lane][1], t3 = s_cnt[lane][2];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          t1 += __shfl_xor_sync(FULL, t1, o); t2 += __shfl_xor_sync(FULL, t2, o); t3 += __shfl_xor_sync(FULL, t3, o);
        }
        if (lane == 0) s_prefix = t3 >= k ? c3 : (t2 >= k ? c2 : (t1 >= k ? c1 : prefix));
      }
    }
    __syncthreads();
    prefix = s_prefix;
  }
  const uint32_t T = prefix;
  stamp(2);

  // ---- 2. compact winners in ascending index order (ties at T: lowest index first); r_ties = how many keys == T to take
  const int64_t per = (N + TK_THREADS - 1) / TK_THREADS;
  const int64_t i0 = min(N, (int64_t)tid * per), i1 = min(N, i0 + per);
  int cg = 0, ct = 0;
  for (int64_t i = i0; i < i1; ++i) {
    const uint32_t key = key_at(i);
    cg += key > T; ct += key == T;
  }
  const int sg = warp_inclusive_scan(cg, lane), st = warp_inclusive_scan(ct, lane);
  if (lane == 31) { s_wg[warp] = sg; s_wt[warp] = st; }
  __syncthreads();
  const int vg = warp_inclusive_scan(s_wg[lane], lane), vt = warp_inclusive_scan(s_wt[lane], lane);   // every warp scans the 32 warp totals
  const int64_t n_gt = __shfl_sync(FULL, vg, 31);
  const int64_t r_ties = k - n_gt;                           // 1 <= r_ties <= #{key == T}
  int64_t gt_before = (warp ? __shfl_sync(FULL, vg, warp - 1) : 0) + sg - cg;
  int64_t tie_before = (warp ? __shfl_sync(FULL, vt, warp - 1) : 0) + st - ct;
  for (int64_t i = i0; i < i1; ++i) {
    const uint32_t key = key_at(i);
    if (key > T) {
      const int64_t pos = gt_before + min(tie_before, r_ties);
      keyA[pos] = key; idxA[pos] = i;
      ++gt_before;
    } else if (key == T) {
      if (tie_before < r_ties) {
        const int64_t pos = gt_before + tie_before;
        keyA[pos] = key; idxA[pos] = i;
      }
      ++tie_before;
    }
  }
  __syncthreads();                       // the key cache is dead from here on: its bytes become wc; keyA / idxA are complete
  stamp(3);

  // ---- 3a. k <= 1024: rank sort.  (key, position in the index-ordered winner list) is a total order without ties, so the output slot of
  // a winner is the number of winners that beat it: every thread counts that for its own winner over the list in shared memory.
  if (k <= TK_THREADS) {
    unsigned long long* sv = reinterpret_cast<unsigned long long*>(dyn);
    unsigned long long mine = 0;
    if (tid < k) { mine = ((unsigned long long)keyA[tid] << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)tid); sv[tid] = mine; }
    __syncthreads();
    if (tid < k) {
      int r = 0;
      const int kk = (int)k;
#pragma unroll 4
      for (int j = 0; j < kk; ++j) r += sv[j] > mine;
      idx_out[r] = idxA[tid];
    }
    stamp(8);
    return;
  }
  // ---- 3b. stable LSD radix sort of the k winners by key, descending
  uint32_t* kin = keyA; uint32_t* kout = keyB;
  int64_t* iin = idxA; int64_t* iout = idxB;
  for (int shift = 0; shift < 32; shift += 8) {
    if (tid < 256) hist[tid] = 0;
    if (tid == 0) s_skip = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < k; c0 += TK_THREADS) {
      const int64_t i = c0 + tid;
      const int d = i < k ? (int)((kin[i] >> shift) & 255u) : 0;
      const uint32_t peers = match_digit8(d, i < k);
      if (i < k && (peers & lt_mask) == 0) atomicAdd(&hist[d], __popc(peers));
    }
    __syncthreads();
    if (tid < 256 && hist[tid] == k) s_skip = 1;             // every winner has this digit: the pass would not move anything
    {
      const int64_t suf = suffix_sum256(hist, wsum);         // (two barriers inside: s_skip is visible afterwards)
      if (tid < 256) base[255 - tid] = suf - hist[255 - tid];  // descending order: bucket d starts after all larger digits
    }
    __syncthreads();
    const int skip = s_skip;                                 // block-uniform
    __syncthreads();                                         // everybody has read it before the next pass clears it
    if (skip) { stamp(4 + shift / 8); continue; }
    for (int64_t c0 = 0; c0 < k; c0 += TK_THREADS) {
      for (int j = tid; j < 32 * 256; j += TK_THREADS) wc[j] = 0;
      __syncthreads();
      const int64_t i = c0 + tid;
      const bool on = i < k;
      uint32_t key = 0;
      int64_t id = 0;
      int d = 0;
      if (on) { key = kin[i]; id = iin[i]; d = (key >> shift) & 255u; }
      const uint32_t peers = match_digit8(d, on);
      const int rank = __popc(peers & lt_mask);
      if (on && rank == 0) wc[d * 32 + warp] = __popc(peers);
      __syncthreads();
      // exclusive prefix over the warps for every digit: warp w takes digits 8 w .. 8 w + 7, lane = source warp (shuffle scan)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int dd = warp * 8 + q;
        const int v = wc[dd * 32 + lane];
        const int inc = warp_inclusive_scan(v, lane);
        wc[dd * 32 + lane] = inc - v;
        if (lane == 31) hist[dd] = inc;                      // hist[] reused as the tile total of the digit
      }
      __syncthreads();
      if (on) {
        const int64_t pos = base[d] + wc[d * 32 + warp] + rank;
        kout[pos] = key;
        iout[pos] = id;
      }
      __syncthreads();
      if (tid < 256) base[tid] += hist[tid];
      __syncthreads();
    }
    uint32_t* tk = kin; kin = kout; kout = tk;
    int64_t* ti = iin; iin = iout; iout = ti;
    stamp(4 + shift / 8);
  }
  for (int64_t i = tid; i < k; i += TK_THREADS) idx_out[i] = iin[i];
  stamp(8);
}

__global__ void __launch_bounds__(TK_THREADS) mask_from_indices_kernel(const int64_t* __restrict__ idx, int64_t k, int64_t N, int64_t cap,
                                                                       int64_t* __restrict__ mask_ids, uint8_t* __restrict__ keep,
                                                                       int64_t* __restrict__ len_keep_out) {
  __shared__ int cnt[32];
  extern __shared__ __align__(16) uint8_t sflag[];            // keep flags of the first `cap` instances (1 B each)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int64_t i = tid; i < min(N, cap); i += TK_THREADS) sflag[i] = 1;
  for (int64_t i = cap + tid; i < N; i += TK_THREADS) keep[i] = 1;
  __syncthreads();
  for (int64_t j = tid; j < k; j += TK_THREADS) {
    const int64_t v = idx[j];
    if (v >= 0 && v < N) { if (v < cap) sflag[v] = 0; else keep[v] = 0; }
  }
  __syncthreads();
  // kept instances in ascending order: every thread owns a contiguous index range, one block-wide scan of the per-thread counts (a
  // ballot + prefix per 1024-element chunk cost two barriers and a 32-step shared-memory walk per chunk)
  const int64_t per = (N + TK_THREADS - 1) / TK_THREADS;
  const int64_t i0 = min(N, (int64_t)tid * per), i1 = min(N, i0 + per);
  int c = 0;
  for (int64_t i = i0; i < i1; ++i) c += (i < cap ? sflag[i] : keep[i]) ? 1 : 0;
  const int sc = warp_inclusive_scan(c, lane);
  if (lane == 31) cnt[warp] = sc;
  __syncthreads();
  const int vw = warp_inclusive_scan(cnt[lane], lane);      // every warp scans the 32 warp totals
  const int64_t run = __shfl_sync(0xffffffffu, vw, 31);
  int64_t pos = (warp ? __shfl_sync(0xffffffffu, vw, warp - 1) : 0) + sc - c;
  for (int64_t i = i0; i < i1; ++i) {
    const bool kp = (i < cap ? sflag[i] : keep[i]) != 0;
    if (kp) mask_ids[pos++] = i;
    if (i < cap) keep[i] = kp ? 1 : 0;
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
  long long* stamps = (long long*)(((uintptr_t)(keyB + N) + 7) & ~(uintptr_t)7);      // 9 clock stamps in the workspace's 256 spare bytes
  topk_kernel<<<1, TK_THREADS, dyn, (cudaStream_t)stream>>>(score, N, k, largest, (int64_t)(dyn / sizeof(uint32_t)), keyA, keyB, idxA, idxB, idx_out, stamps);
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
