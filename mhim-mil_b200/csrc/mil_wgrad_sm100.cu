// Weight gradient on the tensor cores (sm_100a): dW[n, k] = sum_m G[m, n] X[m, k] -- the contraction over the INSTANCES that the
// autograd of every N-row Linear on the path needs (abmil.py:213, mhim.py:193/335, dsmil.py:62-70, nystrom_attention.py:52-57), and
// the expensive reduction of the streaming backward (dW1 = sum_n g_pre^T x).  Round 1 ran it as an fp32 FFMA split-K GEMM
// (~50 % of the GPU time of a training step).
//
// Both operands are contracted over their ROW index, i.e. both are "MN-major" for the MMA.  Instead of MN-major descriptors the
// converter warps transpose while they split: a thread owns one operand row (one output row n of G^T, or one output column k of X^T),
// reads its 32 values of the 32-instance slab from the fp32 staging tile (bank-conflict-free: consecutive threads read consecutive
// words of one staged row) and writes them as ONE 64-byte row of the K-major SWIZZLE_64B operand tile -- the same tile layout,
// descriptors and instruction descriptor as the forward kernels.
//
//   grid = (tiles_n x tiles_k) output tiles of 128 x 256  x  `splits` slices of the instance range (multiples of 32 rows)
//   per CTA, per 32-instance step:
//     warp 0      : TMA  G slab [32 x 128] + X slab [32 x 256] fp32 (no swizzle) -> staging ring (2 x 48 KB), L2 prefetch 6 steps ahead
//     warps 4-15  : converters: 128 + 256 operand rows, one per thread: bf16 hi + lo, K-major SW64 -> operand ring (2 x 48 KB)
//     warp 1      : tcgen05.mma (cta_group::1, M = 128, N = 256, K = 16) x 2 k16 x 3 products -> 256 TMEM columns
//   tail          : warps 4-7 read the accumulator (tcgen05.ld) and store the 128 x 256 partial; bias gradient partials (column sums of
//                   G, accumulated by the A-row converter threads in registers) are written by the k-tile-0 CTAs.
//   a second small kernel sums the `splits` partials in a fixed order (deterministic) into dW / db.
#include "mil_umma.cuh"

namespace mil {
namespace wgrad {

constexpr int TM = 128;                       // output rows per tile  (n: columns of G)
constexpr int TN = 256;                       // output cols per tile  (k: columns of X)
constexpr int BKM = 32;                       // instances per step (two UMMA K = 16 steps)
constexpr int NSLOT = 2;                      // fp32 staging slots
constexpr int NST = 2;                        // operand stages
constexpr int G_SLAB = BKM * TM * 4;          // 16384
constexpr int X_SLAB = BKM * TN * 4;          // 32768
constexpr int SLOT_BYTES = G_SLAB + X_SLAB;   // 49152
constexpr int A_OP = TM * BKM * 2;            // 8192   one 16-bit A tile [128 x 32]
constexpr int B_OP = TN * BKM * 2;            // 16384  one 16-bit B tile [256 x 32]
constexpr int STAGE_BYTES = 2 * (A_OP + B_OP);   // hi + lo: 49152
constexpr int NUM_THREADS = 512;
constexpr int CONV_WARP0 = 4, N_CONV_WARPS = 12;
constexpr int PF = 6;                         // L2 prefetch distance in steps

struct Params {
  int64_t M;            // instances (rows of G and X)
  int Nn, K;            // dW is [Nn, K]
  int tiles_n, tiles_k, splits;
  int64_t rows_per_split;   // multiple of 32
  float* out;           // splits == 1: dW itself; else partials [splits][Nn][K]
  float* db_part;       // nullable: [splits][Nn] column sums of G
  int* err;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sStage = smem;                                    // NSLOT x [G slab | X slab]
  uint8_t* sOp = sStage + NSLOT * SLOT_BYTES;                // NST x [A hi | A lo | B hi | B lo]
  uint8_t* sMisc = sOp + NST * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMisc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sMisc + 128);

  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_XFULL = 0, B_XEMPTY = B_XFULL + NSLOT, B_FULL = B_XEMPTY + NSLOT, B_EMPTY = B_FULL + NST, B_ACCFULL = B_EMPTY + NST;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile / slice of this CTA: n tiles fastest, so that the CTAs sharing an X slab (same slice, same k tile) are co-scheduled
  const int tn = blockIdx.x % p.tiles_n, tk = (blockIdx.x / p.tiles_n) % p.tiles_k, sp = blockIdx.x / (p.tiles_n * p.tiles_k);
  const int64_t m_begin = (int64_t)sp * p.rows_per_split;
  int64_t m_end = m_begin + p.rows_per_split;
  if (m_end > p.M) m_end = p.M;
  const int steps = m_end > m_begin ? (int)((m_end - m_begin + BKM - 1) / BKM) : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(BAR(B_XFULL + i), 1); mbar_init(BAR(B_XEMPTY + i), N_CONV_WARPS); }
    for (int i = 0; i < NST; ++i) { mbar_init(BAR(B_FULL + i), N_CONV_WARPS); mbar_init(BAR(B_EMPTY + i), 1); }
    mbar_init(BAR(B_ACCFULL), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapG); tma_prefetch_desc(&mapX); }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: HBM / L2 -> fp32 staging =====================
    if (lane == 0) {
      for (int i = 0; i < PF && i < steps; ++i) {
        tma_prefetch_2d(&mapG, tn * TM, (int)(m_begin + (int64_t)i * BKM));
        tma_prefetch_2d(&mapX, tk * TN, (int)(m_begin + (int64_t)i * BKM));
      }
      for (int it = 0; it < steps; ++it) {
        if (it + PF < steps) {
          tma_prefetch_2d(&mapG, tn * TM, (int)(m_begin + (int64_t)(it + PF) * BKM));
          tma_prefetch_2d(&mapX, tk * TN, (int)(m_begin + (int64_t)(it + PF) * BKM));
        }
        const uint32_t s = (uint32_t)it % NSLOT, ph = ((uint32_t)it / NSLOT) & 1u;
        mbar_wait(BAR(B_XEMPTY + s), ph ^ 1u, p.err, 1);
        mbar_expect_tx(BAR(B_XFULL + s), SLOT_BYTES);
        const uint32_t dst = smem_u32(sStage + s * SLOT_BYTES);
        const int m0 = (int)(m_begin + (int64_t)it * BKM);
        tma_load_2d(dst, &mapG, BAR(B_XFULL + s), tn * TM, m0);
        tma_load_2d(dst + G_SLAB, &mapX, BAR(B_XFULL + s), tk * TN, m0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane) =====================
    const uint32_t idesc = make_idesc(0, TN, TM);
    const uint32_t op0 = smem_u32(sOp);
    for (int it = 0; it < steps; ++it) {
      const uint32_t s = (uint32_t)it % NST, ph = ((uint32_t)it / NST) & 1u;
      mbar_wait(BAR(B_FULL + s), ph, p.err, 2);
      tc_fence_after();
      const uint32_t a0 = op0 + s * STAGE_BYTES;
      const uint64_t ah0 = make_desc_sw64(a0), al0 = make_desc_sw64(a0 + A_OP), bh0 = make_desc_sw64(a0 + 2 * A_OP), bl0 = make_desc_sw64(a0 + 2 * A_OP + B_OP);
      if (elect_one()) {
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16) {
          const uint64_t o = (uint64_t)(k16 * 2);             // +32 bytes per K = 16 step, in 16-byte units
          umma_f16(tmem, ah0 + o, bh0 + o, idesc, (it | k16) ? 1u : 0u);
          umma_f16(tmem, al0 + o, bh0 + o, idesc, 1u);
          umma_f16(tmem, ah0 + o, bl0 + o, idesc, 1u);
        }
        umma_commit(BAR(B_EMPTY + s));
        if (it == steps - 1) umma_commit(BAR(B_ACCFULL));
      }
      __syncwarp();
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== converters (transpose + hi/lo split), then the accumulator read-out =====================
    const int cw = warp - CONV_WARP0;                          // 0..3: A rows (n), 4..11: B rows (k)
    const bool isA = cw < 4;
    const int row = isA ? cw * 32 + lane : (cw - 4) * 32 + lane;     // operand row = output row n (A) / output column k (B)
    const uint32_t src_off = isA ? (uint32_t)row * 4u : (uint32_t)G_SLAB + (uint32_t)row * 4u;
    const uint32_t src_stride = isA ? (uint32_t)TM * 4u : (uint32_t)TN * 4u;
    const uint32_t dst_off = isA ? 0u : 2u * (uint32_t)A_OP;
    const uint32_t lo_off = isA ? (uint32_t)A_OP : (uint32_t)B_OP;
    float colsum = 0.f;
    for (int it = 0; it < steps; ++it) {
      const uint32_t xs = (uint32_t)it % NSLOT, xph = ((uint32_t)it / NSLOT) & 1u;
      const uint32_t s = (uint32_t)it % NST, ph = ((uint32_t)it / NST) & 1u;
      mbar_wait(BAR(B_XFULL + xs), xph, p.err, 3);
      float x[32];
      const uint32_t src = smem_u32(sStage + xs * SLOT_BYTES) + src_off;
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[i]) : "r"(src + (uint32_t)i * src_stride));
      if (isA) {
#pragma unroll
        for (int i = 0; i < 32; ++i) colsum += x[i];
      }
      uint32_t hi[16], lo[16];
      pack_operand_row<false, true>(x, hi, lo);                // consumes every staged value
      mbar_wait(BAR(B_EMPTY + s), ph ^ 1u, p.err, 4);
      const uint32_t dst = smem_u32(sOp + s * STAGE_BYTES) + dst_off;
      store_operand_row<true>(dst, dst + lo_off, row, hi, lo);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(BAR(B_FULL + s)); mbar_arrive(BAR(B_XEMPTY + xs)); }
    }
    if (isA) {
      // ---- read-out: TMEM lane = output row n (warp cw owns lanes 32 cw .. +32), 256 columns = this tile's k range ----
      const int n = tn * TM + row;
      if (p.db_part && tk == 0 && n < p.Nn) p.db_part[(int64_t)sp * p.Nn + n] = colsum;
      float* dst = p.out + (int64_t)sp * p.Nn * p.K + (int64_t)n * p.K + (int64_t)tk * TN;
      if (steps > 0) {
        mbar_wait(BAR(B_ACCFULL), 0, p.err, 5);
        tc_fence_after();
      }
      const uint32_t tq = tmem + ((uint32_t)(cw * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < TN / 32; ++c) {
        float v[32];
        if (steps > 0) {
          tmem_ld32f(tq + (uint32_t)(c * 32), v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (n < p.Nn) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + c * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// out[i] = sum_s part[s][i] (fixed order); the same for the bias-gradient partials
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out, const float* __restrict__ db_part,
                                    int nb, float* __restrict__ db) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i < n) {
    float4 a = *reinterpret_cast<const float4*>(part + i);
    for (int s = 1; s < splits; ++s) {
      const float4 b = *reinterpret_cast<const float4*>(part + (int64_t)s * n + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(out + i) = a;
  }
  if (db && blockIdx.x == 0) {
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      float a = 0.f;
      for (int s = 0; s < splits; ++s) a += db_part[(int64_t)s * nb + j];
      db[j] = a;
    }
  }
}

static int plan(int64_t M, int Nn, int K, int* tiles_n, int* tiles_k, int* splits, int64_t* rows_per_split) {
  *tiles_n = Nn / TM;
  *tiles_k = K / TN;
  const int tiles = *tiles_n * *tiles_k;
  const int64_t steps = (M + BKM - 1) / BKM;
  int s = num_sms() / tiles;
  if (s < 1) s = 1;
  if (s > steps / 4) s = (int)(steps / 4);                     // at least 4 steps (128 instances) per slice
  if (s < 1) s = 1;
  int64_t per = (steps + s - 1) / s;
  s = (int)((steps + per - 1) / per);                          // drop empty slices
  *splits = s;
  *rows_per_split = per * BKM;
  return tiles * s;
}

}  // namespace wgrad
}  // namespace mil

using namespace mil;

extern "C" size_t mil_wgrad_tc_workspace_bytes(int64_t M, int Nn, int K) {
  if (M <= 0 || Nn <= 0 || K <= 0 || Nn % wgrad::TM || K % wgrad::TN) return 0;
  int tn, tk, sp;
  int64_t per;
  wgrad::plan(M, Nn, K, &tn, &tk, &sp, &per);
  return (size_t)sp * ((size_t)Nn * K + (size_t)Nn) * sizeof(float) + 256;
}

extern "C" int mil_wgrad_tc_f32(const float* G, int64_t ldg, const float* X, int64_t ldx, int64_t M, int Nn, int K, float* dW, float* db,
                                void* ws, size_t ws_bytes, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(mil_device_supported(), "mil_wgrad_tc_f32: needs a compute-capability 10.x device (tcgen05/TMEM/TMA)");
  MIL_CHECK_ARG(G && X && dW && ws && M > 0 && M < (1ll << 31) - 64, "mil_wgrad_tc_f32: bad argument");
  MIL_CHECK_ARG(Nn >= wgrad::TM && Nn % wgrad::TM == 0 && K >= wgrad::TN && K % wgrad::TN == 0,
                "mil_wgrad_tc_f32: dW must be [multiple of %d, multiple of %d] (got %d x %d)", wgrad::TM, wgrad::TN, Nn, K);
  MIL_CHECK_ARG(ldg >= Nn && ldx >= K && ldg % 4 == 0 && ldx % 4 == 0, "mil_wgrad_tc_f32: leading dimensions must cover the rows and be multiples of 4");
  MIL_CHECK_ARG((uintptr_t)G % 16 == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)dW % 16 == 0, "mil_wgrad_tc_f32: pointers must be 16-byte aligned");
  MIL_CHECK_ARG(ws_bytes >= mil_wgrad_tc_workspace_bytes(M, Nn, K), "mil_wgrad_tc_f32: workspace needs %zu bytes", mil_wgrad_tc_workspace_bytes(M, Nn, K));
  wgrad::Params p;
  p.M = M; p.Nn = Nn; p.K = K;
  const int grid = wgrad::plan(M, Nn, K, &p.tiles_n, &p.tiles_k, &p.splits, &p.rows_per_split);
  uint8_t* w = (uint8_t*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* part = (float*)w;
  float* db_part = part + (size_t)p.splits * Nn * K;
  p.out = p.splits == 1 ? dW : part;
  p.db_part = db ? (p.splits == 1 ? db : db_part) : nullptr;
  p.err = nullptr;
  CUtensorMap mg, mx;
  int rc;
  // 2-D maps over the row-major operands with their leading dimensions; rows past M read as zeros (TMA out-of-bounds fill)
  if ((rc = make_map_2d_ld(&mg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, G, (uint64_t)M, (uint64_t)Nn, (uint64_t)ldg, wgrad::BKM, wgrad::TM, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  if ((rc = make_map_2d_ld(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, wgrad::BKM, wgrad::TN, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  const size_t smem = 1024 + (size_t)wgrad::NSLOT * wgrad::SLOT_BYTES + (size_t)wgrad::NST * wgrad::STAGE_BYTES + 256;
  MIL_CUDA(cudaFuncSetAttribute(wgrad::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mil_set_notrap();
  wgrad::wgrad_kernel<<<grid, wgrad::NUM_THREADS, smem, stream>>>(mg, mx, p);
  MIL_LAUNCH_CHECK();
  if (p.splits > 1) {
    const int64_t n = (int64_t)Nn * K;
    wgrad::wgrad_reduce_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream>>>(part, p.splits, n, dW, db ? db_part : nullptr, Nn, db);
    MIL_LAUNCH_CHECK();
  }
  return 0;
}
