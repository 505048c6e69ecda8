// Fused ABMIL forward for sm_100a: one persistent, warp-specialised kernel that streams the bag X[N,D] (fp32) from HBM
// exactly once and produces per-CTA softmax-pool partials.
//
//   HBM --TMA(SW128)--> fp32 staging ring --converter warps--> 16-bit hi(/lo) operand ring (UMMA K-major SW64)
//   L2  --TMA(SW64)---> weight ring (W1 hi/lo chunks, then Wa hi/lo chunks)
//   tcgen05.mma (one thread) : pre[128 x 512] = X_tile W1^T   -> TMEM columns [0,512)       (GEMM1)
//   epilogue warps           : h = act(pre + b1) -> 16-bit hi/lo -> operand ring (A2), h kept in TMEM / registers
//   tcgen05.mma              : u[128 x Da] = h Wa^T            -> TMEM columns [0,Da)        (GEMM2)
//   epilogue warps           : s = wc . tanh(u + ba) + bc, online softmax over rows, p += e^{s-m} h  (warp shuffles)
//
// Precision: MIL_PREC_BF16X3 splits every fp32 operand into bf16 hi + lo and issues 3 products (hi.hi, lo.hi, hi.lo)
// into the same fp32 TMEM accumulator (fp32-class results); MIL_PREC_FP16 / MIL_PREC_BF16 issue one product.
//
// The same pipeline with a "store" epilogue is the tensor-core Linear(+bias+act) used to materialise h and as the
// self-test of the TMA/UMMA plumbing (mil_umma_selftest_f32).
#include <stdlib.h>
#include <utility>
#include <vector>

#include "mil_umma.cuh"

namespace mil {

// ------------------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------------------
constexpr int BM = 128;                 // rows per tile (UMMA M)
constexpr int BK = 32;                  // K elements per pipeline stage (two UMMA K=16 steps); 64-byte rows -> SWIZZLE_64B
constexpr int XS = 3;                   // fp32 staging slots (128 rows x 32 floats, SWIZZLE_128B)
constexpr int X_SLOT_BYTES = BM * BK * 4;           // 16384
constexpr int A_OP_BYTES = BM * BK * 2;             // 8192  (one 16-bit operand tile)
constexpr int B_OP_BYTES = HMAX * BK * 2;           // 32768
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 8;            // warps 8..15: epilogue; 4..7: converters; 0: X TMA; 1: MMA; 2: TMEM alloc; 3: W TMA

enum { MODE_FUSED = 0, MODE_STORE = 1 };


// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
template <int NPROD, bool FP16, int NST, int MODE, int ACT, int ATT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mil_fused_kernel(const __grid_constant__ CUtensorMap mapX, const FusedParams p) {
  constexpr bool LO = NPROD == 3;
  constexpr int NOP = LO ? 2 : 1;                         // operand tiles per stage (hi, lo)
  constexpr uint32_t A_STAGE = NOP * A_OP_BYTES, B_STAGE = NOP * B_OP_BYTES;
  // converter warps per k-step: the 3-product mode (1536 MMA cycles per k-step) keeps one row per thread on all four warps;
  // the 1-product modes (512 cycles) alternate two groups of two warps (two rows per thread) so that each group has two
  // k-step times for its wait -> load -> pack -> store -> fence -> arrive chain (~800 cycles)
  constexpr int CW = LO ? 4 : 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                                      // XS x 16 KB
  uint8_t* sA = sX + XS * X_SLOT_BYTES;                    // NST x A_STAGE
  uint8_t* sB = sA + NST * A_STAGE;                        // NST x B_STAGE
  uint8_t* sMisc = sB + NST * B_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMisc);     // barrier block
  float* s_part = reinterpret_cast<float*>(sMisc + 512);   // [2][128] partial attention logits (+ [2][128][4] t partials)
  float* t_part = s_part + 256;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sMisc + 512 + 1024 + 4096);
  float* c_b1 = reinterpret_cast<float*>(sMisc + 512 + 1024 + 4096 + 16);   // [512] feature bias
  float* c_ba = c_b1 + HMAX;                                                 // [128] attention bias
  float* c_wc = c_ba + 128;                                                  // [128] attention output weights
  float* p_acc = c_wc + 128;                                                 // [8 warps][8 chunks][32 lanes] pooled partial sums

  // barrier indices
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int XG = LO ? 1 : 2;                               // converter groups (see the converters)
  constexpr int B_XFULL = 0, B_XEMPTY = B_XFULL + XG * XS, B_FULL = B_XEMPTY + XG * XS, B_EMPTY = B_FULL + NST,
                B_ACCFULL = B_EMPTY + NST, B_ACCEMPTY = B_ACCFULL + 1, B_TAILFREE = B_ACCEMPTY + 1, B_UFULL = B_TAILFREE + 1,
                B_G2AFULL = B_UFULL + 1, B_G2BFULL = B_G2AFULL + 4, B_G2EMPTY = B_G2BFULL + 4, B_FIN = B_G2EMPTY + 4, B_COUNT = B_FIN + 2;
  static_assert(B_COUNT * 8 <= 512, "barrier block overflow");

  // GEMM2 has its own 4-deep operand ring (barriers G2*).  Its buffers are carved out of the GEMM1 stage slots, which are idle
  // between ACC_FULL (all GEMM1 MMAs of the tile done) and U_FULL (all GEMM2 MMAs done):
  //   1-product modes: ring entry r = GEMM1 slot r (NST = 4).
  //   3-product mode (NST = 2, slot = A 16 KB + B 64 KB): r = (slot = r & 1, sub = r >> 1);
  //     sub 0: A2 hi/lo in the A slot, Wa hi @ B+0, Wa lo @ B+32K;  sub 1: A2 hi @ B+8K, A2 lo @ B+40K, Wa hi @ B+16K, Wa lo @ B+48K.
  auto g2_a_hi = [&](uint32_t r) -> uint32_t {
    if (LO) return (r >> 1) ? smem_u32(sB + (r & 1) * B_STAGE + 8192) : smem_u32(sA + (r & 1) * A_STAGE);
    return smem_u32(sA + r * A_STAGE);
  };
  auto g2_a_lo = [&](uint32_t r) -> uint32_t { return g2_a_hi(r) + ((r >> 1) ? (uint32_t)B_OP_BYTES : (uint32_t)A_OP_BYTES); };
  auto g2_b_hi = [&](uint32_t r) -> uint32_t {
    if (LO) return smem_u32(sB + (r & 1) * B_STAGE + ((r >> 1) ? 16384 : 0));
    return smem_u32(sB + r * B_STAGE);
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.trace && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[256 + 2 * blockIdx.x] = (long long)gt;
  }
  const int KS = p.D / BK;                                 // GEMM1 k-steps per tile
  const int NCH2 = (MODE == MODE_FUSED) ? HMAX / BK : 0;   // GEMM2 k-steps per tile (16)
  const int64_t n_tiles = (p.N + BM - 1) / BM;
  // store mode, column-split launch: virtual tile v = (row tile v / nblk, 256-column block v % nblk); p.nout = columns per virtual tile
  const int nblk = (MODE == MODE_STORE) ? p.nblk : 1;
  const int64_t n_vt = n_tiles * nblk;

  if (threadIdx.x == 0) {
    // FULL[s]: the converter warps of the k-step + the weight producer's expect_tx arrival (+ its bytes)
    for (int i = 0; i < XG * XS; ++i) { mbar_init(BAR(B_XFULL + i), 1); mbar_init(BAR(B_XEMPTY + i), CW); }
    for (int i = 0; i < NST; ++i) { mbar_init(BAR(B_FULL + i), CW + 1); mbar_init(BAR(B_EMPTY + i), 1); }
    mbar_init(BAR(B_ACCFULL), 1); mbar_init(BAR(B_ACCEMPTY), 8); mbar_init(BAR(B_TAILFREE), 8); mbar_init(BAR(B_UFULL), 1);
    for (int i = 0; i < 4; ++i) { mbar_init(BAR(B_G2AFULL + i), 4); mbar_init(BAR(B_G2BFULL + i), 1); mbar_init(BAR(B_G2EMPTY + i), 1); }
    mbar_init(BAR(B_FIN), 1); mbar_init(BAR(B_FIN + 1), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&mapX);
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Register file re-balancing between the roles (512 threads x 128 at launch; 4 x 32 x (56 + 104 + 176 + 176) = 65536): the
  // setmaxnreg of a role sits at the top of that role's branch so that ptxas allocates the branch against the new budget.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== X producer: HBM -> fp32 staging (TMA, SWIZZLE_128B) =====================
    // In the 2-group modes consecutive uses of a staging slot (XS is odd) go to alternating converter groups.  A parity wait is only
    // sound if the waiter sees every phase of its barrier, so each (group, slot) pair has its own XFULL / XEMPTY barrier: entry
    // (g, slot) is used every 2 * XS k-steps, its n-th use is k-step it with n = (it / XS) >> 1.
    static_assert(XG == 1 || (XS & 1), "the (group, slot) barrier scheme assumes an odd staging ring");
    // The staging ring (3 x 16 KB) cannot cover an HBM round trip of the bag stream (~1800 cycles x 32 B/cycle), so every box is
    // first pulled into L2 PF k-steps ahead (cp.async.bulk.prefetch.tensor); the staged load then sees L2 latency.
    if (lane == 0) {
      uint32_t it = 0;
      constexpr int PF = 16;
      int64_t ptile = blockIdx.x;
      int pks = 0;
      auto prefetch_next = [&]() {
        if (ptile < n_vt && !(p.dbg & 2)) {
          tma_prefetch_2d(&mapX, pks * BK, (int)((ptile / nblk) * BM));
          if (++pks == KS) { pks = 0; ptile += gridDim.x; }
        }
      };
      for (int i = 0; i < PF; ++i) prefetch_next();
      for (int64_t tile = blockIdx.x; tile < n_vt; tile += gridDim.x) {
        for (int ks = 0; ks < KS; ++ks, ++it) {
          prefetch_next();
          const uint32_t s = it % XS, m = it / XS;
          if (XG == 1) {
            mbar_wait(BAR(B_XEMPTY + s), (m & 1) ^ 1, p.err, 1);
          } else if (m > 0) {                                 // the slot's previous use (k-step it - XS) belonged to the other group
            const uint32_t gp = (it & 1) ^ 1;
            mbar_wait(BAR(B_XEMPTY + gp * XS + s), ((m - 1) >> 1) & 1, p.err, 1);
          }
          const uint32_t full = BAR(B_XFULL + (XG == 1 ? 0 : (it & 1) * XS) + s);
          if (p.dbg & 2) { mbar_arrive(full); continue; }
          mbar_expect_tx(full, X_SLOT_BYTES);
          tma_load_2d(smem_u32(sX + s * X_SLOT_BYTES), &mapX, full, ks * BK, (int)((tile / nblk) * BM));
        }
      }
    }
  } else if (warp == 3) {
    // ===================== weight producer: L2 -> weight ring =====================
    // The weights were pre-arranged (split_weights_kernel) as the exact shared-memory image of every stage's operand tile
    // (UMMA K-major SWIZZLE_64B), so each stage is ONE contiguous bulk copy per operand instead of 256-512 64-byte TMA rows.
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      // W1 image: 256-row blocks (split_weights_kernel); this CTA consumes nimg of them per stage, starting at block blk0 of its virtual tile
      const uint32_t w1_tile = (uint32_t)(p.nout < 256 ? p.nout : 256) * BK * 2, wa_tile = (uint32_t)p.Da * BK * 2;
      const int nimg = (p.nout + 255) / 256;
      for (int64_t tile = blockIdx.x; tile < n_vt; tile += gridDim.x, ++tl) {
        const int blk0 = (int)(tile % nblk);
        // the GEMM1 stage slots host GEMM2's ring between ACC_FULL and U_FULL of the previous tile
        if (MODE == MODE_FUSED && tl > 0) mbar_wait(BAR(B_UFULL), (tl - 1) & 1, p.err, 17);
        for (int ks = 0; ks < KS; ++ks, ++it) {
          const uint32_t s = it % NST, ph = (it / NST) & 1;
          mbar_wait(BAR(B_EMPTY + s), ph ^ 1, p.err, 2);
          if (p.dbg & 1) { mbar_arrive(BAR(B_FULL + s)); continue; }
          mbar_expect_tx(BAR(B_FULL + s), NOP * nimg * w1_tile);
          const uint32_t dst = smem_u32(sB + s * B_STAGE);
          for (int hf = 0; hf < nimg; ++hf) {
            const uint8_t* src = p.w1_img + ((size_t)(blk0 + hf) * KS + ks) * NOP * w1_tile;
            bulk_load(dst + hf * w1_tile, src, w1_tile, BAR(B_FULL + s));
            if (LO) bulk_load(dst + B_OP_BYTES + hf * w1_tile, src + w1_tile, w1_tile, BAR(B_FULL + s));
          }
        }
        if (NCH2 > 0) mbar_wait(BAR(B_ACCFULL), tl & 1, p.err, 18);
        for (int c = 0; c < NCH2; ++c) {
          const uint32_t r = c & 3, ph = (tl * 4 + (c >> 2)) & 1;
          mbar_wait(BAR(B_G2EMPTY + r), ph ^ 1, p.err, 3);
          if (p.dbg & 16) { mbar_arrive(BAR(B_G2BFULL + r)); continue; }
          mbar_expect_tx(BAR(B_G2BFULL + r), NOP * wa_tile);
          const uint32_t dst = g2_b_hi(r);
          const uint8_t* src = p.wa_img + (size_t)c * NOP * wa_tile;
          bulk_load(dst, src, wa_tile, BAR(B_G2BFULL + r));
          if (LO) bulk_load(dst + B_OP_BYTES, src + wa_tile, wa_tile, BAR(B_G2BFULL + r));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The warp runs CONVERGED: all lanes wait and compute the (uniform) descriptors, one elected lane issues.  With a lone diverged
    // lane ptxas moved every descriptor into uniform registers through an ELECT + R2UR waterfall per MMA: ~780 cycles per k-step.
    {
      uint32_t s = 0, ph = 0, tl = 0;
      const int n1 = p.nout < 256 ? p.nout : 256, nhalf = (p.nout + 255) / 256;
      const uint32_t idesc1 = make_idesc(FP16, n1), idesc2 = make_idesc(FP16, p.Da);
      const uint32_t sa0 = smem_u32(sA), sb0 = smem_u32(sB);
      for (int64_t tile = blockIdx.x; tile < n_vt; tile += gridDim.x, ++tl) {
        mbar_wait(BAR(B_ACCEMPTY), (tl & 1) ^ 1, p.err, 4);
        tc_fence_after();
        if (lane == 0) trace_stamp(p, tl, 0);                // GEMM1 may start
        for (int ks = 0; ks < KS; ++ks) {
          mbar_wait(BAR(B_FULL + s), ph, p.err, 5);
          tc_fence_after();
          const uint32_t a0 = sa0 + s * A_STAGE, b0 = sb0 + s * B_STAGE;
          const uint64_t ah0 = make_desc_sw64(a0), bh0 = make_desc_sw64(b0), al0 = make_desc_sw64(a0 + A_OP_BYTES), bl0 = make_desc_sw64(b0 + B_OP_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k16 = 0; k16 < 2; ++k16) {
              if (p.dbg & 4) break;
              for (int hf = 0; hf < nhalf; ++hf) {
                // descriptor start addresses are in 16-byte units: +2 per K = 16 step (32 B), +1024 per 256-row N half (16 KB)
                const uint64_t oa = (uint64_t)(k16 * 2), ob = (uint64_t)(hf * (256 * BK * 2 / 16) + k16 * 2);
                const uint32_t d = tmem + (uint32_t)(hf * 256);
                umma_f16(d, ah0 + oa, bh0 + ob, idesc1, (ks | k16) ? 1u : 0u);
                if (LO) {
                  umma_f16(d, al0 + oa, bh0 + ob, idesc1, 1u);
                  umma_f16(d, ah0 + oa, bl0 + ob, idesc1, 1u);
                }
              }
            }
            umma_commit(BAR(B_EMPTY + s));
            if (ks == KS - 1) umma_commit(BAR(B_ACCFULL));
          }
          __syncwarp();
          if (++s == (uint32_t)NST) { s = 0; ph ^= 1; }
        }
        if (lane == 0) trace_stamp(p, tl, 1);                // GEMM1 fully issued
        if (MODE == MODE_FUSED) {
          mbar_wait(BAR(B_TAILFREE), tl & 1, p.err, 7);
          tc_fence_after();
          if (lane == 0) trace_stamp(p, tl, 2);              // GEMM2 may start
          for (int c = 0; c < NCH2; ++c) {
            const uint32_t r = c & 3, ph2 = (tl * 4 + (c >> 2)) & 1;
            mbar_wait(BAR(B_G2AFULL + r), ph2, p.err, 8);
            mbar_wait(BAR(B_G2BFULL + r), ph2, p.err, 9);
            tc_fence_after();
            const uint32_t a0 = g2_a_hi(r), a0l = g2_a_lo(r), b0 = g2_b_hi(r);
            const uint64_t ah0 = make_desc_sw64(a0), bh0 = make_desc_sw64(b0), al0 = make_desc_sw64(a0l), bl0 = make_desc_sw64(b0 + B_OP_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16) {
                if (p.dbg & 64) break;
                const uint64_t o = (uint64_t)(k16 * 2);
                umma_f16(tmem, ah0 + o, bh0 + o, idesc2, (c | k16) ? 1u : 0u);
                if (LO) {
                  umma_f16(tmem, al0 + o, bh0 + o, idesc2, 1u);
                  umma_f16(tmem, ah0 + o, bl0 + o, idesc2, 1u);
                }
              }
              umma_commit(BAR(B_G2EMPTY + r));
              if (c == NCH2 - 1) umma_commit(BAR(B_UFULL));
            }
            __syncwarp();
          }
          if (lane == 0) trace_stamp(p, tl, 3);              // GEMM2 fully issued
        }
      }
    }
  }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // ===================== converters: fp32 staging -> 16-bit hi/lo operand tiles =====================
    constexpr int RPT = LO ? 1 : 2;                           // rows per thread
    const int grp = LO ? 0 : (warp - 4) >> 1;                 // 1-product modes: group 0 = warps 4, 5 (even k-steps), group 1 = warps 6, 7 (odd)
    const int row0 = LO ? (warp - 4) * 32 + lane : ((warp - 4) & 1) * 32 + lane;   // rows row0 and (RPT == 2) row0 + 64
    uint32_t itx = 0, ita = 0, tl = 0;
    for (int64_t tile = blockIdx.x; tile < n_vt; tile += gridDim.x, ++tl) {
      // The stage slots host GEMM2's operand ring until U_FULL of the previous tile: do not overwrite them earlier.
      if (MODE == MODE_FUSED && tl > 0) mbar_wait(BAR(B_UFULL), (tl - 1) & 1, p.err, 16);
      for (int ks = 0; ks < KS; ++ks, ++itx, ++ita) {
        if (!LO && (int)(itx & 1) != grp) continue;
        static_assert(LO || !(NST & 1), "two converter groups need an even stage ring (each stage always belongs to the same group)");
        const uint32_t xs = itx % XS, xph = (LO ? itx / XS : (itx / XS) >> 1) & 1, xb = (LO ? 0 : grp * XS) + xs;   // this group's barrier of the slot
        const uint32_t s = ita % NST, ph = (ita / NST) & 1;
        mbar_wait(BAR(B_XFULL + xb), xph, p.err, 10);
        uint32_t hi[RPT][16], lo[RPT][16];
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
          const int row = row0 + rr * 64;
          float x[32];
          const uint32_t src = smem_u32(sX + xs * X_SLOT_BYTES) + (uint32_t)row * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = src + (((uint32_t)j ^ ((uint32_t)row & 7u)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[4 * j]), "=f"(x[4 * j + 1]), "=f"(x[4 * j + 2]), "=f"(x[4 * j + 3]) : "r"(a));
          }
          pack_operand_row<FP16, LO>(x, hi[rr], lo[rr]);      // consumes every staged value: the loads are complete afterwards
        }
        // (Handing the staging slot back right here, before the stage wait, looked free but produced a wrong row about once in 10^5
        // k-steps: the slot is released only after the operand stores below.)
        mbar_wait(BAR(B_EMPTY + s), ph ^ 1, p.err, 11);
        const uint32_t a_hi = smem_u32(sA + s * A_STAGE);
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int rr = 0; rr < RPT; ++rr) store_operand_row<LO>(a_hi, a_hi + A_OP_BYTES, row0 + rr * 64, hi[rr], lo[rr]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(BAR(B_FULL + s)); mbar_arrive(BAR(B_XEMPTY + xb)); }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    // ===================== epilogue warps =====================
    const int q = warp & 3;                                 // TMEM lane quarter this warp may touch
    const int half = (warp - EPI_WARP0) >> 2;               // two warps per quarter split the columns
    const int row = q * 32 + lane;                          // row inside the tile
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    const int et = threadIdx.x - EPI_WARP0 * 32;            // 0..255
    uint32_t tl = 0;

    // per-column constants -> shared memory (broadcast reads in the hot loops)
    for (int i = et; i < HMAX; i += 256) c_b1[i] = (p.b1 && i < p.nout * nblk) ? p.b1[i] : 0.f;   // the bias has nout * nblk entries (<= 512)
    if (MODE == MODE_FUSED) {
      for (int i = et; i < 128; i += 256) { c_ba[i] = p.ba ? p.ba[i] : 0.f; c_wc[i] = p.wc[i]; }
    } else {
      for (int i = et; i < 128; i += 256) c_ba[i] = 0.f;
    }
    const float* c_zero = c_ba;                             // 32+ zeros (store mode only)
    named_bar_sync(1, 256);

    if (MODE == MODE_STORE) {
      const int nch = p.nout / 32;
      for (int64_t tile = blockIdx.x; tile < n_vt; tile += gridDim.x, ++tl) {
        mbar_wait(BAR(B_ACCFULL), tl & 1, p.err, 12);
        tc_fence_after();
        const int64_t grow = (tile / nblk) * BM + row;
        const int cb = (int)(tile % nblk) * nch;              // first 32-column chunk of this virtual tile's column block
#pragma unroll 1
        for (int cc = half; cc < nch; cc += 2) {
          const int c = cb + cc;                              // chunk index in the full output row
          float hv[32];
          tmem_ld32f(tq + (uint32_t)(cc * 32), hv);
          if (p.h_out) {                                      // pre-activation (needed by the GELU backward)
            bias_act32<MIL_ACT_NONE>(hv, c_b1 + c * 32, 0, p.w1_inv);
            if (grow < p.N) {
              float* dst = p.h_out + grow * p.ldc + c * 32;
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(hv[j], hv[j + 1], hv[j + 2], hv[j + 3]);
            }
            bias_act32<-1>(hv, c_zero, p.act);
          } else {
            bias_act32<-1>(hv, c_b1 + c * 32, p.act, p.w1_inv);
          }
          if (p.drop_mode) drop_apply32(hv, drop_keep_word(p, grow, c, nch * nblk), p.drop_scale);
          if (grow < p.N) {
            float* dst = p.c_out + grow * p.ldc + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(hv[j], hv[j + 1], hv[j + 2], hv[j + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ACCEMPTY));
      }
    } else {
      // running online-softmax state of this warp: rows = its 32 lanes, columns = its 8 chunks (chunk c = 2 j + half);
      // pooled partial sums: prun[j] (this lane) = column (2 j + half) * 32 + lane
      float m_run = -INFINITY, l_run = 0.f;
      float prun[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // registers; handed over through p_acc after the last tile
      const float bc = p.bc ? p.bc[0] : 0.f;

      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const int64_t grow = tile * BM + row;
        mbar_wait(BAR(B_ACCFULL), tl & 1, p.err, 13);
        tc_fence_after();
        if (et == 0) trace_stamp(p, tl, 4);                  // accumulator complete

        // E1: vacate the first 128 accumulator columns (they become GEMM2's accumulator); keep h for them in registers
        float keep_h[2][32];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = 2 * j + half;
          tmem_ld32f(tq + (uint32_t)(c * 32), keep_h[j]);
          bias_act32<ACT>(keep_h[j], c_b1 + c * 32, p.act, p.w1_inv);
          if (p.drop_mode) drop_apply32(keep_h[j], drop_keep_word(p, grow, c, HMAX / 32), p.drop_scale);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_TAILFREE));
        if (et == 0) trace_stamp(p, tl, 5);                  // E1 done

        // E2: h chunk -> 16-bit operand tiles for GEMM2 (+ h back into TMEM for the pooling pass, + optional outputs)
        float tacc[4] = {0.f, 0.f, 0.f, 0.f};
        auto emit_chunk = [&](int c, const float (&hv)[32]) {
          if (p.h_out && grow < p.N) {
            float* dst = p.h_out + grow * HMAX + c * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(hv[i], hv[i + 1], hv[i + 2], hv[i + 3]);
          }
          if (p.t_out) {
#pragma unroll 1
            for (int cc = 0; cc < p.C; ++cc) {
              const float4* wp = reinterpret_cast<const float4*>(p.Wp + cc * HMAX + c * 32);
              float a = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 w4 = __ldg(wp + i);
                a = fmaf(hv[4 * i], w4.x, a); a = fmaf(hv[4 * i + 1], w4.y, a); a = fmaf(hv[4 * i + 2], w4.z, a); a = fmaf(hv[4 * i + 3], w4.w, a);
              }
              tacc[cc] += a;
            }
          }
          const uint32_t r = (uint32_t)c & 3u, ph = (tl * 4 + ((uint32_t)c >> 2)) & 1u;
          mbar_wait(BAR(B_G2EMPTY + r), ph ^ 1, p.err, 14);
          write_operand_row<FP16, LO>(g2_a_hi(r), g2_a_lo(r), row, hv);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_G2AFULL + r));
        };
        emit_chunk(half, keep_h[0]);
        emit_chunk(2 + half, keep_h[1]);
#pragma unroll 1
        for (int j = 2; j < 8; ++j) {
          const int c = 2 * j + half;
          float hv[32];
          tmem_ld32f(tq + (uint32_t)(c * 32), hv);
          bias_act32<ACT>(hv, c_b1 + c * 32, p.act, p.w1_inv);
          if (p.drop_mode) drop_apply32(hv, drop_keep_word(p, grow, c, HMAX / 32), p.drop_scale);
          tmem_st32f(tq + (uint32_t)(c * 32), hv);
          emit_chunk(c, hv);
        }
        tmem_wait_st();
        if (et == 0) trace_stamp(p, tl, 6);                  // E2 done (all my A2 chunks produced)

        // E3: attention logit of every row: s = wc . f(u + ba) + bc   (this warp: 64 of the Da = 128 columns)
        mbar_wait(BAR(B_UFULL), tl & 1, p.err, 15);
        tc_fence_after();
        if (et == 0) trace_stamp(p, tl, 7);                  // u complete
        float sp = 0.f;
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const int c0 = half * 64 + j * 32;
          float uv[32];
          tmem_ld32f(tq + (uint32_t)c0, uv);
          bias_act32<ATT>(uv, c_ba + c0, p.att_act, p.wa_inv);
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(c_wc + c0 + i);
            s4[0] = fmaf(uv[i], w4.x, s4[0]); s4[1] = fmaf(uv[i + 1], w4.y, s4[1]);
            s4[2] = fmaf(uv[i + 2], w4.z, s4[2]); s4[3] = fmaf(uv[i + 3], w4.w, s4[3]);
          }
          sp += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
        s_part[half * 128 + row] = sp;
        if (p.t_out)
          for (int cc = 0; cc < p.C; ++cc) t_part[(half * 128 + row) * 4 + cc] = tacc[cc];
        named_bar_sync(1, 256);
        float sv = s_part[row] + s_part[128 + row] + bc;
        const bool valid = grow < p.N && (!p.keep || p.keep[grow]);
        if (!valid) sv = -INFINITY;
        if (half == 0 && grow < p.N) {
          if (p.s_out) p.s_out[grow] = sv;
          if (p.t_out)
            for (int cc = 0; cc < p.C; ++cc) p.t_out[grow * p.C + cc] = t_part[row * 4 + cc] + t_part[(128 + row) * 4 + cc];
        }
        named_bar_sync(1, 256);                             // s_part / t_part may be overwritten by the next tile after this

        if (et == 0) trace_stamp(p, tl, 8);                  // E3 done (scores exchanged)
        // online softmax over the 32 rows of this warp
        const float m_new = fmaxf(m_run, warp_max(sv));
        float w = 0.f, scale = 1.f;
        if (m_new > -INFINITY) {
          scale = (m_run > -INFINITY) ? expf(m_run - m_new) : 0.f;
          w = valid ? expf(sv - m_new) : 0.f;
        }
        l_run = l_run * scale + warp_sum(w);
        m_run = m_new;

        // E4: p += sum_rows w * h  (h from registers for the vacated columns, from TMEM otherwise); two chunks per step
        {
#pragma unroll
          for (int i = 0; i < 32; ++i) { keep_h[0][i] *= w; keep_h[1][i] *= w; }
          float r0, r1;
          warp_transpose_sum2(keep_h[0], keep_h[1], r0, r1);
          prun[0] = prun[0] * scale + r0;
          prun[1] = prun[1] * scale + r1;
        }
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
          float ha[32], hb[32];
          uint32_t va[32], vb[32];
          tmem_ld32(tq + (uint32_t)((2 * j + half) * 32), va);
          tmem_ld32(tq + (uint32_t)((2 * j + 2 + half) * 32), vb);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) { ha[i] = __uint_as_float(va[i]) * w; hb[i] = __uint_as_float(vb[i]) * w; }
          float r0, r1;
          warp_transpose_sum2(ha, hb, r0, r1);
          prun[j] = prun[j] * scale + r0;
          prun[j + 1] = prun[j + 1] * scale + r1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ACCEMPTY));
        if (et == 0) trace_stamp(p, tl, 9);                  // E4 done
      }

      // CTA partial: merge the 4 row quarters (the two column halves share m, l) -> part[blockIdx] = (m, l, P[512])
      float* red_m = s_part;                                // [4]
      float* red_l = s_part + 4;                            // [4]
      named_bar_sync(1, 256);
#pragma unroll
      for (int j = 0; j < 8; ++j) p_acc[(warp - EPI_WARP0) * 256 + j * 32 + lane] = prun[j];
      if (half == 0 && lane == 0) { red_m[q] = m_run; red_l[q] = l_run; }
      named_bar_sync(1, 256);
      const float m_cta = fmaxf(fmaxf(red_m[0], red_m[1]), fmaxf(red_m[2], red_m[3]));
      float* out = p.part + (int64_t)blockIdx.x * (2 + HMAX);
      // column c = (2 j + half) * 32 + lane lives in prun of warp (half * 4 + q'), q' = 0..3: sum the quarters in fixed order
      for (int c = et; c < HMAX; c += 256) {
        const int ch = c >> 5, ln = c & 31, hf = ch & 1, j = ch >> 1;
        float v = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float f = (red_m[qq] > -INFINITY) ? expf(red_m[qq] - m_cta) : 0.f;
          v = fmaf(p_acc[(hf * 4 + qq) * 256 + j * 32 + ln], f, v);
        }
        out[2 + c] = v;
      }
      if (et == 0) {
        float l = 0.f;
        for (int i = 0; i < 4; ++i) l += (red_m[i] > -INFINITY) ? red_l[i] * expf(red_m[i] - m_cta) : 0.f;
        out[0] = m_cta;
        out[1] = l;
      }

      grid_finalize(p, et, lane, t_part, smem, (uint32_t)(sMisc - smem), BAR(B_FIN), reinterpret_cast<int*>(s_part + 16));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[256 + 2 * blockIdx.x + 1] = (long long)gt;
  }
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------------------
// fp32 weights W[R, K] -> pre-swizzled 16-bit operand image (once per weight version; 2.4 MB of weights vs 205 MB of bag)
//   image = for every 256-row block of W (one block if R <= 256), for every k-step ks (32 elements): [hi tile | lo tile], each tile =
//   min(R, 256) rows x 64 B laid out exactly as the
//   UMMA K-major SWIZZLE_64B shared-memory tile: byte(r, c, e) = (r/8)*512 + (r%8)*64 + ((c ^ ((r>>1)&3))*16) + 2e,
//   c = 16-byte chunk (8 elements) inside the 64-byte row.  One thread converts one (row, chunk).
// ------------------------------------------------------------------------------------------------------------
template <bool FP16, bool LO>
__global__ void split_weights_kernel(const float* __restrict__ w, int R, int K, uint8_t* __restrict__ img, float scale) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (r, kc): kc = k / 8
  const int kchunks = K / 8;
  if (item >= (int64_t)R * kchunks) return;
  const int r = (int)(item / kchunks), kc = (int)(item % kchunks);
  const int ks = kc >> 2, c = kc & 3;
  float4 a = *reinterpret_cast<const float4*>(w + (int64_t)r * K + kc * 8);
  float4 b = *reinterpret_cast<const float4*>(w + (int64_t)r * K + kc * 8 + 4);
  a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
  uint32_t h[4], l[4];
  h[0] = pack_hi<FP16>(a.x, a.y); h[1] = pack_hi<FP16>(a.z, a.w); h[2] = pack_hi<FP16>(b.x, b.y); h[3] = pack_hi<FP16>(b.z, b.w);
  constexpr int NOPK = LO ? 2 : 1;
  // R > 256: one image per 256-row block, block after block, so that a CTA can take either the whole width (two block tiles per stage) or
  // one 256-column block of it (store mode, column-split launches) from the SAME image
  const int rb = R > 256 ? (r & 255) : r, blk = R > 256 ? (r >> 8) : 0;
  const size_t tile = (size_t)(R > 256 ? 256 : R) * 64;
  const size_t off = (size_t)(rb >> 3) * 512 + (size_t)(rb & 7) * 64 + (size_t)((c ^ ((rb >> 1) & 3)) << 4);
  uint8_t* dst = img + ((size_t)blk * (K / 32) + ks) * NOPK * tile + off;
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (LO) {
    l[0] = pack_lo<FP16>(a.x, a.y, h[0]); l[1] = pack_lo<FP16>(a.z, a.w, h[1]); l[2] = pack_lo<FP16>(b.x, b.y, h[2]); l[3] = pack_lo<FP16>(b.z, b.w, h[3]);
    *reinterpret_cast<uint4*>(dst + tile) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

int make_map_2d_ld(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                   uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -2; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  static const int promo_env = getenv("MHIMK_L2PROMO") ? atoi(getenv("MHIMK_L2PROMO")) : 128;     // tuning probe: 0 / 64 / 128 / 256
  const CUtensorMapL2promotion promo = promo_env == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo_env == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo_env == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols); return -3; }
  return 0;
}

int make_map_2d(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint32_t box_cols, CUtensorMapSwizzle sw) {
  return make_map_2d_ld(m, dt, elem_bytes, base, rows, cols, cols, box_rows, box_cols, sw);
}

// optional kernel-only timing (bench.py roofline): CUDA events recorded on the launching stream around the fused kernel
static bool g_profile = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;

static cudaEvent_t g_prof_open = nullptr;
void prof_begin(cudaStream_t stream) {
  if (!g_profile) return;
  cudaEventCreate(&g_prof_open);
  cudaEventRecord(g_prof_open, stream);
}
void prof_end(cudaStream_t stream) {
  if (!g_profile || !g_prof_open) return;
  cudaEvent_t e1;
  cudaEventCreate(&e1);
  cudaEventRecord(e1, stream);
  g_prof_events.push_back({g_prof_open, e1});
  g_prof_open = nullptr;
}

template <int NPROD, bool FP16, int NST, int MODE, int ACT, int ATT>
static int launch_fused(const CUtensorMap& mx, const FusedParams& p, int grid, cudaStream_t stream) {
  constexpr int NOP = NPROD == 3 ? 2 : 1;
  const size_t smem = 1024 + (size_t)XS * X_SLOT_BYTES + (size_t)NST * NOP * (A_OP_BYTES + B_OP_BYTES) + 512 + 1024 + 4096 + 16 + (HMAX + 256) * 4 + 8 * 256 * 4;
  auto kern = mil_fused_kernel<NPROD, FP16, NST, MODE, ACT, ATT>;
  static bool attr_set_dev[64] = {false};        // the attribute is per device (context): one flag per device ordinal
  int attr_dev = 0;
  cudaGetDevice(&attr_dev);
  bool& attr_set = attr_set_dev[attr_dev & 63];
  if (!attr_set) {
    MIL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  mil_set_notrap();
  prof_begin(stream);
  kern<<<grid, NUM_THREADS, smem, stream>>>(mx, p);
  prof_end(stream);
  MIL_LAUNCH_CHECK();
  return 0;
}

int debug_mask() {
  const char* e = getenv("MHIMK_DEBUG");
  return e ? atoi(e) : 0;
}
static int grid_override(int grid) {
  const char* e = getenv("MHIMK_GRID");
  const int g = e ? atoi(e) : 0;
  return (g > 0 && g < grid) ? g : grid;
}

template <int MODE, int ACT, int ATT>
static int dispatch_prec(int precision, const CUtensorMap& mx, const FusedParams& p, int grid, cudaStream_t stream) {
  if (precision == MIL_PREC_BF16X3) return launch_fused<3, false, 2, MODE, ACT, ATT>(mx, p, grid, stream);
  if (precision == MIL_PREC_FP16X3) return launch_fused<3, true, 2, MODE, ACT, ATT>(mx, p, grid, stream);
  if (precision == MIL_PREC_FP16) return launch_fused<1, true, 4, MODE, ACT, ATT>(mx, p, grid, stream);
  return launch_fused<1, false, 4, MODE, ACT, ATT>(mx, p, grid, stream);
}

// Fused mode is instantiated for the activation pairs the reference can produce: feature act relu / gelu (abmil.py:183-186,
// mhim.py:71-74) x attention act tanh (abmil.py:195) or relu / gelu / tanh (MHIM da_act, baseline.py:17-22).
static int dispatch_fused(int precision, int mode, const CUtensorMap& mx, const FusedParams& p, int grid, cudaStream_t stream) {
  if (mode == MODE_STORE) return dispatch_prec<MODE_STORE, -1, -1>(precision, mx, p, grid, stream);
#define MIL_CASE(A, T) if (p.act == A && p.att_act == T) return dispatch_prec<MODE_FUSED, A, T>(precision, mx, p, grid, stream);
  MIL_CASE(MIL_ACT_RELU, MIL_ACT_TANH) MIL_CASE(MIL_ACT_GELU, MIL_ACT_TANH)
  MIL_CASE(MIL_ACT_RELU, MIL_ACT_RELU) MIL_CASE(MIL_ACT_GELU, MIL_ACT_RELU)
  MIL_CASE(MIL_ACT_RELU, MIL_ACT_GELU) MIL_CASE(MIL_ACT_GELU, MIL_ACT_GELU)
#undef MIL_CASE
  set_error("fused pass: unsupported activation pair act=%d att_act=%d (use the composed path)", p.act, p.att_act);
  return -1;
}

int set_dropout(FusedParams& p, const mil_dropout_t* drop, int ncols) {
  p.drop_mode = 0; p.drop_bits = nullptr; p.drop_thresh = 65536u; p.drop_scale = 1.f;
  p.drop_seed[0] = p.drop_seed[1] = p.drop_off[0] = p.drop_off[1] = 0u;
  if (!drop || drop->mode == MIL_DROP_NONE || drop->p == 0.f) return 0;
  MIL_CHECK_ARG(drop->mode == MIL_DROP_BITS || drop->mode == MIL_DROP_PHILOX || drop->mode == MIL_DROP_PHILOX_DEV, "dropout: bad mode %d", drop->mode);
  MIL_CHECK_ARG(drop->p > 0.f && drop->p < 1.f, "dropout: p=%g must be in [0, 1)", (double)drop->p);
  MIL_CHECK_ARG(ncols % 32 == 0, "dropout: the dropped tensor must have a multiple of 32 columns (got %d)", ncols);
  MIL_CHECK_ARG(drop->mode == MIL_DROP_PHILOX || drop->keep_bits, "dropout: modes 1 and 3 need keep_bits");
  p.drop_mode = drop->mode;
  p.drop_bits = drop->keep_bits;
  p.drop_scale = 1.f / (1.f - drop->p);
  p.drop_thresh = (uint32_t)lrint((1.0 - (double)drop->p) * 65536.0);
  p.drop_seed[0] = (uint32_t)drop->seed; p.drop_seed[1] = (uint32_t)(drop->seed >> 32);
  p.drop_off[0] = (uint32_t)drop->offset; p.drop_off[1] = (uint32_t)(drop->offset >> 32);
  return 0;
}

__global__ void dropout_bits_kernel(int64_t words, int words_per_row, uint32_t thresh, uint32_t s0, uint32_t s1, uint32_t o0, uint32_t o1,
                                    const uint32_t* __restrict__ dev, uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= words) return;
  if (dev) { s0 = dev[0]; s1 = dev[1]; o0 = dev[2]; o1 = dev[3]; }
  const uint32_t seed[2] = {s0, s1}, off[2] = {o0, o1};
  out[i] = philox_keep_word((uint32_t)(i / words_per_row), (uint32_t)(i % words_per_row), thresh, seed, off);
}

static int split_weights(const float* w, int R, int K, uint8_t* img, int precision, cudaStream_t stream) {
  const int64_t items = (int64_t)R * (K / 8);
  const unsigned blocks = (unsigned)((items + 255) / 256);
  const float sc = prec_wscale(precision);
  if (precision == MIL_PREC_BF16X3) split_weights_kernel<false, true><<<blocks, 256, 0, stream>>>(w, R, K, img, sc);
  else if (precision == MIL_PREC_FP16X3) split_weights_kernel<true, true><<<blocks, 256, 0, stream>>>(w, R, K, img, sc);
  else if (precision == MIL_PREC_FP16) split_weights_kernel<true, false><<<blocks, 256, 0, stream>>>(w, R, K, img, sc);
  else split_weights_kernel<false, false><<<blocks, 256, 0, stream>>>(w, R, K, img, sc);
  MIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mil

using namespace mil;

extern "C" int mil_fused_num_partials(void) { return num_sms() + 1; }   // one spare record: the bulk copies of the final merge round up to 16 bytes

extern "C" void mil_profile_enable(int on) { g_profile = on != 0; }
// Synchronises the recorded event pairs; returns their count and writes the summed kernel time in ms.
extern "C" int mil_profile_collect(double* total_ms) {
  double tot = 0.0;
  int n = 0;
  for (auto& ev : g_prof_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) { tot += ms; ++n; }
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  g_prof_events.clear();
  if (total_ms) *total_ms = tot;
  return n;
}

extern "C" int mil_dropout_bits(int64_t rows, int ncols, const mil_dropout_t* drop, uint32_t* keep_bits_out, mil_stream_t stream_) {
  MIL_CHECK_ARG(rows > 0 && rows < (1ll << 31) && ncols > 0 && ncols % 32 == 0 && drop && keep_bits_out, "mil_dropout_bits: bad argument");
  MIL_CHECK_ARG((drop->mode == MIL_DROP_PHILOX || drop->mode == MIL_DROP_PHILOX_DEV) && drop->p > 0.f && drop->p < 1.f, "mil_dropout_bits: needs mode 2 or 3 and 0 < p < 1");
  FusedParams p;
  int rc;
  if ((rc = set_dropout(p, drop, ncols))) return rc;
  const int64_t words = rows * (ncols / 32);
  dropout_bits_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(words, ncols / 32, p.drop_thresh, p.drop_seed[0], p.drop_seed[1],
                                                                                         p.drop_off[0], p.drop_off[1],
                                                                                         p.drop_mode == 3 ? p.drop_bits : nullptr, keep_bits_out);
  MIL_LAUNCH_CHECK();
  return 0;
}

// workspace: [weight images][err, counter][trace stamps][tail-split flags (1 KB) | partial accumulators of the pair pipeline's tail split:
// fewer than num_sms / 2 partials of 128 x 512 fp32 (mil_fused2_sm100.cu)]
static size_t fused_ws_base_bytes(int D, int H, int Da) {
  return ((size_t)H * D + (size_t)Da * H) * 2 /*hi+lo*/ * sizeof(uint16_t) + 1024 + sizeof(int) * 4 + 64 + (16 * 16 + 2 * 1024) * sizeof(long long);
}
static size_t fused_ws_split_bytes() { return 256 + 1024 + (size_t)(num_sms() / 2) * 128 * HMAX * sizeof(float); }
extern "C" size_t mil_fused_workspace_bytes(int D, int H, int Da, int gated) {
  (void)gated;
  return fused_ws_base_bytes(D, H, Da) + fused_ws_split_bytes();
}

extern "C" int mil_abmil_fused_fwd_f32(const float* X, int64_t N, int D, int H, const float* W1, const float* b1, int act, const float* Wa,
                                       const float* ba, const float* Wb, const float* bb, int Da, int att_act, const float* wc, const float* bc,
                                       const uint8_t* keep, const float* Wp, int C, float* s_out, float* t_out, float* h_out, float* part,
                                       float* stats, float* pooled, float* rec_out, const float* Wcls, const float* bcls, int n_cls, float* logits,
                                       const mil_dropout_t* drop, void* ws, size_t ws_bytes, int ws_ready, int precision, mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  // bits 8..15 of `precision` select the pipeline: 0 = default (pair), MIL_PIPE_SINGLE, MIL_PIPE_PAIR; MHIMK_PIPELINE=1|2 overrides the default
  int pipeline = (precision >> 8) & 0xFF;
  precision &= 0xFF;
  if (pipeline == 0) {
    const char* e = getenv("MHIMK_PIPELINE");
    pipeline = e ? atoi(e) : 0;
    if (pipeline != 1 && pipeline != 2) pipeline = 2;
  }
  MIL_CHECK_ARG(pipeline == 1 || pipeline == 2, "mil_abmil_fused_fwd_f32: bad pipeline %d", pipeline);
  MIL_CHECK_ARG(mil_device_supported(), "mil_abmil_fused_fwd_f32: needs a compute-capability 10.x device (tcgen05/TMEM/TMA)");
  MIL_CHECK_ARG(X && W1 && b1 && Wa && wc && part && stats && pooled && ws, "mil_abmil_fused_fwd_f32: null argument");
  MIL_CHECK_ARG(N > 0 && N < (1ll << 31) - 256, "mil_abmil_fused_fwd_f32: N=%lld out of range", (long long)N);
  MIL_CHECK_ARG(H == HMAX, "mil_abmil_fused_fwd_f32: H=%d (only 512 is supported)", H);
  MIL_CHECK_ARG(D >= BK && D % BK == 0, "mil_abmil_fused_fwd_f32: D=%d must be a positive multiple of %d", D, BK);
  MIL_CHECK_ARG(Wb == nullptr && bb == nullptr, "mil_abmil_fused_fwd_f32: the gated branch is not fused yet; use the composed path");
  MIL_CHECK_ARG(Da == 128, "mil_abmil_fused_fwd_f32: Da=%d (only 128 is fused)", Da);
  MIL_CHECK_ARG(precision >= 0 && precision <= 3, "mil_abmil_fused_fwd_f32: bad precision %d", precision);
  MIL_CHECK_ARG(!t_out || (Wp && C >= 1 && C <= 4), "mil_abmil_fused_fwd_f32: t_out needs Wp and 1 <= C <= 4");
  MIL_CHECK_ARG((uintptr_t)X % 16 == 0 && (!h_out || (uintptr_t)h_out % 16 == 0), "mil_abmil_fused_fwd_f32: X / h_out must be 16-byte aligned");
  MIL_CHECK_ARG(ws_bytes >= mil_fused_workspace_bytes(D, H, Da, 0), "mil_abmil_fused_fwd_f32: workspace needs %zu bytes", mil_fused_workspace_bytes(D, H, Da, 0));

  uint8_t* w = (uint8_t*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  uint8_t* w1_img = w;                                        // [D/32][hi|lo][H x 64 B]
  uint8_t* wa_img = w1_img + (size_t)H * D * 4;               // [H/32][hi|lo][Da x 64 B]
  int* err = (int*)(wa_img + (size_t)Da * H * 4);     // err[0] = pipeline error code, err[1] = finished-CTA counter
  int rc;
  MIL_CHECK_ARG(!logits || (Wcls && n_cls >= 1 && n_cls <= 64), "mil_abmil_fused_fwd_f32: logits needs Wcls and 1 <= n_cls <= 64");
  if (!ws_ready) {   // 16-bit operand images of the weights: reusable across calls until the weights change (ws_ready = 1)
    MIL_CHECK_ARG((uintptr_t)W1 % 16 == 0 && (uintptr_t)Wa % 16 == 0, "mil_abmil_fused_fwd_f32: weights must be 16-byte aligned");
    if (pipeline == 2) {
      if ((rc = pair_build_images(W1, H, D, Wa, Da, w1_img, wa_img, precision, stream))) return rc;
    } else {
      if ((rc = split_weights(W1, H, D, w1_img, precision, stream))) return rc;
      if ((rc = split_weights(Wa, Da, H, wa_img, precision, stream))) return rc;
    }
    MIL_CUDA(cudaMemsetAsync(err, 0, 4 * sizeof(int), stream));
  }
  long long* trace_base = (long long*)(((uintptr_t)(err + 4) + 63) & ~(uintptr_t)63);
  int* split_flags = (int*)(((uintptr_t)(trace_base + 16 * 16 + 2 * 1024) + 255) & ~(uintptr_t)255);
  if (!ws_ready) MIL_CUDA(cudaMemsetAsync(split_flags, 0, 1024, stream));
  CUtensorMap mx;
  if (pipeline == 1 && (rc = make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, (uint64_t)N, (uint64_t)D, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;

  FusedParams p;
  p.N = N; p.D = D; p.nout = H; p.Da = Da; p.act = act; p.att_act = att_act;
  p.b1 = b1; p.ba = ba; p.wc = wc; p.bc = bc; p.keep = keep; p.Wp = Wp; p.C = t_out ? C : 0;
  p.s_out = s_out; p.t_out = t_out; p.h_out = h_out; p.part = part; p.c_out = nullptr; p.ldc = 0; p.err = err; p.dbg = debug_mask(); p.w1_img = w1_img; p.wa_img = wa_img;
  p.stats = stats; p.pooled = pooled; p.rec_out = rec_out; p.counter = (unsigned int*)(err + 1); p.Wcls = Wcls; p.bcls = bcls; p.n_cls = n_cls; p.logits = logits;
  p.trace = getenv("MHIMK_TRACE") ? trace_base : nullptr;
  p.trace_cta = getenv("MHIMK_TRACE_CTA") ? atoi(getenv("MHIMK_TRACE_CTA")) : 0;
  p.split_flags = split_flags; p.split_buf = (float*)((uint8_t*)split_flags + 1024);
  p.w1_inv = p.wa_inv = 1.f / prec_wscale(precision);
  if ((rc = set_dropout(p, drop, H))) return rc;
  if (pipeline == 2) return pair_fused_launch(X, p, precision, stream);
  const int64_t n_tiles = (N + BM - 1) / BM;
  const int grid = grid_override((int)(n_tiles < num_sms() ? n_tiles : num_sms()));
  return dispatch_fused(precision, MODE_FUSED, mx, p, grid, stream);
}

extern "C" int mil_pair_plan_item(int64_t N, int D, int precision, int pair, int i, int64_t* out9) {
  MIL_CHECK_ARG(out9 && N > 0 && D >= 32 && D % 32 == 0 && precision >= 0 && precision <= 3 && pair >= 0, "mil_pair_plan_item: bad arguments");
  return pair_plan_item(N, D, precision, pair, i, out9);
}

extern "C" int mil_umma_selftest_f32(const float* A, const float* B, float* C, int M, int N, int K, int precision, void* ws, size_t ws_bytes,
                                     mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(mil_device_supported(), "mil_umma_selftest_f32: needs a compute-capability 10.x device");
  MIL_CHECK_ARG(A && B && C && ws && M > 0, "mil_umma_selftest_f32: null argument");
  MIL_CHECK_ARG((N == 64 || N == 128 || N == 256 || N == 512) && K >= BK && K % BK == 0, "mil_umma_selftest_f32: unsupported N=%d K=%d", N, K);
  MIL_CHECK_ARG(precision >= 0 && precision <= 3, "mil_umma_selftest_f32: bad precision");
  MIL_CHECK_ARG(ws_bytes >= (size_t)N * K * 4 + 1024, "mil_umma_selftest_f32: workspace needs %zu bytes", (size_t)N * K * 4 + 1024);
  MIL_CHECK_ARG(K % 32 == 0, "mil_umma_selftest_f32: K must be a multiple of 32");
  uint8_t* b_img = (uint8_t*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  int* err = (int*)(b_img + (size_t)N * K * 4);
  int rc;
  MIL_CHECK_ARG((uintptr_t)B % 16 == 0, "mil_umma_selftest_f32: B must be 16-byte aligned");
  if ((rc = split_weights(B, N, K, b_img, precision, stream))) return rc;
  MIL_CUDA(cudaMemsetAsync(err, 0, sizeof(int), stream));
  CUtensorMap mx;
  if ((rc = make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, (uint64_t)M, (uint64_t)K, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  FusedParams p;
  p.N = M; p.D = K; p.nout = N; p.Da = 128; p.act = MIL_ACT_NONE; p.att_act = MIL_ACT_NONE;
  p.b1 = nullptr; p.ba = nullptr; p.wc = nullptr; p.bc = nullptr; p.keep = nullptr; p.Wp = nullptr; p.C = 0;
  p.s_out = nullptr; p.t_out = nullptr; p.h_out = nullptr; p.part = nullptr; p.c_out = C; p.ldc = N; p.err = err; p.dbg = debug_mask(); p.w1_img = b_img; p.wa_img = b_img; p.trace = nullptr;
  p.stats = nullptr; p.pooled = nullptr; p.rec_out = nullptr; p.counter = nullptr; p.Wcls = nullptr; p.bcls = nullptr; p.n_cls = 0; p.logits = nullptr;
  set_dropout(p, nullptr, N);
  p.w1_inv = p.wa_inv = 1.f / prec_wscale(precision);
  const int64_t n_tiles = ((int64_t)M + BM - 1) / BM;
  const int grid = grid_override((int)(n_tiles < num_sms() ? n_tiles : num_sms()));
  return dispatch_fused(precision, MODE_STORE, mx, p, grid, stream);
}

extern "C" size_t mil_linear_tc_workspace_bytes(int N, int K) { return (size_t)N * K * 4 + 1024 + 64; }

extern "C" int mil_linear_act_tc_ld_f32(const float* X, int64_t ldx, int64_t M, int K, const float* W, const float* bias, int N, int act, float* pre_out,
                                        float* Y, int64_t ldy, const mil_dropout_t* drop, void* ws, size_t ws_bytes, int ws_ready, int precision,
                                        mil_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MIL_CHECK_ARG(mil_device_supported(), "mil_linear_act_tc_f32: needs a compute-capability 10.x device");
  MIL_CHECK_ARG(X && W && Y && ws && M > 0 && M < (1ll << 31) - 256, "mil_linear_act_tc_f32: bad argument");
  MIL_CHECK_ARG(ldx >= K && ldx % 4 == 0 && ldy >= N && ldy % 4 == 0, "mil_linear_act_tc_f32: leading dimensions must cover the rows and be multiples of 4");
  MIL_CHECK_ARG(N >= 64 && N <= 512 && N % 64 == 0 && (N <= 256 || N == 512), "mil_linear_act_tc_f32: N=%d must be 64,128,192,256 or 512", N);
  MIL_CHECK_ARG(K >= BK && K % BK == 0, "mil_linear_act_tc_f32: K=%d must be a positive multiple of %d", K, BK);
  MIL_CHECK_ARG(precision >= 0 && precision <= 3, "mil_linear_act_tc_f32: bad precision");
  MIL_CHECK_ARG((uintptr_t)X % 16 == 0 && (uintptr_t)W % 16 == 0 && (uintptr_t)Y % 16 == 0 && (!pre_out || (uintptr_t)pre_out % 16 == 0),
                "mil_linear_act_tc_f32: pointers must be 16-byte aligned");
  MIL_CHECK_ARG(ws_bytes >= mil_linear_tc_workspace_bytes(N, K), "mil_linear_act_tc_f32: workspace needs %zu bytes", mil_linear_tc_workspace_bytes(N, K));
  uint8_t* w_img = (uint8_t*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  int* err = (int*)(w_img + (size_t)N * K * 4);
  int rc;
  if (!ws_ready) {
    if ((rc = split_weights(W, N, K, w_img, precision, stream))) return rc;
    MIL_CUDA(cudaMemsetAsync(err, 0, 4 * sizeof(int), stream));
  }
  CUtensorMap mx;
  if ((rc = make_map_2d_ld(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  FusedParams p;
  p.N = M; p.D = K; p.nout = N; p.Da = 128; p.act = act; p.att_act = MIL_ACT_NONE;
  p.b1 = bias; p.ba = nullptr; p.wc = nullptr; p.bc = nullptr; p.keep = nullptr; p.Wp = nullptr; p.C = 0;
  p.s_out = nullptr; p.t_out = nullptr; p.h_out = pre_out; p.part = nullptr; p.c_out = Y; p.ldc = ldy; p.err = err; p.dbg = 0;
  p.w1_img = w_img; p.wa_img = w_img; p.trace = nullptr;
  p.stats = nullptr; p.pooled = nullptr; p.rec_out = nullptr; p.counter = nullptr; p.Wcls = nullptr; p.bcls = nullptr; p.n_cls = 0; p.logits = nullptr;
  if ((rc = set_dropout(p, drop, N))) return rc;
  p.w1_inv = p.wa_inv = 1.f / prec_wscale(precision);
  const int64_t n_tiles = (M + BM - 1) / BM;
  int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  // Column split: with at most half of the SMs' worth of row tiles a 512-wide output is computed as two 256-column blocks by separate CTAs
  // (half the MMA time and weight traffic per CTA; the bag tile is read and converted twice, the second read is an L2 hit).
  static const bool nocolsplit = getenv("MHIMK_NOCOLSPLIT") && atoi(getenv("MHIMK_NOCOLSPLIT"));
  if (N == 512 && 2 * n_tiles <= num_sms() && !nocolsplit) { p.nout = 256; p.nblk = 2; grid = (int)(2 * n_tiles); }
  return dispatch_fused(precision, MODE_STORE, mx, p, grid, stream);
}

extern "C" int mil_linear_act_tc_f32(const float* X, int64_t M, int K, const float* W, const float* bias, int N, int act, float* pre_out,
                                     float* Y, const mil_dropout_t* drop, void* ws, size_t ws_bytes, int ws_ready, int precision,
                                     mil_stream_t stream_) {
  return mil_linear_act_tc_ld_f32(X, K, M, K, W, bias, N, act, pre_out, Y, N, drop, ws, ws_bytes, ws_ready, precision, stream_);
}
