// Pair pipeline of the fused ABMIL forward (sm_100a): two CTAs of a thread-block cluster (one TPC) share every
// tcgen05.mma (cta_group::2, M = 128 -> 64 bag rows per CTA).  A 64-row x 512-column fp32 accumulator is 256 TMEM
// columns, so TMEM holds TWO of them: GEMM1 of tile t+1 runs while the epilogue warps work on tile t, and the HBM stream of
// the bag never pauses (the single-CTA pipeline of mil_fused_sm100.cu owns all 512 columns and has to serialise).
//
//   per CTA (rank r of the pair), per 128-row pair tile (its rows: tile * 128 + r * 64 .. + 64):
//   HBM --TMA(SW128)--> fp32 staging [64 x 32] --converter warps--> 16-bit hi(/lo) A tiles (UMMA K-major SW64, 4 KB)
//   L2  --TMA---------> W1 image tiles: this CTA's half (128 rows) of each 256-row N block; Wa image tiles (64 rows)
//   leader CTA, warp 13 (converged, one elected lane): tcgen05.mma.cta_group::2  pre[128 x 256] x 2 N blocks -> TMEM buffer (tile & 1)   (GEMM1)
//   epilogue warps (both CTAs): h = act(pre + b1) -> 16-bit A2 tiles; h kept in TMEM / registers
//   leader CTA, warp 15: tcgen05.mma.cta_group::2  u[128 x 128] = h Wa^T -> the first 64 columns of the same buffer          (GEMM2)
//   epilogue warps: s = wc . f(u + ba) + bc, online softmax over rows, p += e^{s-m} h (warp-shuffle transposes)
//
// TMEM layout of a pair MMA with M = 128 ("2x2", cute::UMMA::tmem_frg_2sm): rows 0..63 of the CTA on lanes 0..63 for
// columns n < N/2 and on lanes 64..127 for n >= N/2, N/2 TMEM columns per N block.
//
// Barriers: every "full" barrier (operands ready, accumulator drained) lives in the leader CTA and is arrived on by both
// CTAs (remote mbarrier.arrive / cta_group::2 TMA complete_tx); every "empty"/"done" barrier is signalled in both CTAs by
// one multicast tcgen05.commit.  All waits trap after ~2 s instead of hanging the GPU.
#include <stdlib.h>

#include "mil_umma.cuh"

namespace mil {
namespace pairk {

constexpr int BMP = 128;                      // rows per pair tile (UMMA M)
constexpr int BMC = 64;                       // rows per CTA
constexpr int BK = 32;                        // K elements per stage (two UMMA K = 16 steps)
constexpr int X_SLOT = BMC * BK * 4;          // 8192   fp32 staging slot
constexpr int A_OP = BMC * BK * 2;            // 4096   one 16-bit A tile [64 x 32]
constexpr int B_OP = 2 * 128 * BK * 2;        // 16384  this CTA's 128 rows of both 256-row N blocks: [blk 0 | blk 1]
constexpr int A2_OP = A_OP;                   // 4096   GEMM2 A tile [64 x 32]
constexpr int B2_OP = 64 * BK * 2;            // 4096   this CTA's 64 rows of Wa
constexpr int G2S = 4;                        // GEMM2 A2 slots (one per epilogue-warp column strip)
constexpr int G2B_BUF = 16384;                // one Wa load group (GS chunks x NOP x 4 KB); two buffers

constexpr int NUM_THREADS = 512;
// Warp roles.  The epilogue of tile t runs concurrently with the operand pipeline of tile t+1.  The short latency-critical loops
// (producers, MMA issue, converters) have the high warp ids and the ALU-heavy epilogue warps the low ones because the SM's warp
// arbiter is reported to favour the highest id among eligible warps; measured here the order made no difference (the control
// loops are bound by their own issue latency, ~300-450 cycles per iteration, not by lost arbitration).
constexpr int EPI_WARP0 = 0;                  // warps 0..7 epilogue
constexpr int CONV_WARP0 = 8;                 // warps 8..11 converters
constexpr int W_X = 12, W_MMA1 = 13, W_W1 = 14, W_MMA2 = 15;   // bag TMA; GEMM1 issue; TMEM alloc + W1 TMA; Wa TMA + GEMM2 issue
constexpr int MISC_BYTES = 512 + 1024 + 4096 + 16 + (HMAX + 256) * 4;

// k-step stamps of CTA 0 (MHIMK_TRACE=1): series x first 64 k-steps, see tools/trace_ksteps.py.  Compiled in only with
// -DMIL_KSTAMP (KSTAMP=1 csrc/build.sh): even a predicated-off stamp costs the control loops issue slots.
__device__ __forceinline__ void kstamp(const FusedParams& p, int series, uint32_t it) {
#ifdef MIL_KSTAMP
  if (p.trace && blockIdx.x == 0 && it < 64) p.trace[1024 + series * 64 + it] = clock64();
#endif
}

// Work items of one pair.  Whole waves of tiles go round-robin over the pairs (tile = pair + i * n_pairs); the tiles of a partly filled wave
// are split by K range over otherwise idle pairs (FusedParams::split_*): the pairs S g .. S g + S - 1 share tile split_full * n_pairs + g, pair
// S g ("owner") runs the tile's epilogue after adding the helper's dumped accumulator (S = 2: one partial, deterministic).
// The helper does its part FIRST and the owner its part LAST: the exchange (dump 128 KB per CTA through L2 at ~32 B/clk, barrier, release flag:
// 6.5-10 k cycles, measured) is then long over when the owner's last epilogue asks for the partial.  (Both last: the dump sits on the kernel's
// tail.  Both first: the owner's wait delays the release of its TMEM buffer and stalls GEMM1 two items later; both measured, both gain nothing.)
// A helper's item has no GEMM2: the GEMM2-side barriers count GEMM2 tiles (item index - 1 on a helper pair), the accumulator barriers count items.
struct WorkItem {
  int64_t tile;
  int kb, ke;       // pipeline stages [kb, ke) of the tile's K loop
  int kind;         // 0 whole tile, 1 owner of a split tile, 2 helper
  int pidx;         // helper: index of its partial; owner: index of its first helper's partial (S - 1 consecutive)
};
__host__ __device__ __forceinline__ int work_count(const FusedParams& p, int64_t pair_id, int64_t n_pairs, int64_t n_tiles) {
  if (p.split_s <= 1) return pair_id < n_tiles ? (int)((n_tiles - pair_id + n_pairs - 1) / n_pairs) : 0;
  return p.split_full + (pair_id < (int64_t)p.split_rem * p.split_s ? 1 : 0);
}
__host__ __device__ __forceinline__ WorkItem work_item(const FusedParams& p, int i, int64_t pair_id, int64_t n_pairs, int KST) {
  WorkItem w;
  const bool has_split = p.split_s > 1 && pair_id < (int64_t)p.split_rem * p.split_s;
  const int part = has_split ? (int)pair_id % p.split_s : 0;
  const int split_pos = part ? 0 : p.split_full;           // helper: its first item; owner: its last
  if (!has_split || i != split_pos) {
    w.tile = pair_id + (int64_t)(i - (part ? 1 : 0)) * n_pairs; w.kb = 0; w.ke = KST; w.kind = 0; w.pidx = 0;
    return w;
  }
  const int grp = (int)pair_id / p.split_s;
  w.tile = (int64_t)p.split_full * n_pairs + grp;
  w.kb = KST * part / p.split_s;
  w.ke = KST * (part + 1) / p.split_s;
  w.kind = part ? 2 : 1;
  w.pidx = grp * (p.split_s - 1) + (part ? part - 1 : 0);
  return w;
}

template <int NPROD, bool FP16, int NST, int XS, int KSUB, int ACT, int ATT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
mil_fused2_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1,
                  const __grid_constant__ CUtensorMap mapWa, const FusedParams p) {
  constexpr bool LO = NPROD == 3;
  constexpr int NOP = LO ? 2 : 1;                              // operand tiles per stage (hi, lo)
  // one pipeline stage = KSUB sub-steps of 32 K elements: the fixed cost of a stage hand-over (two mbarrier round trips, a commit,
  // ~300 issue-latency-bound cycles in the MMA warp) is amortised over KSUB x 4 (x3) MMAs
  constexpr uint32_t A_SUB = NOP * A_OP, B_SUB = NOP * B_OP, A_STAGE = KSUB * A_SUB, B_STAGE = KSUB * B_SUB, G2A_STAGE = NOP * A2_OP;
  constexpr int GS = 4 / NOP;                                  // GEMM2 k-chunks per Wa load (16 KB), NG groups per tile
  constexpr int NG = 16 / GS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                                          // XS x 8 KB
  uint8_t* sA = sX + XS * X_SLOT;                              // NST x A_STAGE
  uint8_t* sB = sA + NST * A_STAGE;                            // NST x B_STAGE
  uint8_t* sA2 = sB + NST * B_STAGE;                           // G2S x G2A_STAGE
  uint8_t* sB2 = sA2 + G2S * G2A_STAGE;                        // 2 x 16 KB Wa group buffers
  uint8_t* sMisc = sB2 + 2 * G2B_BUF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMisc);         // barrier block (512 B)
  float* s_part = reinterpret_cast<float*>(sMisc + 512);       // [4 strips][64 rows] partial attention logits
  float* t_part = s_part + 256;                                // [4 strips][64 rows][4] partial t (also grid_finalize scratch)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sMisc + 512 + 1024 + 4096);
  float* c_b1 = reinterpret_cast<float*>(sMisc + 512 + 1024 + 4096 + 16);   // [512] feature bias
  float* c_ba = c_b1 + HMAX;                                                 // [128] attention bias
  float* c_wc = c_ba + 128;                                                  // [128] attention output weights
  float* p_acc = reinterpret_cast<float*>(sX);                               // [8 warps][4 chunks][32 lanes]: hand-over buffer of the CTA tail (the staging ring is idle then)

  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  // FULL[s] (leader): 4 converter warps of the pair + the W1 producer's expect_tx arrival + both CTAs' W1 bytes; G2FULL[r] likewise
  constexpr int B_XFULL = 0, B_XEMPTY = B_XFULL + XS, B_FULL = B_XEMPTY + XS, B_EMPTY = B_FULL + NST,
                B_ACCFULL = B_EMPTY + NST, B_ACCEMPTY = B_ACCFULL + 2, B_TAILFREE = B_ACCEMPTY + 2, B_UFULL = B_TAILFREE + 2,
                B_G2AFULL = B_UFULL + 1, B_G2AEMPTY = B_G2AFULL + G2S, B_G2BFULL = B_G2AEMPTY + G2S, B_G2BEMPTY = B_G2BFULL + 2,
                B_FIN = B_G2BEMPTY + 2,
                B_COUNT = B_FIN + 2;
  static_assert(B_COUNT * 8 <= 512, "barrier block overflow");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                    // 0 = leader (issues every MMA of the pair)
  auto LEADER = [&](int i) { return mapa_shared(BAR(i), 0); };  // cluster address of barrier i in the leader CTA
  auto ARRIVE_LEADER = [&](int i) {                            // the leader takes the plain shared::cta path for its own barriers
    if (rank == 0) mbar_arrive(BAR(i));
    else mbar_arrive_cluster(mapa_shared(BAR(i), 0));
  };
  if (p.trace && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[256 + 2 * blockIdx.x] = (long long)gt;
    if (blockIdx.x == (unsigned)p.trace_cta) p.trace[238] = clock64();
  }
  const int KS = p.D / BK;                                     // GEMM1 32-wide k sub-steps per tile
  const int KST = KS / KSUB;                                   // pipeline stages per tile
  constexpr int NCH2 = HMAX / BK;                              // GEMM2 k-steps per tile (16)
  const int64_t n_tiles = (p.N + BMP - 1) / BMP;
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_it = work_count(p, pair_id, n_pairs, n_tiles);    // work items of this pair (whole tiles, then at most one split item)

  if (threadIdx.x == 0) {
    for (int i = 0; i < XS; ++i) { mbar_init(BAR(B_XFULL + i), 1); mbar_init(BAR(B_XEMPTY + i), LO ? 4 : 2); }
    for (int i = 0; i < NST; ++i) { mbar_init(BAR(B_FULL + i), (LO ? 8 : 4) * KSUB + 1); mbar_init(BAR(B_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_ACCFULL + i), 1); mbar_init(BAR(B_ACCEMPTY + i), 16); mbar_init(BAR(B_TAILFREE + i), 8); }
    mbar_init(BAR(B_UFULL), 1);
    for (int i = 0; i < G2S; ++i) { mbar_init(BAR(B_G2AFULL + i), 4); mbar_init(BAR(B_G2AEMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_G2BFULL + i), 1); mbar_init(BAR(B_G2BEMPTY + i), 1); mbar_init(BAR(B_FIN + i), 1); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == W_X && lane == 0) { tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapW1); tma_prefetch_desc(&mapWa); }
  if (warp == W_W1) tmem_alloc_pair(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                          // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for the compiler

  if (warp >= W_X) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // Control warps run their loops CONVERGED (all 32 lanes wait on the barriers and compute the uniform operands; one
    // elected lane issues the TMA / MMA / commit).  A lone diverged lane made ptxas move every descriptor into uniform
    // registers through an ELECT + 5 x R2UR waterfall per instruction: ~700 cycles per k-step, 3x the MMA time.
    if (warp == W_X) {
      // ===================== bag producer: HBM -> fp32 staging (this CTA's 64 rows of the pair tile) =====================
      // The staging ring (XS x 8 KB) cannot hold an HBM round trip of the bag stream (~1800 cycles x 32 B/cycle = 58 KB), so the
      // tiles are first pulled into L2 PF sub-steps ahead (cp.async.bulk.prefetch.tensor); the staged load then sees L2 latency.
      uint32_t s = 0, ph = 1, xit = 0;
      const uint32_t sx0 = smem_u32(sX);
#ifdef MIL_PROBE
      const int PF = ((p.dbg >> 16) & 0xFF) ? ((p.dbg >> 16) & 0xFF) : 24;       // MHIMK_DEBUG bits 16..23 override the L2 prefetch distance (PROBE=1 builds only)
#else
      constexpr int PF = 24;                                                     // 8 ... 96 measured: no sensitivity (profiles/round2_summary.md)
#endif
      int pi = 0;
      WorkItem pw = work_item(p, 0, pair_id, n_pairs, KST);
      int pks = pw.kb * KSUB;
      auto prefetch_next = [&]() {                                               // one box further down this CTA's stream
        if (pi < n_it) {
          if (elect_one()) tma_prefetch_2d(&mapX, pks * BK, (int)(pw.tile * BMP + rank * BMC));
          __syncwarp();
          if (++pks == pw.ke * KSUB && ++pi < n_it) { pw = work_item(p, pi, pair_id, n_pairs, KST); pks = pw.kb * KSUB; }
        }
      };
      if (!(p.dbg & 2))
        for (int i = 0; i < PF; ++i) prefetch_next();
      for (int wi = 0; wi < n_it; ++wi) {
        const WorkItem w = work_item(p, wi, pair_id, n_pairs, KST);
        const int row0 = (int)(w.tile * BMP + rank * BMC);
        for (int ks = w.kb * KSUB; ks < w.ke * KSUB; ++ks) {
          if (!(p.dbg & 2)) prefetch_next();
          mbar_wait(BAR(B_XEMPTY + s), ph, p.err, 1);
          if (lane == 0) kstamp(p, 6, xit++);
          if (elect_one()) {
            if (p.dbg & 2) mbar_arrive(BAR(B_XFULL + s));                          // timing attribution: no bag traffic
            else {
              mbar_expect_tx(BAR(B_XFULL + s), X_SLOT);
              tma_load_2d(sx0 + s * X_SLOT, &mapX, BAR(B_XFULL + s), ks * BK, row0);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)XS) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == W_W1) {
      // ===================== W1 producer: L2 -> this CTA's half of the B operand of every k-step =====================
      // Image rows are 256 B; one stage of one CTA is KSUB x NOP x 16 KB contiguous: per sub-step [hi: N block 0 | N block 1][lo: ...].
      uint32_t s = 0, ph = 1, wit = 0;
      const uint32_t full0 = LEADER(B_FULL), sb0 = smem_u32(sB);
      for (int wi = 0; wi < n_it; ++wi) {
        const WorkItem w = work_item(p, wi, pair_id, n_pairs, KST);
        for (int kst = w.kb; kst < w.ke; ++kst) {
          mbar_wait(BAR(B_EMPTY + s), ph, p.err, 2);
          if (lane == 0) kstamp(p, 0, wit++);
          if (elect_one()) {
            if (p.dbg & 1) { if (rank == 0) mbar_arrive(BAR(B_FULL + s)); }        // timing attribution: no W1 traffic
            else {
              if (rank == 0) mbar_expect_tx(BAR(B_FULL + s), 2 * B_STAGE);       // both CTAs' bytes land on the leader's barrier
              tma_load_2d_pair(sb0 + s * B_STAGE, &mapW1, full0 + 8u * s, 0, (int)((kst * 2 + rank) * (B_STAGE / 256)));
            }
          }
          __syncwarp();
          if (++s == (uint32_t)NST) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == W_MMA1 && rank == 0) {
      // ===================== GEMM1 issuer (leader CTA) =====================
      const uint32_t idesc1 = make_idesc(FP16, 256, BMP);
      const uint32_t sa0 = smem_u32(sA), sb0 = smem_u32(sB);
      uint32_t s = 0, ph = 0, tl = 0, iit = 0, sq = 0, phq = 0, gi = 0;
      const uint32_t qd = ((uint32_t)p.dbg >> 12) & 15u;
      for (int wi = 0; wi < n_it; ++wi, ++tl) {
        const WorkItem w = work_item(p, wi, pair_id, n_pairs, KST);
        const uint32_t b1 = tl & 1;
        mbar_wait(BAR(B_ACCEMPTY + b1), ((tl >> 1) & 1) ^ 1, p.err, 4);
        tc_fence_after();
        if (lane == 0) trace_stamp(p, tl, 0);
        const uint32_t d0 = tmem + b1 * 256;
        for (int kst = w.kb; kst < w.ke; ++kst) {
          if (lane == 0) kstamp(p, 2, iit);
          // optional queue-depth limit (MHIMK_DEBUG bits 12..15 = qd): at most qd stages queued in the tensor pipe, so that GEMM2's
          // short MMAs of the previous tile (other warp) do not wait behind a full ring of GEMM1 work
          if (qd && gi >= qd) {
            mbar_wait(BAR(B_EMPTY + sq), phq, p.err, 5);
            if (++sq == (uint32_t)NST) { sq = 0; phq ^= 1; }
          }
          ++gi;
          mbar_wait(BAR(B_FULL + s), ph, p.err, 6);
          if (lane == 0) kstamp(p, 1, iit++);
          tc_fence_after();
          const uint32_t a0 = sa0 + s * A_STAGE, b0 = sb0 + s * B_STAGE;
          const uint64_t ah0 = make_desc_sw64(a0), bh0 = make_desc_sw64(b0), al0 = make_desc_sw64(a0 + A_OP), bl0 = make_desc_sw64(b0 + B_OP);
          if (elect_one()) {
#pragma unroll
            for (int sub = 0; sub < KSUB; ++sub) {
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16) {
                if (p.dbg & 4) break;                                              // timing attribution: no GEMM1 MMAs
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                  // descriptor start addresses are in 16-byte units: +2 per K = 16 step (32 B), +512 per N block (8 KB)
                  const uint64_t oa = (uint64_t)(sub * (A_SUB / 16) + k16 * 2), ob = (uint64_t)(sub * (B_SUB / 16) + blk * 512 + k16 * 2);
                  const uint32_t acc = ((kst - w.kb) | sub | k16) ? 1u : 0u;
                  umma_f16_pair(d0 + blk * 128, ah0 + oa, bh0 + ob, idesc1, acc);
                  if (LO) {
                    umma_f16_pair(d0 + blk * 128, al0 + oa, bh0 + ob, idesc1, 1u);
                    umma_f16_pair(d0 + blk * 128, ah0 + oa, bl0 + ob, idesc1, 1u);
                  }
                }
              }
            }
            umma_commit_pair(BAR(B_EMPTY + s));
            if (kst == w.ke - 1) umma_commit_pair(BAR(B_ACCFULL + b1));
          }
          __syncwarp();
          if (++s == (uint32_t)NST) { s = 0; ph ^= 1; }
        }
        if (lane == 0) trace_stamp(p, tl, 1);
      }
    } else if (warp == W_MMA2) {
      // ===================== Wa producer (both CTAs) + GEMM2 issuer (leader) =====================
      // GEMM2 of tile t runs against the epilogue of tile t while GEMM1 is on tile t+1.  Wa arrives in groups of GS k-chunks
      // (16 KB per CTA: this CTA's 64 rows of each chunk), double-buffered; group G+1 is requested when group G starts.
      const uint32_t idesc2 = make_idesc(FP16, 128, BMP);
      const uint32_t bfull0 = LEADER(B_G2BFULL), sa20 = smem_u32(sA2), sb20 = smem_u32(sB2);
      const uint32_t skip = (n_it && work_item(p, 0, pair_id, n_pairs, KST).kind == 2) ? 1u : 0u;   // a helper's split item (its first) has no GEMM2
      const uint32_t T = (uint32_t)n_it - skip;
      const uint32_t total = T * NG;
      auto load_group = [&](uint32_t G) {                      // G-th group of this CTA's sequence (groups repeat every NG)
        const uint32_t bf = G & 1;
        mbar_wait(BAR(B_G2BEMPTY + bf), ((G >> 1) & 1) ^ 1, p.err, 3);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(BAR(B_G2BFULL + bf), 2 * G2B_BUF);
          tma_load_2d_pair(sb20 + bf * G2B_BUF, &mapWa, bfull0 + 8u * bf, 0, (int)(((G % NG) * 2 + rank) * (G2B_BUF / 256)));
        }
        __syncwarp();
      };
      if (total) load_group(0);
      uint32_t G = 0;
      for (uint32_t tl = 0; tl < T; ++tl) {                  // tl: GEMM2 tile index; its accumulator buffer follows the ITEM index tl + skip
        const uint32_t b2 = (tl + skip) & 1;
        if (rank == 0) {
          // (tl >> 1) = earlier GEMM2 tiles on the same buffer, with or without a skipped first item
          mbar_wait(BAR(B_TAILFREE + b2), (tl >> 1) & 1, p.err, 7);             // u's columns are vacated (implies GEMM1 of the tile is complete)
          tc_fence_after();
          if (lane == 0) trace_stamp(p, tl + skip, 2);
        }
        const uint32_t d = tmem + b2 * 256;
        for (int g = 0; g < NG; ++g, ++G) {
          if (G + 1 < total) load_group(G + 1);
          if (rank != 0) continue;
          const uint32_t bf = G & 1;
          mbar_wait(BAR(B_G2BFULL + bf), (G >> 1) & 1, p.err, 8);
#pragma unroll
          for (int sg = 0; sg < GS; ++sg) {
            const uint32_t c = (uint32_t)(g * GS + sg), strip = c & 3, j = c >> 2;
            mbar_wait(BAR(B_G2AFULL + strip), (tl * 4 + j) & 1, p.err, 9);
            tc_fence_after();
            const uint32_t a0 = sa20 + strip * G2A_STAGE, b0 = sb20 + bf * G2B_BUF + sg * NOP * B2_OP;
            const uint64_t ah0 = make_desc_sw64(a0), bh0 = make_desc_sw64(b0), al0 = make_desc_sw64(a0 + A2_OP), bl0 = make_desc_sw64(b0 + B2_OP);
            if (elect_one()) {
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16) {
                const uint64_t ah = ah0 + (uint64_t)(k16 * 2), bh = bh0 + (uint64_t)(k16 * 2);
                umma_f16_pair(d, ah, bh, idesc2, (c | (uint32_t)k16) ? 1u : 0u);
                if (LO) {
                  umma_f16_pair(d, al0 + (uint64_t)(k16 * 2), bh, idesc2, 1u);
                  umma_f16_pair(d, ah, bl0 + (uint64_t)(k16 * 2), idesc2, 1u);
                }
              }
              umma_commit_pair(BAR(B_G2AEMPTY + strip));
              if (sg == GS - 1) umma_commit_pair(BAR(B_G2BEMPTY + bf));
              if (sg == GS - 1 && g == NG - 1) umma_commit_pair(BAR(B_UFULL));
            }
            __syncwarp();
          }
        }
        if (rank == 0 && lane == 0) trace_stamp(p, tl + skip, 3);
      }
    }
  } else if (warp >= CONV_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // ===================== converters: fp32 staging -> 16-bit hi/lo A tiles =====================
    if (LO) {
      // 3-product mode (768 MMA cycles per sub-step): all four warps work on every sub-step, half a row (16 values) per thread.
      // No group alternation, so any ring size is sound for the parity waits.
      const int tid = threadIdx.x - CONV_WARP0 * 32, row = tid >> 1, hf = tid & 1;
      const uint32_t row_off = (uint32_t)(row >> 3) * 512u + (uint32_t)(row & 7) * 64u, sw = (uint32_t)(row >> 1) & 3u;
      uint32_t it = 0;
      for (int wi = 0; wi < n_it; ++wi) {
        const WorkItem w = work_item(p, wi, pair_id, n_pairs, KST);
        for (int ks = w.kb * KSUB; ks < w.ke * KSUB; ++ks, ++it) {
          const uint32_t xs = it % XS, xph = (it / XS) & 1;
          const uint32_t st = it / KSUB, sub = it % KSUB, s = st % NST, ph = (st / NST) & 1;
          mbar_wait(BAR(B_XFULL + xs), xph, p.err, 10);
          if (tid == 0) kstamp(p, 5, it);
#ifdef MIL_PROBE
          if (p.dbg & 8) {                                      // timing attribution (PROBE=1 builds only): hand-shakes only, no conversion work
            mbar_wait(BAR(B_EMPTY + s), ph ^ 1, p.err, 11);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { ARRIVE_LEADER(B_FULL + s); mbar_arrive(BAR(B_XEMPTY + xs)); }
            continue;
          }
#endif
          float x[16];
          const uint32_t src = smem_u32(sX + xs * X_SLOT) + (uint32_t)row * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t a = src + ((((uint32_t)(4 * hf + j)) ^ ((uint32_t)row & 7u)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[4 * j]), "=f"(x[4 * j + 1]), "=f"(x[4 * j + 2]), "=f"(x[4 * j + 3]) : "r"(a));
          }
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            hi[i] = pack_hi<FP16>(x[2 * i], x[2 * i + 1]);
            lo[i] = pack_lo<FP16>(x[2 * i], x[2 * i + 1], hi[i]);
          }
          mbar_wait(BAR(B_EMPTY + s), ph ^ 1, p.err, 11);
          if (tid == 0) kstamp(p, 3, it);
          const uint32_t a_hi = smem_u32(sA + s * A_STAGE + sub * A_SUB) + row_off;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const uint32_t off = (((uint32_t)(2 * hf + cc)) ^ sw) << 4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(hi[4 * cc]), "r"(hi[4 * cc + 1]), "r"(hi[4 * cc + 2]), "r"(hi[4 * cc + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + A_OP + off), "r"(lo[4 * cc]), "r"(lo[4 * cc + 1]), "r"(lo[4 * cc + 2]), "r"(lo[4 * cc + 3]) : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { ARRIVE_LEADER(B_FULL + s); mbar_arrive(BAR(B_XEMPTY + xs)); }
          if (tid == 0) kstamp(p, 4, it);
        }
      }
    } else {
    // Single-product modes: two groups of two warps alternate sub-steps (one row of the 64-row slab per thread): each group has two
    // sub-step times for its wait -> load -> convert -> store -> fence -> arrive latency chain.
    const int grp = (warp - CONV_WARP0) >> 1;
    const int row = ((warp - CONV_WARP0) & 1) * 32 + lane;
    uint32_t it = 0;
    for (int wi = 0; wi < n_it; ++wi) {
      const WorkItem w = work_item(p, wi, pair_id, n_pairs, KST);
      for (int ks = w.kb * KSUB; ks < w.ke * KSUB; ++ks, ++it) {
        if ((int)(it & 1) != grp) continue;
        // Parity waits are only sound if the waiter sees every phase of its barrier: the rings are even (XS) or shared by both groups
        // within every stage (KSUB == 2), so a staging slot / stage always belongs to the same group.
        static_assert(LO || (!(XS & 1) && (KSUB == 2 || !(NST & 1))), "two converter groups need even rings");
        const uint32_t xs = it % XS, xph = (it / XS) & 1;
        const uint32_t st = it / KSUB, sub = it % KSUB, s = st % NST, ph = (st / NST) & 1;
        mbar_wait(BAR(B_XFULL + xs), xph, p.err, 10);
        float x[32];
        const uint32_t src = smem_u32(sX + xs * X_SLOT) + (uint32_t)row * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a = src + (((uint32_t)j ^ ((uint32_t)row & 7u)) << 4);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[4 * j]), "=f"(x[4 * j + 1]), "=f"(x[4 * j + 2]), "=f"(x[4 * j + 3]) : "r"(a));
        }
        uint32_t hi[16], lo[16];
        pack_operand_row<FP16, LO>(x, hi, lo);                                      // consumes every staged value
        mbar_wait(BAR(B_EMPTY + s), ph ^ 1, p.err, 11);
        const uint32_t a_hi = smem_u32(sA + s * A_STAGE + sub * A_SUB);
        store_operand_row<LO>(a_hi, a_hi + A_OP, row, hi, lo);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { ARRIVE_LEADER(B_FULL + s); mbar_arrive(BAR(B_XEMPTY + xs)); }
      }
    }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    // ===================== epilogue warps =====================
    // Warp (q, half): TMEM lanes 32 q .. +32 = rows (q & 1) * 32 + lane of this CTA; feature group g = q >> 1 (lanes 64..127
    // hold the upper 128 columns of each 256-wide N block); half = N block.  Its strip: 128 features f0 .. f0 + 128, four
    // 32-column chunks at TMEM columns half * 128 + 32 j of the tile's buffer.  strip id = half * 2 + g = its GEMM2 ring slot.
    const int q = warp & 3, half = (warp - EPI_WARP0) >> 2, g = q >> 1, rg = q & 1;
    const int row = rg * 32 + lane;
    const int strip = half * 2 + g;
    const int f0 = half * 256 + g * 128;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    const int et = threadIdx.x - EPI_WARP0 * 32;             // 0..255
    uint32_t tl = 0;
    const uint32_t gskip = (n_it && work_item(p, 0, pair_id, n_pairs, KST).kind == 2) ? 1u : 0u;   // GEMM2 tiles = items - gskip (see WorkItem)

    for (int i = et; i < HMAX; i += 256) c_b1[i] = p.b1 ? p.b1[i] : 0.f;
    for (int i = et; i < 128; i += 256) { c_ba[i] = p.ba ? p.ba[i] : 0.f; c_wc[i] = p.wc[i]; }
    named_bar_sync(1, 256);

    float m_run = -INFINITY, l_run = 0.f;
    float prun[4] = {0.f, 0.f, 0.f, 0.f};                  // registers: prun[j] (this lane) = pooled partial of feature f0 + 32 j + lane
    const float bc = p.bc ? p.bc[0] : 0.f;
    const uint32_t a2_hi = smem_u32(sA2 + strip * G2A_STAGE);

    for (int wi = 0; wi < n_it; ++wi, ++tl) {
      const WorkItem wk = work_item(p, wi, pair_id, n_pairs, KST);
      const int64_t tile = wk.tile;
      const uint32_t b = tl & 1;
      const int64_t grow = tile * BMP + rank * BMC + row;
      const uint32_t tb = tq + b * 256 + (uint32_t)(half * 128);        // this warp's strip in the tile's accumulator buffer
      if (wi == n_it - 2) {
        // the owner's split item comes next (last): pull this warp's 16 KB of the helper's partial (dumped long ago, possibly evicted by the bag
        // stream) back into L2 while this tile's epilogue runs.  A hint only: the flag is still checked before the loads.
        const WorkItem nx = work_item(p, wi + 1, pair_id, n_pairs, KST);
        if (nx.kind == 1) {
          const char* base = reinterpret_cast<const char*>(p.split_buf) + ((size_t)(nx.pidx * 2 + rank) * 8 + (warp - EPI_WARP0)) * (4 * 8 * 32 * 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(lane + 32 * k) * 128));
        }
      }
      mbar_wait(BAR(B_ACCFULL + b), (tl >> 1) & 1, p.err, 13);
      tc_fence_after();
      if (et == 0) trace_stamp(p, tl, 4);
      if (wk.kind == 2) {
        // helper of a split tile: dump the raw partial accumulator (its K range) for the owner pair and raise this CTA's flag
        // layout [partial][CTA rank][warp][chunk j][4-column group][lane] float4: every store / load instruction of a warp covers 512 contiguous bytes
        float4* dst = reinterpret_cast<float4*>(p.split_buf) + ((size_t)(wk.pidx * 2 + rank) * 8 + (warp - EPI_WARP0)) * (4 * 8 * 32) + lane;
#pragma unroll 1
        for (int j = 0; j < 4; j += 2) {                      // two chunks in flight: the second TMEM load overlaps the first chunk's stores
          uint32_t va[32], vb[32];
          tmem_ld32(tb + (uint32_t)(j * 32), va);
          tmem_ld32(tb + (uint32_t)((j + 1) * 32), vb);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            __stcg(dst + (j * 8 + i) * 32, make_float4(__uint_as_float(va[4 * i]), __uint_as_float(va[4 * i + 1]), __uint_as_float(va[4 * i + 2]), __uint_as_float(va[4 * i + 3])));
#pragma unroll
          for (int i = 0; i < 8; ++i)
            __stcg(dst + ((j + 1) * 8 + i) * 32, make_float4(__uint_as_float(vb[4 * i]), __uint_as_float(vb[4 * i + 1]), __uint_as_float(vb[4 * i + 2]), __uint_as_float(vb[4 * i + 3])));
        }
        if (et == 0) trace_stamp(p, tl, 5);
        tc_fence_before();
        named_bar_sync(1, 256);                               // every epilogue thread's stores are ordered before thread 0's release below (cumulativity)
        if (et == 0) trace_stamp(p, tl, 7);
        if (et == 0) flag_raise(p.split_flags + wk.pidx * 2 + rank);   // st.release.gpu = fence + store, the grid-sync pattern
        if (et == 0) trace_stamp(p, tl, 8);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(LEADER(B_ACCEMPTY + b));
        if (et == 0) trace_stamp(p, tl, 9);
        continue;
      }
      // owner: accumulator chunk j of this warp's strip += the helper's partial sum.  The partial of chunk j + 1 is fetched (L2 loads: another
      // SM wrote it) while chunk j is processed, so one L2 latency is exposed per tile instead of one per chunk.
      float4 pf[8];
      const float4* pf_src = reinterpret_cast<const float4*>(p.split_buf) + ((size_t)(wk.pidx * 2 + rank) * 8 + (warp - EPI_WARP0)) * (4 * 8 * 32) + lane;
      auto pf_load = [&](int j) {
        if (wk.kind != 1) return;
#pragma unroll
        for (int i = 0; i < 8; ++i) pf[i] = __ldcg(pf_src + (j * 8 + i) * 32);
      };
      auto pf_add = [&](float (&hv)[32]) {
        if (wk.kind != 1) return;
#pragma unroll
        for (int i = 0; i < 8; ++i) { hv[4 * i] += pf[i].x; hv[4 * i + 1] += pf[i].y; hv[4 * i + 2] += pf[i].z; hv[4 * i + 3] += pf[i].w; }
      };
      if (wk.kind == 1) {                                     // the helper's partial of this CTA's 64 rows must have landed
        if (lane == 0) flag_wait(p.split_flags + wk.pidx * 2 + rank, p.err, 30);
        __syncwarp();
        pf_load(0);
      }
      if (p.dbg & 512) {                                    // timing attribution: handshakes only, no epilogue work
        if (half == 0) { __syncwarp(); if (lane == 0) mbar_arrive_cluster(LEADER(B_TAILFREE + b)); }
        for (int j = 0; j < 4; ++j) {
          mbar_wait(BAR(B_G2AEMPTY + strip), ((tl * 4 + (uint32_t)j) & 1u) ^ 1, p.err, 14);
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(LEADER(B_G2AFULL + strip));
        }
        mbar_wait(BAR(B_UFULL), (tl - gskip) & 1, p.err, 15);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(LEADER(B_ACCEMPTY + b));
        continue;
      }

      // E1 (N-block-0 warps): vacate the buffer's first 64 columns (they become GEMM2's accumulator); keep h in registers
      float keep_h[2][32];
      if (half == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          tmem_ld32f(tb + (uint32_t)(j * 32), keep_h[j]);
          pf_add(keep_h[j]);
          pf_load(j + 1);
          bias_act32<ACT>(keep_h[j], c_b1 + f0 + j * 32, p.act, p.w1_inv);
          if (p.drop_mode) drop_apply32(keep_h[j], drop_keep_word(p, grow, (f0 >> 5) + j, HMAX / 32), p.drop_scale);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(LEADER(B_TAILFREE + b));
      }
      if (et == 0) trace_stamp(p, tl, 5);

      // E2: h chunk -> 16-bit A2 tile of this strip's ring slot (+ h back into TMEM for the pooling pass, + optional outputs)
      float tacc[4] = {0.f, 0.f, 0.f, 0.f};
      auto emit_chunk = [&](int j, const float (&hv)[32]) {
        const int fc = f0 + j * 32;
        if (p.h_out && grow < p.N) {
          float* dst = p.h_out + grow * HMAX + fc;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(hv[i], hv[i + 1], hv[i + 2], hv[i + 3]);
        }
        if (p.t_out) {
#pragma unroll 1
          for (int cc = 0; cc < p.C; ++cc) {
            const float4* wp = reinterpret_cast<const float4*>(p.Wp + cc * HMAX + fc);
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 w4 = __ldg(wp + i);
              a = fmaf(hv[4 * i], w4.x, a); a = fmaf(hv[4 * i + 1], w4.y, a); a = fmaf(hv[4 * i + 2], w4.z, a); a = fmaf(hv[4 * i + 3], w4.w, a);
            }
            tacc[cc] += a;
          }
        }
        const uint32_t ph = (tl * 4 + (uint32_t)j) & 1u;
        mbar_wait(BAR(B_G2AEMPTY + strip), ph ^ 1, p.err, 14);
        write_operand_row<FP16, LO>(a2_hi, a2_hi + A2_OP, row, hv);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(LEADER(B_G2AFULL + strip));
      };
      if (half == 0) {
        emit_chunk(0, keep_h[0]);
        emit_chunk(1, keep_h[1]);
      }
#pragma unroll 1
      for (int j = (half == 0 ? 2 : 0); j < 4; ++j) {
        float hv[32];
        tmem_ld32f(tb + (uint32_t)(j * 32), hv);
        pf_add(hv);
        if (j < 3) pf_load(j + 1);
        bias_act32<ACT>(hv, c_b1 + f0 + j * 32, p.act, p.w1_inv);
        if (p.drop_mode) drop_apply32(hv, drop_keep_word(p, grow, (f0 >> 5) + j, HMAX / 32), p.drop_scale);
        tmem_st32f(tb + (uint32_t)(j * 32), hv);
        emit_chunk(j, hv);
      }
      tmem_wait_st();
      if (et == 0) trace_stamp(p, tl, 6);

      // E3: attention logit of every row: s = wc . f(u + ba) + bc.  u (N = 128) sits in the buffer's first 64 columns:
      // lanes 0..63 hold Da 0..63, lanes 64..127 hold Da 64..127; this warp takes 32 of its 64 columns.
      mbar_wait(BAR(B_UFULL), (tl - gskip) & 1, p.err, 15);
      tc_fence_after();
      if (et == 0) trace_stamp(p, tl, 7);
      {
        const int da0 = g * 64 + half * 32;
        float uv[32];
        tmem_ld32f(tq + b * 256 + (uint32_t)(half * 32), uv);
        bias_act32<ATT>(uv, c_ba + da0, p.att_act, p.wa_inv);
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(c_wc + da0 + i);
          s4[0] = fmaf(uv[i], w4.x, s4[0]); s4[1] = fmaf(uv[i + 1], w4.y, s4[1]);
          s4[2] = fmaf(uv[i + 2], w4.z, s4[2]); s4[3] = fmaf(uv[i + 3], w4.w, s4[3]);
        }
        s_part[strip * 64 + row] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      if (p.t_out)
        for (int cc = 0; cc < p.C; ++cc) t_part[(strip * 64 + row) * 4 + cc] = tacc[cc];
      named_bar_sync(1, 256);
      float sv = ((s_part[row] + s_part[64 + row]) + (s_part[128 + row] + s_part[192 + row])) + bc;
      const bool valid = grow < p.N && (!p.keep || p.keep[grow]);
      if (!valid) sv = -INFINITY;
      if (strip == 0 && grow < p.N) {
        if (p.s_out) p.s_out[grow] = sv;
        if (p.t_out)
          for (int cc = 0; cc < p.C; ++cc)
            p.t_out[grow * p.C + cc] = (t_part[row * 4 + cc] + t_part[(64 + row) * 4 + cc]) + (t_part[(128 + row) * 4 + cc] + t_part[(192 + row) * 4 + cc]);
      }
      named_bar_sync(1, 256);                               // s_part / t_part may be overwritten by the next tile after this
      if (wk.kind == 1 && et == 0) p.split_flags[wk.pidx * 2 + rank] = 0;   // every warp is past E2: the partial was consumed; re-arm for the next launch
      if (et == 0) trace_stamp(p, tl, 8);

      // online softmax over the 32 rows of this warp (the four warps that share these rows compute identical m, l)
      const float m_new = fmaxf(m_run, warp_max(sv));
      float w = 0.f, scale = 1.f;
      if (m_new > -INFINITY) {
        scale = (m_run > -INFINITY) ? expf(m_run - m_new) : 0.f;
        w = valid ? expf(sv - m_new) : 0.f;
      }
      l_run = l_run * scale + warp_sum(w);
      m_run = m_new;

      // E4: p += sum_rows w * h over this warp's 128 features, two chunks per transpose step
      if (half == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { keep_h[0][i] *= w; keep_h[1][i] *= w; }
        float r0, r1;
        warp_transpose_sum2(keep_h[0], keep_h[1], r0, r1);
        prun[0] = prun[0] * scale + r0;
        prun[1] = prun[1] * scale + r1;
      }
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        if (half == 0 && j == 0) continue;
        float ha[32], hb[32];
        uint32_t va[32], vb[32];
        tmem_ld32(tb + (uint32_t)(j * 32), va);
        tmem_ld32(tb + (uint32_t)((j + 1) * 32), vb);
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
      if (lane == 0) mbar_arrive_cluster(LEADER(B_ACCEMPTY + b));
      if (et == 0) trace_stamp(p, tl, 9);
    }

    // CTA partial: merge the 2 row groups -> part[blockIdx] = (m, l, P[512])
    float* red_m = s_part;                                  // [2]
    float* red_l = s_part + 2;                              // [2]
    named_bar_sync(1, 256);
#pragma unroll
    for (int j = 0; j < 4; ++j) p_acc[(warp - EPI_WARP0) * 128 + j * 32 + lane] = prun[j];   // every staged load was consumed: the ring is idle
    if (strip == 0 && lane == 0) { red_m[rg] = m_run; red_l[rg] = l_run; }
    named_bar_sync(1, 256);
    const float m_cta = fmaxf(red_m[0], red_m[1]);
    const float f_0 = (red_m[0] > -INFINITY) ? expf(red_m[0] - m_cta) : 0.f, f_1 = (red_m[1] > -INFINITY) ? expf(red_m[1] - m_cta) : 0.f;
    float* out = p.part + (int64_t)blockIdx.x * (2 + HMAX);
    for (int c = et; c < HMAX; c += 256) {
      const int hf = c >> 8, gg = (c >> 7) & 1, j = (c >> 5) & 3, ln = c & 31;
      const float* base = p_acc + (hf * 4 + gg * 2) * 128 + j * 32 + ln;     // warp (hf, q = 2 gg + rg)
      out[2 + c] = fmaf(base[0], f_0, base[128] * f_1);
    }
    if (et == 0) {
      out[0] = m_cta;
      out[1] = red_l[0] * f_0 + red_l[1] * f_1;
    }
    grid_finalize(p, et, lane, t_part, smem, (uint32_t)(sMisc - smem), BAR(B_FIN), reinterpret_cast<int*>(s_part + 16));
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // nobody exits while the peer may still signal its barriers
  if (p.trace && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[256 + 2 * blockIdx.x + 1] = (long long)gt;
    if (blockIdx.x == (unsigned)p.trace_cta) p.trace[239] = clock64();
  }
  if (warp == W_W1) { tc_fence_after(); tmem_dealloc_pair(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------------------
// fp32 weights -> pre-swizzled 16-bit operand images in the pair layout (once per weight version).
// Every tile is the exact shared-memory image of a UMMA K-major SWIZZLE_64B operand tile (rows of 64 B = 32 elements):
//   byte(r, c, e) = (r/8)*512 + (r%8)*64 + ((c ^ ((r>>1)&3))*16) + 2e,  c = 16-byte chunk inside the row.
// W1 image: for k-step ks, CTA rank, operand op (hi, lo), N block blk: the 128 rows  blk*256 + rank*128 + [0,128)  (8 KB);
//           order (stage = ks / ksub, rank, sub = ks % ksub, op, blk) -> one pipeline stage of one CTA is ksub x NOP x 16 KB contiguous.
// Wa image: GEMM2 chunk c = 4 j + half*2 + g covers features half*256 + g*128 + 32 j .. + 32 (the order in which the epilogue
//           strips produce A2 tiles); for chunk c, CTA rank, operand op: the 64 rows rank*64 + [0,64) (4 KB); order (c, rank, op).
// ------------------------------------------------------------------------------------------------------------
template <bool FP16, bool LO>
__global__ void pair_split_w1_kernel(const float* __restrict__ w, int H, int K, uint8_t* __restrict__ img, int ksub, float scale) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (row f, 16-byte chunk kc = k / 8)
  const int kchunks = K / 8;
  if (item >= (int64_t)H * kchunks) return;
  const int f = (int)(item / kchunks), kc = (int)(item % kchunks);
  const int ks = kc >> 2, c = kc & 3;
  const int blk = f >> 8, rank = (f >> 7) & 1, r = f & 127;
  float4 a = *reinterpret_cast<const float4*>(w + (int64_t)f * K + kc * 8);
  float4 b = *reinterpret_cast<const float4*>(w + (int64_t)f * K + kc * 8 + 4);
  a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
  uint32_t h[4], l[4];
  h[0] = pack_hi<FP16>(a.x, a.y); h[1] = pack_hi<FP16>(a.z, a.w); h[2] = pack_hi<FP16>(b.x, b.y); h[3] = pack_hi<FP16>(b.z, b.w);
  constexpr int NOPK = LO ? 2 : 1;
  const size_t off = (size_t)(r >> 3) * 512 + (size_t)(r & 7) * 64 + (size_t)((c ^ ((r >> 1) & 3)) << 4);
  uint8_t* dst = img + ((size_t)(((ks / ksub) * 2 + rank) * ksub + ks % ksub) * NOPK) * 16384 + (size_t)blk * 8192 + off;
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (LO) {
    l[0] = pack_lo<FP16>(a.x, a.y, h[0]); l[1] = pack_lo<FP16>(a.z, a.w, h[1]); l[2] = pack_lo<FP16>(b.x, b.y, h[2]); l[3] = pack_lo<FP16>(b.z, b.w, h[3]);
    *reinterpret_cast<uint4*>(dst + 16384) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

template <bool FP16, bool LO>
__global__ void pair_split_wa_kernel(const float* __restrict__ w, int Da, int K, uint8_t* __restrict__ img, float scale) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (row d, 16-byte chunk kc)
  const int kchunks = K / 8;
  if (item >= (int64_t)Da * kchunks) return;
  const int d = (int)(item / kchunks), kc = (int)(item % kchunks);
  const int fi = kc >> 2, c = kc & 3;                                       // fi = feature / 32
  const int half = fi >> 3, g = (fi >> 2) & 1, j = fi & 3;
  const int chunk = 4 * j + half * 2 + g;
  const int rank = d >> 6, r = d & 63;
  float4 a = *reinterpret_cast<const float4*>(w + (int64_t)d * K + kc * 8);
  float4 b = *reinterpret_cast<const float4*>(w + (int64_t)d * K + kc * 8 + 4);
  a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
  uint32_t h[4], l[4];
  h[0] = pack_hi<FP16>(a.x, a.y); h[1] = pack_hi<FP16>(a.z, a.w); h[2] = pack_hi<FP16>(b.x, b.y); h[3] = pack_hi<FP16>(b.z, b.w);
  constexpr int NOPK = LO ? 2 : 1;
  const size_t off = (size_t)(r >> 3) * 512 + (size_t)(r & 7) * 64 + (size_t)((c ^ ((r >> 1) & 3)) << 4);
  constexpr int GSK = 4 / NOPK;                                             // chunks per 16 KB load group
  uint8_t* dst = img + ((size_t)((chunk / GSK) * 2 + rank) * GSK + (size_t)(chunk % GSK)) * NOPK * 4096 + off;
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (LO) {
    l[0] = pack_lo<FP16>(a.x, a.y, h[0]); l[1] = pack_lo<FP16>(a.z, a.w, h[1]); l[2] = pack_lo<FP16>(b.x, b.y, h[2]); l[3] = pack_lo<FP16>(b.z, b.w, h[3]);
    *reinterpret_cast<uint4*>(dst + 4096) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

template <int NPROD, bool FP16, int NST, int XS, int KSUB, int ACT, int ATT>
static int launch_pair(const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mwa, const FusedParams& p, int grid, cudaStream_t stream) {
  constexpr int NOP = NPROD == 3 ? 2 : 1;
  const size_t smem = 1024 + (size_t)XS * X_SLOT + (size_t)NST * KSUB * NOP * (A_OP + B_OP) + (size_t)G2S * NOP * A2_OP + 2 * (size_t)G2B_BUF + MISC_BYTES;
  auto kern = mil_fused2_kernel<NPROD, FP16, NST, XS, KSUB, ACT, ATT>;
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
  kern<<<grid, NUM_THREADS, smem, stream>>>(mx, mw1, mwa, p);
  prof_end(stream);
  MIL_LAUNCH_CHECK();
  return 0;
}

template <int ACT, int ATT>
static int dispatch_prec(int precision, const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mwa, const FusedParams& p, int grid,
                         cudaStream_t stream) {
  if (precision == MIL_PREC_BF16X3) return launch_pair<3, false, 3, 4, 1, ACT, ATT>(mx, mw1, mwa, p, grid, stream);
  if (precision == MIL_PREC_FP16X3) return launch_pair<3, true, 3, 4, 1, ACT, ATT>(mx, mw1, mwa, p, grid, stream);
  if (p.D % 64 == 0) {
    if (precision == MIL_PREC_FP16) return launch_pair<1, true, 3, 4, 2, ACT, ATT>(mx, mw1, mwa, p, grid, stream);
    return launch_pair<1, false, 3, 4, 2, ACT, ATT>(mx, mw1, mwa, p, grid, stream);
  }
  if (precision == MIL_PREC_FP16) return launch_pair<1, true, 6, 4, 1, ACT, ATT>(mx, mw1, mwa, p, grid, stream);
  return launch_pair<1, false, 6, 4, 1, ACT, ATT>(mx, mw1, mwa, p, grid, stream);
}

}  // namespace pairk

// 64-wide stages (two 32-wide sub-steps) for the single-product modes when D allows; the 3-product mode has no room for them
static int pair_ksub(int precision, int D) { return (!prec_split(precision) && D % 64 == 0) ? 2 : 1; }

size_t pair_weight_image_bytes(int D, int H, int Da) { return ((size_t)H * D + (size_t)Da * H) * 2 * sizeof(uint16_t); }

int pair_build_images(const float* W1, int H, int D, const float* Wa, int Da, uint8_t* w1_img, uint8_t* wa_img, int precision, cudaStream_t stream) {
  using namespace pairk;
  const int64_t i1 = (int64_t)H * (D / 8), i2 = (int64_t)Da * (H / 8);
  const unsigned b1 = (unsigned)((i1 + 255) / 256), b2 = (unsigned)((i2 + 255) / 256);
  const float sc = prec_wscale(precision);
  if (precision == MIL_PREC_BF16X3) {
    pair_split_w1_kernel<false, true><<<b1, 256, 0, stream>>>(W1, H, D, w1_img, pair_ksub(precision, D), sc);
    pair_split_wa_kernel<false, true><<<b2, 256, 0, stream>>>(Wa, Da, H, wa_img, sc);
  } else if (precision == MIL_PREC_FP16X3) {
    pair_split_w1_kernel<true, true><<<b1, 256, 0, stream>>>(W1, H, D, w1_img, pair_ksub(precision, D), sc);
    pair_split_wa_kernel<true, true><<<b2, 256, 0, stream>>>(Wa, Da, H, wa_img, sc);
  } else if (precision == MIL_PREC_FP16) {
    pair_split_w1_kernel<true, false><<<b1, 256, 0, stream>>>(W1, H, D, w1_img, pair_ksub(precision, D), sc);
    pair_split_wa_kernel<true, false><<<b2, 256, 0, stream>>>(Wa, Da, H, wa_img, sc);
  } else {
    pair_split_w1_kernel<false, false><<<b1, 256, 0, stream>>>(W1, H, D, w1_img, pair_ksub(precision, D), sc);
    pair_split_wa_kernel<false, false><<<b2, 256, 0, stream>>>(Wa, Da, H, wa_img, sc);
  }
  MIL_LAUNCH_CHECK();
  return 0;
}

int pair_plan(FusedParams& p, int precision);

// p: every field of the fused pass filled in by the caller (mil_abmil_fused_fwd_f32), w1_img / wa_img in the pair layout.
int pair_fused_launch(const float* X, FusedParams p, int precision, cudaStream_t stream) {
  using namespace pairk;
  const int NOP = prec_split(precision) ? 2 : 1;
  CUtensorMap mx, mw1, mwa;
  int rc;
  if ((rc = make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, (uint64_t)p.N, (uint64_t)p.D, BMC, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const uint64_t w1_rows = (uint64_t)HMAX * p.D * 2 * NOP / 256, wa_rows = (uint64_t)128 * HMAX * 2 * NOP / 256;
  if ((rc = make_map_2d(&mw1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.w1_img, w1_rows, 256, (uint32_t)(pair_ksub(precision, p.D) * NOP * B_OP / 256), 256, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  if ((rc = make_map_2d(&mwa, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.wa_img, wa_rows, 256, (uint32_t)(G2B_BUF / 256), 256, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  const int grid = 2 * pair_plan(p, precision);
#define MIL_CASE(A, T) if (p.act == A && p.att_act == T) return dispatch_prec<A, T>(precision, mx, mw1, mwa, p, grid, stream);
  MIL_CASE(MIL_ACT_RELU, MIL_ACT_TANH) MIL_CASE(MIL_ACT_GELU, MIL_ACT_TANH)
  MIL_CASE(MIL_ACT_RELU, MIL_ACT_RELU) MIL_CASE(MIL_ACT_GELU, MIL_ACT_RELU)
  MIL_CASE(MIL_ACT_RELU, MIL_ACT_GELU) MIL_CASE(MIL_ACT_GELU, MIL_ACT_GELU)
#undef MIL_CASE
  set_error("fused pass: unsupported activation pair act=%d att_act=%d (use the composed path)", p.act, p.att_act);
  return -1;
}

// Grid and tail-split plan of the pair pipeline for p.N rows: fills p.split_* and returns the number of CTA pairs.
int pair_plan(FusedParams& p, int precision) {
  using namespace pairk;
  const int64_t n_tiles = (p.N + BMP - 1) / BMP;
  int pmax = num_sms() / 2;
  const char* e = getenv("MHIMK_GRID");
  if (e && atoi(e) >= 2 && atoi(e) / 2 < pmax) pmax = atoi(e) / 2;
  int pairs = n_tiles < pmax ? (int)n_tiles : pmax;
  // Tail split: the tiles of a last wave that fills at most half of the pairs are shared by K range between two pairs each (the owner adds one
  // partial accumulator of 128 KB per CTA; >= 4 pipeline stages per part).  More parts do not pay: the split item's epilogue cannot start
  // before the previous tile's (~25 k cycles) has finished, and half a GEMM1 (~16 k) already fits under it; every extra partial costs the owner
  // an L2 round trip per chunk.  391 tiles on 74 pairs: 5 waves + 21 tiles x 2 halves instead of a 6th tile period for 21 pairs.
  p.split_s = 1; p.split_full = 0; p.split_rem = 0;
  static const bool nosplit = getenv("MHIMK_NOSPLIT") && atoi(getenv("MHIMK_NOSPLIT"));
  if (!nosplit && p.split_buf && p.split_flags) {
    const int kst = p.D / BK / pair_ksub(precision, p.D);
    const int rem = (int)(n_tiles % pmax);                  // == n_tiles when the bag has fewer tiles than pairs
    int S = rem ? pmax / rem : 1;
    if (S > 2) S = 2;
    while (S > 1 && kst / S < 4) --S;
    if (S >= 2) {
      p.split_s = S; p.split_full = (int)(n_tiles / pmax); p.split_rem = rem;
      pairs = p.split_full ? pmax : rem * S;
    }
  }
  return pairs;
}

// Test hook (mil_pair_plan_item): the i-th work item of CTA pair `pair` under the plan for (N, D, precision) on this device's SM count.
// out[0..4] = tile, first stage, end stage, kind (0 whole tile, 1 owner, 2 helper), partial index; out[5..8] = pairs, split_s, split_full, split_rem.
int pair_plan_item(int64_t N, int D, int precision, int pair, int i, int64_t* out) {
  using namespace pairk;
  FusedParams p;
  p.N = N; p.D = D;
  p.split_buf = reinterpret_cast<float*>(1); p.split_flags = reinterpret_cast<int*>(1);      // "the workspace has the exchange area"
  const int pairs = pair_plan(p, precision);
  const int64_t n_tiles = (N + BMP - 1) / BMP;
  const int n_it = pair < pairs ? work_count(p, pair, pairs, n_tiles) : 0;
  out[5] = pairs; out[6] = p.split_s; out[7] = p.split_full; out[8] = p.split_rem;
  if (i < 0 || i >= n_it) return n_it;
  const WorkItem w = work_item(p, i, pair, pairs, D / BK / pair_ksub(precision, D));
  out[0] = w.tile; out[1] = w.kb; out[2] = w.ke; out[3] = w.kind; out[4] = w.pidx;
  return n_it;
}

}  // namespace mil
