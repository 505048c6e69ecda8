// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and epilogue helpers shared by the fused-pass kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "mil_common.cuh"

namespace mil {

constexpr uint64_t WAIT_TIMEOUT_CYCLES = 4000000000ull;   // ~2 s: trap instead of hanging the GPU on a pipeline bug
// fp16 hi+lo weight images are built from W * 16: typical weights (|w| ~ 0.03 .. 1) then keep their rounding residual in fp16's normal
// range (>= 6.1e-5) instead of its subnormals; the epilogue multiplies the accumulator by 1/16 (exact).
constexpr float W_SCALE_FP16X3 = 16.f;
constexpr int HMAX = 512;               // embedding width = accumulator columns of one 128-row tile

struct FusedParams {
  int64_t N;            // rows
  int D;                // K of GEMM1
  int nout;             // N of GEMM1 (512 in fused mode)
  int Da;               // N of GEMM2
  int act, att_act;
  const float* b1; const float* ba; const float* wc; const float* bc;
  const uint8_t* keep; const float* Wp; int C;
  float* s_out; float* t_out; float* h_out; float* part;
  float* c_out; int64_t ldc;   // MODE_STORE
  const uint8_t* w1_img; const uint8_t* wa_img;   // pre-swizzled 16-bit weight images (see split_weights_kernel)
  float* stats; float* pooled; unsigned int* counter;      // in-kernel finalisation by the last CTA to finish
  float* rec_out;       // nullable [2 + H]: (m, l, P = sum e^{s-m} h) = the exchange record of an instance-sharded bag
  const float* Wcls; const float* bcls; int n_cls; float* logits;
  int* err;
  long long* trace;     // optional [16 tiles][16 slots] clock64 stamps of CTA trace_cta (MHIMK_TRACE=1, MHIMK_TRACE_CTA=<block>), see tools/trace_fused.py
  int trace_cta = 0;
  // dropout on h (mhim.py:193-194, abmil.py:188-189): 0 = none, 1 = caller-supplied keep bits, 2 = in-kernel Philox4x32-10
  int drop_mode; const uint32_t* drop_bits; uint32_t drop_thresh; float drop_scale; uint32_t drop_seed[2]; uint32_t drop_off[2];
  float w1_inv, wa_inv; // 1 / (power-of-two scale of the W1 / Wa image): MIL_W_SCALE_FP16X3 in the fp16x3 arithmetic, else 1
  // tail split of the pair pipeline (mil_fused2_sm100.cu): the last, partly filled wave of tiles is split over idle pairs by K range.
  // split_s = parts per tile (1 = off), split_full = whole-tile waves, split_rem = tiles of the split wave; helpers dump their
  // accumulators to split_buf ([partial][CTA rank][epilogue warp][...] in the warps' own register order, 128 KB per CTA) and raise
  // split_flags [partial][2 CTA ranks]; the owner adds them before the bias.
  int split_s = 1, split_full = 0, split_rem = 0; float* split_buf = nullptr; int* split_flags = nullptr;
  int nblk = 1;         // store mode of the single pipeline: 256-column blocks per row tile handled by separate CTAs (nout = columns per CTA)
  int dbg;              // MHIMK_DEBUG bitmask (timing attribution only): 1 skip W1 TMA, 2 skip X TMA, 4 skip GEMM1 MMA, 8 skip convert, 16 skip Wa TMA, 32 skip pooling, 64 skip GEMM2 MMA
};

// ------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// MHIMK_NOTRAP=1 (debugging): a timed-out wait records its code (first one wins, plus block id) and gives up after ~0.05 s
// instead of trapping, so that the launch completes and the host can read the code from the workspace.
static __device__ int g_mil_notrap = 0;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    const uint64_t dt = (uint64_t)(clock64() - t0);
    if (g_mil_notrap && dt > 100000000ull) {
      if (err) atomicCAS(err, 0, code | ((int)blockIdx.x << 8) | ((int)(threadIdx.x >> 5) << 20));
      return;
    }
    if (dt > WAIT_TIMEOUT_CYCLES) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}
// acquire-spin on a global flag raised by another CTA of the same launch (st.release.gpu); same time-out policy as mbar_wait
__device__ __forceinline__ void flag_wait(const int* f, int* err, int code) {
  const long long t0 = clock64();
  for (;;) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v) return;
    const uint64_t dt = (uint64_t)(clock64() - t0);
    if (g_mil_notrap && dt > 100000000ull) {
      if (err) atomicCAS(err, 0, code | ((int)blockIdx.x << 8) | ((int)(threadIdx.x >> 5) << 20));
      return;
    }
    if (dt > WAIT_TIMEOUT_CYCLES) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void flag_raise(int* f) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(1) : "memory"); }
static void mil_set_notrap() {
  static bool done_dev[64] = {false};           // the symbol lives per device
  int dev = 0;
  cudaGetDevice(&dev);
  bool& done = done_dev[dev & 63];
  if (done) return;
  done = true;
  const char* e = getenv("MHIMK_NOTRAP");
  const int v = e ? atoi(e) : 0;
  if (v) cudaMemcpyToSymbol(g_mil_notrap, &v, sizeof(int));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): one request moves a whole pre-swizzled operand tile
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// pull one box of a tensor map into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// one lane of a fully converged warp (always the same one)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- thread-block cluster / CTA-pair (cta_group::2) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (.release.cta) semantics, as cutlass::arch::ClusterBarrier::arrive(cta_id): the .release.cluster form costs a
  // MEMBAR.ALL.GPU per arrival (measured: ~1000 cycles per k-step in the converter chain)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// TMA tile load issued by either CTA of a pair; `bar_cluster` may live in the peer CTA (cta_group::2 form)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T over the CTA pair: M = 128 (64 rows per CTA) or 256; each CTA supplies its half of A's rows and of B's rows
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once all earlier MMAs of the pair are complete) on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

#define TMEM_REGS32(v) \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), \
  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),  \
  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define TMEM_REGS32_IN(v) \
  "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),   \
  "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),    \
  "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])

// 32 lanes x 32 consecutive columns: lane i of the warp gets row (lane_base + i), v[j] = column (col + j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : TMEM_REGS32(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), TMEM_REGS32_IN(v) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld();
__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float (&f)[32]) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tmem_st32f(uint32_t taddr, const float (&f)[32]) {
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(f[i]);
  tmem_st32(taddr, v);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void trace_stamp(const FusedParams& p, uint32_t tl, int slot) {
  if (p.trace && blockIdx.x == (unsigned)p.trace_cta && tl < 16) p.trace[tl * 16 + slot] = clock64();
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows are 64 B, 8-row groups are 512 B apart.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major) bits [16,30)
  d |= (uint64_t)(512 >> 4) << 32;                      // stride byte offset = 512  bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell) bits [46,48)
  d |= (uint64_t)4 << 61;                               // layout type SWIZZLE_64B  bits [61,64)
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: fp32 accumulate, A/B both K-major.
__device__ __forceinline__ uint32_t make_idesc(int fp16, int n, int m = 128) {
  const uint32_t fmt = fp16 ? 0u : 1u;                  // 0 = F16, 1 = BF16
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// fp32 pair -> packed 16-bit hi (and the packed 16-bit rounding residual lo)
template <bool FP16>
__device__ __forceinline__ uint32_t pack_hi(float x0, float x1) {
  uint32_t r;
  if (FP16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));   // |x| > 65504 saturates instead of becoming inf
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
  return r;
}
__device__ __forceinline__ uint32_t pack_lo_bf16(float x0, float x1, uint32_t hi) {
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
  return pack_hi<false>(x0 - h0, x1 - h1);
}

// packed 16-bit rounding residual of (x0, x1) against their packed 16-bit hi parts, in the same format as hi
template <bool FP16>
__device__ __forceinline__ uint32_t pack_lo(float x0, float x1, uint32_t hi) {
  if (FP16) {
    float h0, h1;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    return pack_hi<true>(x0 - h0, x1 - h1);
  }
  return pack_lo_bf16(x0, x1, hi);
}

// Write one row (32 consecutive K elements) of a [128 x 32] 16-bit operand tile in the UMMA K-major SWIZZLE_64B layout.
template <bool FP16, bool LO>
__device__ __forceinline__ void write_operand_row(uint32_t a_hi, uint32_t a_lo, int row, const float (&x)[32]) {
  const uint32_t row_off = (uint32_t)(row >> 3) * 512u + (uint32_t)(row & 7) * 64u;
  const uint32_t sw = (uint32_t)(row >> 1) & 3u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = pack_hi<FP16>(x[8 * c + 2 * i], x[8 * c + 2 * i + 1]);
      if (LO) l[i] = pack_lo<FP16>(x[8 * c + 2 * i], x[8 * c + 2 * i + 1], h[i]);
    }
    const uint32_t off = row_off + (((uint32_t)c ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
    if (LO) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
  }
}

// The same in two phases (pack in registers, store later): lets a converter hand its fp32 staging slot back before it waits for
// the operand stage.  The packs consume every loaded value, so all staging loads have completed when they are done.
template <bool FP16, bool LO>
__device__ __forceinline__ void pack_operand_row(const float (&x)[32], uint32_t (&h)[16], uint32_t (&l)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    h[i] = pack_hi<FP16>(x[2 * i], x[2 * i + 1]);
    l[i] = LO ? pack_lo<FP16>(x[2 * i], x[2 * i + 1], h[i]) : 0u;
  }
}
template <bool LO>
__device__ __forceinline__ void store_operand_row(uint32_t a_hi, uint32_t a_lo, int row, const uint32_t (&h)[16], const uint32_t (&l)[16]) {
  const uint32_t row_off = (uint32_t)(row >> 3) * 512u + (uint32_t)(row & 7) * 64u;
  const uint32_t sw = (uint32_t)(row >> 1) & 3u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t off = row_off + (((uint32_t)c ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(h[4 * c]), "r"(h[4 * c + 1]), "r"(h[4 * c + 2]), "r"(h[4 * c + 3]) : "memory");
    if (LO) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(l[4 * c]), "r"(l[4 * c + 1]), "r"(l[4 * c + 2]), "r"(l[4 * c + 3]) : "memory");
  }
}

// Short inline transcendental activations (MUFU ex2/rcp based).  Inlining libm's erff/tanhf/expf at every element made the
// kernel 600 KB of SASS (instruction-cache bound) and calling them out of line cost ~250 cycles per element.
//   tanh(x)    = 1 - 2 / (1 + e^{2x})                                   |abs err| <~ 5e-7
//   sigmoid(x) = 1 / (1 + e^{-x})                                       |abs err| <~ 2e-7
//   erf(z)     = 1 - (a1 t + ... + a5 t^5) e^{-z^2}, t = 1/(1 + p z)    |abs err| <= 1.5e-7 (Abramowitz-Stegun 7.1.26)
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
// v[i] = act(v[i] + bias[i]); bias points into shared memory (broadcast reads).  ACT is a compile-time constant in the fused
// kernel (one variant per instantiation keeps the epilogue small); ACT = -1 selects at run time (store mode only).
template <int ACT>
// inv_scale undoes the power-of-two scale the weight image was built with (fp16x3 arithmetic; 1 otherwise: fma(v, 1, b) == v + b).
__device__ __forceinline__ void bias_act32(float (&v)[32], const float* bias, int act_rt, float inv_scale = 1.f) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + i);
    v[i] = fmaf(v[i], inv_scale, b.x); v[i + 1] = fmaf(v[i + 1], inv_scale, b.y);
    v[i + 2] = fmaf(v[i + 2], inv_scale, b.z); v[i + 3] = fmaf(v[i + 3], inv_scale, b.w);
  }
  const int act = ACT >= 0 ? ACT : act_rt;
  if (act == MIL_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (act == MIL_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
  } else if (act == MIL_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = tanh_fast(v[i]);
  } else if (act == MIL_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = sigmoid_fast(v[i]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Dropout keep-mask of h.  One 32-bit word per (row, 32-column chunk): bit i = keep flag of column 32 * chunk + i.
//   mode 1: the caller's words (uint32 [rows][ncols/32]) -- parity tests feed the reference's own mask;
//   mode 2: Philox4x32-10, counter = (row, 4 * chunk + quarter, offset_lo, offset_hi), key = seed: every call yields eight
//           16-bit uniforms, one per column; keep iff u16 < thresh, thresh = round((1 - p) * 65536).  Stateless in (row, column),
//           so a backward pass (or a host-side check, tests/philox_ref.py) regenerates the identical mask.
// ------------------------------------------------------------------------------------------------------------
// words_per_row = ncols / 32 of the tensor the mask belongs to
__device__ __forceinline__ uint32_t drop_keep_word(const FusedParams& p, int64_t row, int chunk, int words_per_row) {
  if (p.drop_mode == 1) return row < p.N ? __ldg(p.drop_bits + row * words_per_row + chunk) : 0u;
  if (p.drop_mode == 3) {                                  // (seed, offset) live in device memory (CUDA-graph replays)
    const uint32_t seed[2] = {__ldg(p.drop_bits), __ldg(p.drop_bits + 1)}, off[2] = {__ldg(p.drop_bits + 2), __ldg(p.drop_bits + 3)};
    return philox_keep_word((uint32_t)row, (uint32_t)chunk, p.drop_thresh, seed, off);
  }
  return philox_keep_word((uint32_t)row, (uint32_t)chunk, p.drop_thresh, p.drop_seed, p.drop_off);
}
__device__ __forceinline__ void drop_apply32(float (&v)[32], uint32_t word, float scale) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = ((word >> i) & 1u) ? v[i] * scale : 0.f;
}

// column sums over the 32 lanes of a warp of v[0..31] (one value per column per lane): afterwards lane j holds sum_rows v_row[j].
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float give = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, give, s);
    }
  }
  return v[0];
}

// two independent transposes interleaved (the single one is a 5-level dependent shuffle chain: latency bound)
__device__ __forceinline__ void warp_transpose_sum2(float (&a)[32], float (&b)[32], float& ra, float& rb) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float ka = up ? a[i + s] : a[i], ga = up ? a[i] : a[i + s];
      const float kb = up ? b[i + s] : b[i], gb = up ? b[i] : b[i + s];
      a[i] = ka + __shfl_xor_sync(0xffffffffu, ga, s);
      b[i] = kb + __shfl_xor_sync(0xffffffffu, gb, s);
    }
  }
  ra = a[0];
  rb = b[0];
}

// ------------------------------------------------------------------------------------------------------------
// Grid-level finalisation, called by the 256 epilogue threads of every CTA after its partial (m, l, P[512]) is in p.part:
// the last CTA to arrive merges all partials (log-sum-exp, fixed order -> deterministic), normalises the pooled vector and
// applies the classifier.  No extra launch on the critical path.
// The partial records (148 x 2056 B) come in through the TMA engine: two bulk copies in flight into a double-buffered staging
// area in the (by then idle) operand rings, summed from shared memory.  Plain loads by one CTA ran at ~20 B/cycle and made
// this tail 15 us; the bulk copies run at the SM's full ingest rate.
//   et: 0..255 thread index inside the epilogue group; scratch: >= 512 floats of shared memory; stage / stage_bytes: staging
//   area (16-byte aligned, >= 2 x 2056 B); fin_bar: two consecutive initialised (count 1) mbarriers nobody else uses.
//   p.part must be readable up to 8 bytes past the last used record (mil_fused_num_partials() reserves one spare record).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_finalize(const FusedParams& p, int et, int lane, float* scratch, uint8_t* stage, uint32_t stage_bytes,
                                              uint32_t fin_bar, int* flag) {
  constexpr uint32_t REC = (2 + HMAX) * 4;               // 2056 bytes per partial record
  __threadfence();
  named_bar_sync(1, 256);
  if (et == 0) *flag = (atomicAdd(p.counter, 1u) == gridDim.x - 1) ? 1 : 0;
  named_bar_sync(1, 256);
  if (!*flag) return;
  auto fstamp = [&](int k) { if (p.trace && et == 0) p.trace[224 + k] = clock64(); };
  fstamp(0);
  __threadfence();
  const int np = (int)gridDim.x;                         // <= 256 partials
  const uint32_t per = ((stage_bytes / 2) / REC) & ~1u;  // records per staging buffer; even, so that every chunk starts 16-byte aligned
  const uint32_t buf_bytes = (stage_bytes / 2) & ~15u;
  const uint32_t stage0 = smem_u32(stage);
  auto issue = [&](int chunk) {                          // thread 0: bulk copy of records [chunk * per, ...) into buffer chunk & 1
    const int i0 = chunk * (int)per;
    if (i0 >= np) return;
    const int n = np - i0 < (int)per ? np - i0 : (int)per;
    const uint32_t bytes = ((uint32_t)n * REC + 15u) & ~15u;
    const uint32_t bar = fin_bar + 8u * (uint32_t)(chunk & 1);
    mbar_expect_tx(bar, bytes);
    bulk_load(stage0 + (uint32_t)(chunk & 1) * buf_bytes, reinterpret_cast<const uint8_t*>(p.part) + (size_t)i0 * REC, bytes, bar);
  };
  // the partials were written with generic-proxy stores (by every CTA, this one included) and are read by the async proxy
  if (et == 0) { asm volatile("fence.proxy.async;" ::: "memory"); issue(0); issue(1); }
  fstamp(1);
  // (m, l) of every partial: one thread each, block-wide max / fixed-order sum through shared memory
  float* wgt = scratch;                                  // [256] exp(m_i - m)
  float* red = scratch + 256;                            // [16]
  float mi = -INFINITY, li = 0.f;
  if (et < np) {
    li = __ldcg(p.part + (int64_t)et * (2 + HMAX) + 1);
    mi = li > 0.f ? __ldcg(p.part + (int64_t)et * (2 + HMAX)) : -INFINITY;
  }
  const float wm = warp_max(mi);
  if (lane == 0) red[et >> 5] = wm;
  named_bar_sync(1, 256);
  float mg = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mg = fmaxf(mg, red[i]);
  const float wi0 = li > 0.f ? expf(mi - mg) : 0.f;
  wgt[et] = wi0;
  const float ls = warp_sum(li * wi0);
  if (lane == 0) red[8 + (et >> 5)] = ls;
  named_bar_sync(1, 256);
  float lg = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) lg += red[8 + i];          // fixed order: identical in every thread
  fstamp(2);
  const int c2 = et * 2;                                 // 256 threads x 2 columns
  float2 v = make_float2(0.f, 0.f);
  const int nchunks = (np + (int)per - 1) / (int)per;
  for (int ch = 0; ch < nchunks; ++ch) {
    mbar_wait(fin_bar + 8u * (uint32_t)(ch & 1), (uint32_t)(ch >> 1) & 1u, p.err, 20);
    const int i0 = ch * (int)per;
    const int n = np - i0 < (int)per ? np - i0 : (int)per;
    const uint8_t* buf = stage + (size_t)(ch & 1) * buf_bytes;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
      const float2 q2 = *reinterpret_cast<const float2*>(buf + (size_t)i * REC + 8 + (size_t)c2 * 4);
      const float w = wgt[i0 + i];
      v.x = fmaf(q2.x, w, v.x); v.y = fmaf(q2.y, w, v.y);
    }
    named_bar_sync(1, 256);                              // everybody is done with this buffer
    if (et == 0) { fence_proxy_async(); issue(ch + 2); }
  }
  fstamp(3);
  float* pooled_s = scratch;                             // wgt is dead: [512] merged pooled vector
  if (p.rec_out) {
    *reinterpret_cast<float2*>(p.rec_out + 2 + c2) = v;
    if (et == 0) { p.rec_out[0] = mg; p.rec_out[1] = lg; }
  }
  v.x /= lg; v.y /= lg;
  *reinterpret_cast<float2*>(pooled_s + c2) = v;
  *reinterpret_cast<float2*>(p.pooled + c2) = v;
  if (et == 0) { p.stats[0] = mg; p.stats[1] = lg; *p.counter = 0u; }
  named_bar_sync(1, 256);
  if (p.logits) {
    const int wid = et >> 5;
    for (int k = wid; k < p.n_cls; k += 8) {
      float a = 0.f;
      for (int c = lane; c < HMAX; c += 32) a = fmaf(pooled_s[c], p.Wcls[k * HMAX + c], a);
      a = warp_sum(a);
      if (lane == 0) p.logits[k] = a + (p.bcls ? p.bcls[k] : 0.f);
    }
  }
  fstamp(4);
}

// ---- host helpers shared by the fused-pass translation units (defined in mil_fused_sm100.cu) ----
int make_map_2d(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                uint32_t box_cols, CUtensorMapSwizzle sw);
// the same with a leading dimension `ld` (elements) different from `cols`; rows past `rows` read as zeros
int make_map_2d_ld(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                   uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle sw);
void prof_begin(cudaStream_t stream);        // kernel-only timing hooks (mil_profile_enable)
void prof_end(cudaStream_t stream);
int debug_mask();
// pair (cta_group::2) pipeline of the fused pass, mil_fused2_sm100.cu
size_t pair_weight_image_bytes(int D, int H, int Da);
// MIL_PREC_* -> (NPROD, FP16) of the kernel templates
inline bool prec_fp16(int precision) { return precision == MIL_PREC_FP16 || precision == MIL_PREC_FP16X3; }
inline bool prec_split(int precision) { return precision == MIL_PREC_BF16X3 || precision == MIL_PREC_FP16X3; }
inline float prec_wscale(int precision) { return precision == MIL_PREC_FP16X3 ? W_SCALE_FP16X3 : 1.f; }
int pair_build_images(const float* W1, int H, int D, const float* Wa, int Da, uint8_t* w1_img, uint8_t* wa_img, int precision, cudaStream_t stream);
int pair_fused_launch(const float* X, FusedParams p, int precision, cudaStream_t stream);
int pair_plan_item(int64_t N, int D, int precision, int pair, int i, int64_t* out);
// fills the dropout fields of p from the ABI struct (nullptr / mode 0 = no dropout); <0 on a bad argument
int set_dropout(FusedParams& p, const mil_dropout_t* drop, int ncols);

}  // namespace mil
