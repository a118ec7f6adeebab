// Thin inline-PTX wrappers for the sm_100a primitives the tensor-core MLP uses:
// mbarrier, bulk async copies (UBLKCP), tcgen05 alloc / mma / commit / ld / st, cluster helpers.
// Everything here is sm_100a-only; there is no fallback path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pnr {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of `local` mapped into CTA `rank`'s window of the cluster address space
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}

// one lane of a converged warp (warp-uniform control flow around it keeps operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier that lives in CTA `rank` of the cluster (may be this CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_local, uint32_t rank) {
  uint32_t remote = mapa(bar_local, rank);
  // default (.release.cta) semantics, like CUTLASS's ClusterBarrier::arrive(cta_id): an explicit
  // .release.cluster costs a full memory barrier (ERRBAR/MEMBAR, >1000 cycles) per arrive, which
  // serialised the peer->leader relay to one ring slot per ~1300 cycles in the first version.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware may park the warp for up to `ns` nanoseconds (it is
// woken as soon as the phase completes), instead of returning after its short default interval.
// Waiting warps then execute almost no instructions: in the first profile of the fused kernel ~40 %
// of all executed instructions were this spin loop, which costs clock under the power cap.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU.  On timeout the barrier id is recorded in
// *err (global) and the wait returns; the kernel then runs to completion with garbage results and
// the host reports the failure.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Slow path kept OUT OF LINE: the kernels wait at ~40 sites, and the instruction footprint of the
// role loops matters (the L1.5 I-cache is 32 KB; code beyond it streams from an L2 that is busy with
// the weight stream).
// err[0] = fault tag of this device (0 = none), err[1] = timeout in ms (0 = wait for ever: debuggers,
// compute-sanitizer, time-sliced GPUs; PNR_WAIT_TIMEOUT_MS)
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* err, int tag) {
  unsigned long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait_hint(bar, parity, 100000u)) return;
    if ((it & 15) == 0) {
      if (err && *(volatile int*)err != 0) return;
      const unsigned long long limit = err ? (unsigned long long)(*(volatile int*)(err + 1)) * 1000000ull : 2000000000ull;
      if (limit == 0) continue;
      unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > limit) {
        // *err lives in mapped pinned host memory: the tag survives the trap (which is the loud,
        // deterministic way out -- draining a half-synchronised tcgen05 pipeline is not safe)
        if (err) {
          atomicCAS(err, 0, tag);
          __threadfence_system();
        }
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int tag) {
  // try_wait itself suspends the thread for a hardware-bounded interval, so this loop does not
  // spin hot.  (Default .acquire.cta semantics on purpose: a cluster-scope acquire makes ptxas emit
  // CCTL.IVALL -- a full L1 invalidate -- after every probe, which dominated the first profile.)
  // Wall-clock bound (err[1] ms, default 2 s) without progress records `tag` in *err; once *err is set every
  // wait in the grid gives up immediately, so a protocol bug drains the kernel in about that time.
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, err, tag);
}

// ---- async proxy / bulk copy ---------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// global -> this CTA's shared memory, completion (bytes) on a LOCAL mbarrier.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// 2-D tiled TMA load issued by either CTA of a cta_group::2 pair: data lands in THIS CTA's shared
// memory, the transaction bytes complete on the LEADER (even) CTA's mbarrier at the same offset
// (bit 24 of a shared::cluster address selects the CTA of the pair).  SASS: UTMALDG.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst_smem),
      "l"(tmap), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
  if constexpr (CG == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved" 8x16B core matrices):
//   core matrix = 8 rows x 16 bytes stored contiguously (128 B)
//   LBO = byte distance between the two core matrices along K of one MMA (K=16 bf16 = 2 x 16 B)
//   SBO = byte distance between consecutive 8-row groups along M/N
// bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1 (sm_100), [61,64) layout=0
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor for kind::f16: bf16 x bf16 -> f32 (or f16 x f16 -> f32), both operands K-major.
// a_format / b_format: 0 = f16, 1 = bf16.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int N) {
  return (1u << 4) /*D=f32*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on `bar` (same smem offset in every CTA of `mask`) once all MMAs issued so far by this
// thread have completed; implies tcgen05.fence::before_thread_sync.
template <int CG>
__device__ __forceinline__ void mma_commit(uint32_t bar, uint16_t mask) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}

// TMEM -> registers: 32 lanes (this warp's quadrant) x 32 consecutive columns, one row per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same shape
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// f16 pair, round-to-nearest, saturating at +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// operand element format of the tensor-core path: 0 = bf16, 1 = f16 (same MMA rate; f16 carries 11
// significand bits instead of 8, in exchange for range: activations and features must stay below 65504)
template <int FMT> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (FMT == 1) return pack_f16x2(lo, hi); else return pack_bf16x2(lo, hi);
}

}  // namespace ptx
}  // namespace pnr
