// Hardware probes for the tcgen05 building blocks (test infrastructure compiled into the library,
// exported as pnr_tc_probe; used by tests/test_gpu_tc_probe.py).  They exercise exactly the operand
// layouts, descriptors, barriers and TMEM addressing the fused MLP kernel relies on and dump the raw
// TMEM image, so a wrong assumption shows up as a readable pattern instead of a wrong render.
//
//   mode 1: cta_group::1, M=128, N=256, one CTA.  A (128 x K), B (256 x K) bf16, K-major, SWIZZLE_NONE.
//   mode 2: cta_group::2, M=128 (64 rows per CTA), N=256 (128 B-rows per CTA), cluster of 2.
// Operands are staged with 1-D bulk copies (UBLKCP) completing on mbarriers, like the real kernel.
// Global operand images are pre-arranged by the host in panel order:
//   A image per CTA: [K/8][rows][8] bf16   (rows = 128 or 64),  B image per CTA: [K/8][nrows][8] bf16.
// Output: D_raw[cta][lane 128][col 256 or 128] fp32 = raw TMEM contents.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pnr {
using namespace ptx;

template <int CG, bool REMOTE = false>
__global__ void __launch_bounds__(128) probe_gemm_kernel(const __nv_bfloat16* __restrict__ A,
                                                         const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                                                         int K, int* __restrict__ err) {
  constexpr int ROWS_A = (CG == 1) ? 128 : 64;   // A rows held by this CTA
  constexpr int ROWS_B = (CG == 1) ? 256 : 128;  // B rows held by this CTA
  constexpr int NCOLS = (CG == 1) ? 256 : 128;   // TMEM columns written per CTA
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)ROWS_A * K * 2;
  uint64_t* bars = (uint64_t*)(sB + (size_t)ROWS_B * K * 2);  // [0]=full(local) [1]=peer_full [2]=mma_done
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_peer = smem_u32(&bars[1]), bar_done = smem_u32(&bars[2]);

  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_peer, 1);
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<CG>(smem_u32(tmem_slot), 256);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const uint32_t bytesA = ROWS_A * K * 2, bytesB = ROWS_B * K * 2;
  if (REMOTE) {
    // experiment: the peer's bulk copies complete on the LEADER's barrier (no relay hop)
    if (threadIdx.x == 0) {
      if (rank == 0) mbar_expect_tx(bar_full, 2 * (bytesA + bytesB));
      uint32_t lead_bar = mapa(bar_full, 0);
      bulk_g2s(smem_u32(sA), A + (size_t)rank * ROWS_A * K, bytesA, rank == 0 ? bar_full : lead_bar);
      bulk_g2s(smem_u32(sB), B + (size_t)rank * ROWS_B * K, bytesB, rank == 0 ? bar_full : lead_bar);
    }
  } else if (threadIdx.x == 0) {
    mbar_expect_tx(bar_full, bytesA + bytesB);
    bulk_g2s(smem_u32(sA), A + (size_t)rank * ROWS_A * K, bytesA, bar_full);
    bulk_g2s(smem_u32(sB), B + (size_t)rank * ROWS_B * K, bytesB, bar_full);
  }
  if (!REMOTE && CG == 2 && rank == 1 && threadIdx.x == 32) {
    // relay: peer operands landed -> tell the leader
    mbar_wait(bar_full, 0, err, 101);
    mbar_arrive_cluster(bar_peer, 0);
  }
  if (rank == 0 && threadIdx.x == 32) {
    mbar_wait(bar_full, 0, err, 102);
    if (CG == 2 && !REMOTE) mbar_wait(bar_peer, 0, err, 103);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16_f32(128, 256);
    for (int k = 0; k < K / 16; ++k) {
      // one MMA consumes 2 core-matrix columns (2 x 8 bf16 along K)
      uint64_t da = smem_desc(smem_u32(sA) + k * 2 * ROWS_A * 16, ROWS_A * 16, 128);
      uint64_t db = smem_desc(smem_u32(sB) + k * 2 * ROWS_B * 16, ROWS_B * 16, 128);
      mma_bf16<CG>(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit<CG>(bar_done, 0x3);
  }
  __syncwarp();
  mbar_wait(bar_done, 0, err, 104);
  tc_fence_after();
  // dump: warp w reads lanes [32w, 32w+32)
  float* out = D + ((size_t)rank * 128 + warp * 32 + lane) * NCOLS;
  for (int c0 = 0; c0 < NCOLS; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) tmem_dealloc<CG>(tmem, 256);
}

// mode 4/5: like mode 2 but the A operand is read from TENSOR MEMORY (tcgen05.mma [d], [a], b-desc): which
// TMEM lanes / columns does the hardware read as A[row][k] when each CTA of the pair owns 64 rows?
// Every thread fills its lane of the A region (columns 256..256+K/2) with packed bf16 pairs that
// identify (mode 4) the lane, 128*rank + lane, or (mode 5) the position 2*column + half; the host passes
// a selector B (B[n][k] = [n == k]), so D[row][n] shows what was read as A[row][n].
__global__ void __launch_bounds__(128) probe_ts_kernel(const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int K,
                                                       int pattern, int* __restrict__ err) {
  constexpr int ROWS_B = 128, NCOLS = 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_ctarank();
  uint8_t* sB = smem;
  uint64_t* bars = (uint64_t*)(sB + (size_t)ROWS_B * K * 2);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_peer = smem_u32(&bars[1]), bar_done = smem_u32(&bars[2]);
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_peer, 1);
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<2>(smem_u32(tmem_slot), 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  {  // A region: 32 columns (K <= 64) at column 256, all 128 lanes
    uint32_t r[32];
    const int L = warp * 32 + lane;
    for (int c = 0; c < 32; ++c) {
      float lo = pattern == 0 ? (float)(128 * (int)rank + L) : (float)(2 * c);
      float hi = pattern == 0 ? lo : (float)(2 * c + 1);
      r[c] = pack_bf16x2(lo, hi);
    }
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256, r);
    tmem_st_wait();
  }
  const uint32_t bytesB = ROWS_B * K * 2;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_full, bytesB);
    bulk_g2s(smem_u32(sB), B + (size_t)rank * ROWS_B * K, bytesB, bar_full);
  }
  tc_fence_before();
  cluster_sync();  // both CTAs' A regions are written
  tc_fence_after();
  if (rank == 1 && threadIdx.x == 32) {
    mbar_wait(bar_full, 0, err, 111);
    mbar_arrive_cluster(bar_peer, 0);
  }
  if (rank == 0 && threadIdx.x == 32) {
    mbar_wait(bar_full, 0, err, 112);
    mbar_wait(bar_peer, 0, err, 113);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16_f32(128, 256);
    for (int k = 0; k < K / 16; ++k) {
      uint64_t db = smem_desc(smem_u32(sB) + k * 2 * ROWS_B * 16, ROWS_B * 16, 128);
      uint32_t a_tmem = tmem + 256 + k * 8;  // hypothesis: 16 bf16 of K = 8 columns
      uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
          "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    mma_commit<2>(bar_done, 0x3);
    if (pattern >= 2) {
      // timing: 1024 back-to-back K=16 MMAs (SS: A = first 64 rows of this CTA's B panel; TS: A from TMEM)
      mbar_wait(bar_done, 0, err, 115);
      tc_fence_after();
      const uint64_t db = smem_desc(smem_u32(sB), ROWS_B * 16, 128);
      const uint64_t da = smem_desc(smem_u32(sB), ROWS_B * 16, 128);  // rows 0..63 of the same panel (LBO = 128 rows)
      const long long t0 = clock64();
      for (int it = 0; it < 1024; ++it) {
        if (pattern == 2)
          mma_bf16<2>(tmem, da, db, idesc, 1u);
        else
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
              "r"(tmem + 256), "l"(db), "r"(idesc), "r"(1u)
              : "memory");
      }
      const long long t1 = clock64();
      mma_commit<2>(bar_peer, 0x1);  // completes on the leader's (re-used) peer barrier: phase 1
      mbar_wait(bar_peer, 1, err, 116);
      const long long t2 = clock64();
      D[0] = (float)(t1 - t0);
      D[1] = (float)(t2 - t0);
    }
  }
  __syncwarp();
  mbar_wait(bar_done, 0, err, 114);
  tc_fence_after();
  if (pattern >= 2) {
    tc_fence_before();
    cluster_sync();
    if (warp == 1) tmem_dealloc<2>(tmem, 512);
    return;
  }
  float* out = D + ((size_t)rank * 128 + warp * 32 + lane) * NCOLS;
  for (int c0 = 0; c0 < NCOLS; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 1) tmem_dealloc<2>(tmem, 512);
}

// Issue-rate probe: `iters` back-to-back tcgen05.mma (K=16, bf16 -> fp32, accumulate) of one shape, operands
// resident in shared memory (contents irrelevant), cycles on the issuing thread until the commit completes.
//   CG: cta_group; M, N: instruction shape (M = rows of the whole CTA group).  ts != 0: A from tensor memory.
template <int CG>
__global__ void __launch_bounds__(128) probe_rate_kernel(int M, int Nn, int ts, int iters, int alt, float* __restrict__ out,
                                                         int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  uint64_t* bars = (uint64_t*)(smem + 256 * 16 * 2 * 2);  // operands: 256 rows x K=16 (two 8-wide panels)
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5;
  const uint32_t bar_done = smem_u32(&bars[0]);
  for (int i = threadIdx.x; i < 256 * 16 * 2 * 2 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    tmem_alloc<CG>(smem_u32(tmem_slot), 512);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (rank == 0 && threadIdx.x == 32) {
    const uint32_t idesc = idesc_bf16_f32(M, Nn);
    const uint64_t da = smem_desc(smem_u32(smem), 256 * 16, 128);
    const uint64_t db = smem_desc(smem_u32(smem), 256 * 16, 128);
    const long long t0 = clock64();
    const uint32_t dstride = (uint32_t)((CG == 2 && M == 128) ? Nn / 2 : Nn);  // TMEM columns of one accumulator tile
    for (int it = 0; it < iters; ++it) {
      if (!ts) {
        mma_bf16<CG>(tmem + (uint32_t)(it % alt) * dstride, da, db, idesc, 1u);  // alt > 1: rotate independent accumulators
      } else if (CG == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 256), "l"(db),
                     "r"(idesc), "r"(1u) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 256), "l"(db),
                     "r"(idesc), "r"(1u) : "memory");
      }
    }
    mma_commit<CG>(bar_done, CG == 2 ? 0x3 : 0x1);
    mbar_wait(bar_done, 0, err, 120);
    out[0] = (float)(clock64() - t0);
  } else if (threadIdx.x == 32) {
    mbar_wait(bar_done, 0, err, 121);
  }
  __syncwarp();
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) tmem_dealloc<CG>(tmem, 512);
}

}  // namespace pnr

using namespace pnr;

// cycles for `iters` MMAs of shape (cta_group, M, N, K=16); *out on the device.  Test/measurement helper.
extern "C" int pnr_tc_rate_probe(int cta_group, int M, int Nn, int ts, int iters, int alt, float* out, int* err, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PNR_CHECK_ARG((cta_group == 1 || cta_group == 2) && iters > 0 && Nn >= 16 && Nn <= 256 && Nn % 16 == 0, "rate probe: bad arguments");
  PNR_CHECK_ARG(cta_group == 1 ? (M == 64 || M == 128) : (M == 128 || M == 256), "rate probe: bad M");
  PNR_CHECK_ARG(alt >= 1 && alt * ((cta_group == 2 && M == 128) ? Nn / 2 : Nn) <= 512, "rate probe: accumulators do not fit tensor memory");
  PNR_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  const size_t smem = 256 * 16 * 2 * 2 + 128;
  cudaLaunchConfig_t cfg;
  memset((void*)&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cta_group);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cta_group == 1) PNR_CUDA(cudaLaunchKernelEx(&cfg, probe_rate_kernel<1>, M, Nn, ts, iters, alt, out, err));
  else PNR_CUDA(cudaLaunchKernelEx(&cfg, probe_rate_kernel<2>, M, Nn, ts, iters, alt, out, err));
  pnr::launch_counter()++;
  return PNR_OK;
}

extern "C" int pnr_tc_probe(int mode, const void* A, const void* B, float* D, int K, int* err, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PNR_CHECK_ARG(K % 16 == 0 && K >= 16 && K <= 256, "probe: K must be a multiple of 16 in [16,256]");
  PNR_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  if (mode == 1) {
    size_t smem = (size_t)(128 + 256) * K * 2 + 64;
    PNR_CUDA(cudaFuncSetAttribute(probe_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_gemm_kernel<1><<<1, 128, smem, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, K, err);
    PNR_LAUNCHED();
    return PNR_OK;
  }
  if (mode == 2 || mode == 3) {
    size_t smem = (size_t)(64 + 128) * K * 2 + 64;
    auto kern = mode == 2 ? probe_gemm_kernel<2, false> : probe_gemm_kernel<2, true>;
    PNR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset((void*)&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PNR_CUDA(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, K, err));
    pnr::launch_counter()++;
    return PNR_OK;
  }
  if (mode >= 4 && mode <= 7) {
    PNR_CHECK_ARG(K <= 64, "probe: TS mode takes K <= 64");
    size_t smem = (size_t)128 * K * 2 + 64;
    PNR_CUDA(cudaFuncSetAttribute(probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset((void*)&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PNR_CUDA(cudaLaunchKernelEx(&cfg, probe_ts_kernel, (const __nv_bfloat16*)B, D, K, mode - 4, err));
    pnr::launch_counter()++;
    return PNR_OK;
  }
  set_err("probe: unknown mode %d", mode);
  return PNR_ERR_BAD_ARG;
}
