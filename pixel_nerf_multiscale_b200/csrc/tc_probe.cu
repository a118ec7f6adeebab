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

}  // namespace pnr

using namespace pnr;

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
  set_err("probe: unknown mode %d", mode);
  return PNR_ERR_BAD_ARG;
}
