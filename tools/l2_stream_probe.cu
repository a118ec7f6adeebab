// L2 -> SM streaming probe: how fast can every SM stream the SAME few-MB weight image out of L2 through
// bulk async copies, (0) all CTAs in lockstep, (1) skewed start offsets, (2) with cluster multicast
// (each CTA of the cluster fetches 1/CL of every chunk and multicasts it to all CL CTAs).
// The fused MLP kernel streams ~2.5 MB of weights per CTA per tile; this decides whether sharing the
// stream across a cluster lifts the L2 (LTS) throughput cap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ab/l2probe tools/l2_stream_probe.cu && ab/l2probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#ifndef STAGES
#define STAGES 6
#endif
constexpr int CHUNK = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spins = 0; !ok; ++spins) {
    if (spins > (1u << 28)) __trap();  // a protocol bug must not hang the box
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t bar, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}

// mode 0/1: unicast (CL = 1 logical).  mode 2: multicast across the CL CTAs of the cluster.
template <int CL>
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* __restrict__ w, size_t wbytes, int passes, int mode, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * CHUNK);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + STAGES * 8;
  const uint32_t rank = CL > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + i * 8, 1); mbar_init(empty0 + i * 8, mode == 2 ? CL : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CL > 1) cluster_sync();
  const int nchunks = (int)(wbytes / CHUNK);
  const long long total = (long long)nchunks * passes;
  int start = 0;
  if (mode == 1) start = (int)(((long long)blockIdx.x * nchunks) / gridDim.x);
  if (threadIdx.x == 0) {  // producer
    int idx = 0; uint32_t ph = 0; int c = start;
    for (long long i = 0; i < total; ++i) {
      mbar_wait(empty0 + idx * 8, ph ^ 1);
      const uint32_t fb = full0 + idx * 8, dst = smem_u32(smem) + idx * CHUNK;
      mbar_expect(fb, CHUNK);
      const uint8_t* src = w + (size_t)c * CHUNK;
      if (mode == 2 && CL > 1) {
        const uint32_t part = CHUNK / CL;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                dst + rank * part),
            "l"(src + rank * part), "r"(part), "r"(fb), "h"((uint16_t)((1u << CL) - 1))
            : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                     "r"((uint32_t)CHUNK), "r"(fb)
                     : "memory");
      }
      if (++c == nchunks) c = 0;
      if (++idx == STAGES) { idx = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer: release each slot as soon as it has landed
    int idx = 0; uint32_t ph = 0; unsigned long long acc = 0;
    for (long long i = 0; i < total; ++i) {
      mbar_wait(full0 + idx * 8, ph);
      acc += *reinterpret_cast<volatile uint32_t*>(smem + idx * CHUNK + 64);
      if (mode == 2 && CL > 1) {
        for (int r = 0; r < CL; ++r) remote_arrive(empty0 + idx * 8, r);
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + idx * 8) : "memory");
      }
      if (++idx == STAGES) { idx = 0; ph ^= 1; }
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
  }
  __syncthreads();
  if (CL > 1) cluster_sync();
}

template <int CL>
static void run(const char* name, const uint8_t* w, size_t wbytes, int passes, int mode, int nctas, unsigned long long* sink) {
  const int smem_bytes = STAGES * CHUNK + 256;
  cudaFuncSetAttribute(stream_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(stream_kernel<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nctas / CL * CL);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int maxc = 0;
  cudaOccupancyMaxActiveClusters(&maxc, stream_kernel<CL>, &cfg);
  if (CL > 1 && maxc * CL < (int)cfg.gridDim.x) cfg.gridDim = dim3(maxc * CL);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, stream_kernel<CL>, w, wbytes, passes, mode, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double per_cta = (double)(wbytes / CHUNK) * CHUNK * passes;
  const double agg = per_cta * cfg.gridDim.x / (best * 1e-3) / 1e12;
  printf("%-34s ctas %3d  %.3f ms  delivered %.2f TB/s aggregate, %.1f GB/s per SM\n", name, cfg.gridDim.x, best, agg,
         per_cta / (best * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
  const size_t wbytes = (argc > 1 ? atol(argv[1]) : 4864) * 1024ull;  // 4.75 MB: one MLP's operand image
  const int passes = argc > 2 ? atoi(argv[2]) : 40;
  uint8_t* w; unsigned long long* sink;
  cudaMalloc(&w, wbytes); cudaMemset(w, 1, wbytes); cudaMalloc(&sink, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d, image %zu KB, passes %d\n", sms, wbytes / 1024, passes);
  if (argc > 3) {  // power mode: one configuration, long enough for nvidia-smi to sample
    const int which = atoi(argv[3]);
    if (which == 0) run<1>("unicast lockstep", w, wbytes, passes, 0, sms, sink);
    if (which == 1) run<4>("cluster 4 multicast", w, wbytes, passes, 2, sms, sink);
    if (which == 2) run<2>("cluster 2 multicast", w, wbytes, passes, 2, sms, sink);
    return cudaDeviceSynchronize() != cudaSuccess;
  }
  run<1>("unicast lockstep", w, wbytes, passes, 0, sms, sink);
  run<1>("unicast skewed", w, wbytes, passes, 1, sms, sink);
  run<2>("cluster 2 unicast skewed", w, wbytes, passes, 1, sms, sink);
  run<2>("cluster 2 multicast", w, wbytes, passes, 2, sms, sink);
  run<4>("cluster 4 multicast", w, wbytes, passes, 2, sms, sink);
  run<8>("cluster 8 multicast", w, wbytes, passes, 2, sms, sink);
  run<1>("unicast lockstep, half the SMs", w, wbytes, passes, 0, sms / 2, sink);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
