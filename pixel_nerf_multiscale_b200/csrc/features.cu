// Kernel (a), fp32 validation variant, and the NCHW -> NHWC pyramid packer.
#include <cuda_fp16.h>

#include <type_traits>

#include "features.cuh"

namespace pnr {

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC (fp32 | bf16).  One block per (view, y): a W x C slab is transposed through
// shared memory so that both the read (x fastest) and the write (c fastest) are coalesced.
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void pack_level_kernel(const float* __restrict__ src, int C, int H, int W, OutT* __restrict__ dst) {
  __shared__ float tile[32][33];
  int vy = blockIdx.z;  // view*H + y
  int view = vy / H, y = vy % H;
  int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, x = x0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && x < W) ? src[(((size_t)view * C + c) * H + y) * W + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int x = x0 + i, c = c0 + threadIdx.x;
    if (x < W && c < C) {
      float v = tile[threadIdx.x][i];
      size_t o = (((size_t)view * H + y) * W + x) * C + c;
      if constexpr (sizeof(OutT) == 4)
        dst[o] = v;
      else if constexpr (std::is_same<OutT, __half>::value)
        dst[o] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
      else
        dst[o] = __float2bfloat16(v);
    }
  }
}

int launch_pack_level(const float* src, int n_views, int C, int H, int W, void* dst, int dtype, cudaStream_t st) {
  dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(C, 32), n_views * H);
  if (dtype == PNR_FP32)
    pack_level_kernel<float><<<grid, block, 0, st>>>(src, C, H, W, (float*)dst);
  else if (dtype == PNR_FP16)
    pack_level_kernel<__half><<<grid, block, 0, st>>>(src, C, H, W, (__half*)dst);
  else
    pack_level_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(src, C, H, W, (__nv_bfloat16*)dst);
  PNR_LAUNCHED();
  return PNR_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32 rows of the ResnetFC input in reference order:  zx[(sb*NS+v)*P + p] = [latent | code]
// One warp per row; lanes stride over channels (NHWC -> 128 B coalesced per tap).
// ---------------------------------------------------------------------------------------------
template <typename FeatT>
__global__ void __launch_bounds__(256)
point_features_f32_kernel(const pnr_scene sc, const float* __restrict__ xyz, const float* __restrict__ viewdirs,
                          const float* __restrict__ rays, const float* __restrict__ z, int K, int SB, int P,
                          float* __restrict__ zx) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long rows = (long long)SB * sc.ns * P;
  if (row >= rows) return;
  const int sb = (int)(row / ((long long)sc.ns * P));
  const long long rem = row - (long long)sb * sc.ns * P;
  const int v = (int)(rem / P);
  const int p = (int)(rem - (long long)v * P);
  const int view = sb * sc.ns + v;
  float X[3], D[3];
  load_point(xyz, viewdirs, rays, z, K, (long long)sb * P + p, X, D);
  PointCam pc;
  camera_project(sc.cams + view * 16, X, D, pc);
  const int width = sc.d_latent + sc.d_in;
  float* out = zx + row * width;
  for (int l = 0; l < sc.n_levels; ++l) {
    const int C = sc.C[l], H = sc.H[l], W = sc.W[l];
    Taps t = make_taps(pc.u, pc.v, H, W, sc.kx[l], sc.ky[l]);
    const FeatT* f = (const FeatT*)sc.level[l] + (size_t)view * H * W * C;
    for (int c = lane; c < C; c += 32) {
      float a, b, cc, d;
      if constexpr (sizeof(FeatT) == 4) {
        a = f[(size_t)t.o00 * C + c];
        b = f[(size_t)t.o01 * C + c];
        cc = f[(size_t)t.o10 * C + c];
        d = f[(size_t)t.o11 * C + c];
      } else {
        a = __bfloat162float(f[(size_t)t.o00 * C + c]);
        b = __bfloat162float(f[(size_t)t.o01 * C + c]);
        cc = __bfloat162float(f[(size_t)t.o10 * C + c]);
        d = __bfloat162float(f[(size_t)t.o11 * C + c]);
      }
      // same accumulation order as grid_sample: nw, ne, sw, se
      float acc = a * t.w00;
      acc += b * t.w01;
      acc += cc * t.w10;
      acc += d * t.w11;
      out[sc.ch_off[l] + c] = acc;
    }
  }
  for (int j = lane; j < sc.d_in; j += 32) out[sc.d_latent + j] = code_entry(sc, pc, j);
}

int launch_point_features_f32(const pnr_scene& sc, const float* xyz, const float* viewdirs, const float* rays,
                              const float* z, int K, int SB, int P, float* zx, cudaStream_t st) {
  long long rows = (long long)SB * sc.ns * P;
  if (rows == 0) return PNR_OK;
  int wpb = 8;
  long long blocks = ceil_div_ll(rows, wpb);
  PNR_CHECK_ARG(blocks < 2147483647LL, "point_features: too many rows (%lld)", rows);
  PNR_UNSUPPORTED(sc.feat_dtype == PNR_FP16, "point_features_f32 reads fp32 or bf16 maps");
  if (sc.feat_dtype == PNR_FP32)
    point_features_f32_kernel<float><<<(unsigned)blocks, wpb * 32, 0, st>>>(sc, xyz, viewdirs, rays, z, K, SB, P, zx);
  else
    point_features_f32_kernel<__nv_bfloat16><<<(unsigned)blocks, wpb * 32, 0, st>>>(sc, xyz, viewdirs, rays, z, K, SB, P, zx);
  PNR_LAUNCHED();
  return PNR_OK;
}

}  // namespace pnr
