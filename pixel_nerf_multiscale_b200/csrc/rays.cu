// Kernel family (c): per-ray stratified sampling, alpha compositing, inverse-CDF importance
// sampling, depth-guided samples and the per-ray sort  (src/render/nerf.py:98-161, 178-244, 286-295).
// One warp per ray; every array is (B, K) fp32 row-major so a warp reads/writes 128 B lines.
// All arithmetic is fp32 with the reference's operation order; products/sums that torch performs
// as separate ops use __fmul_rn/__fadd_rn so the compiler cannot contract them into FMAs.
#include <math_constants.h>

#include "common.cuh"

namespace pnr {

__device__ __forceinline__ float lerp_depth(float near, float far, float s, int lindisp) {
  if (!lindisp)  // near*(1-s) + far*s            nerf.py:113, 145
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, s)), __fmul_rn(far, s));
  // 1 / (1/near*(1-s) + 1/far*s)                  nerf.py:115, 147
  float a = __fmul_rn(__fdiv_rn(1.f, near), __fsub_rn(1.f, s));
  float b = __fmul_rn(__fdiv_rn(1.f, far), s);
  return __fdiv_rn(1.f, __fadd_rn(a, b));
}

// ---- sample_coarse ----------------------------------------------------------------------------
__global__ void sample_coarse_kernel(const float* __restrict__ rays, const float* __restrict__ jitter, long long n,
                                     int Kc, int lindisp, float lin_end, float lin_step, float step,
                                     float* __restrict__ z) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long r = i / Kc;
  int k = (int)(i - r * Kc);
  // torch.linspace(0, 1-step, Kc): first half start + step*k, second half end - step*(Kc-1-k).  torch evaluates the
  // second form as ONE fused multiply-add on the CPU (AVX2/AVX-512 kernels) and on CUDA alike; with a separate
  // multiply the last bit differs whenever step = (1-1/Kc)/(Kc-1) is inexact (any Kc that is not a power of two)
  float s = (k < Kc / 2) ? __fmul_rn(lin_step, (float)k) : __fmaf_rn(-lin_step, (float)(Kc - 1 - k), lin_end);
  s = __fadd_rn(s, __fmul_rn(jitter[i], step));
  z[i] = lerp_depth(rays[r * 8 + 6], rays[r * 8 + 7], s, lindisp);
}

int launch_sample_coarse(const float* rays, const float* jitter, int B, int Kc, int lindisp, float* z,
                         cudaStream_t st) {
  long long n = (long long)B * Kc;
  if (n == 0) return PNR_OK;
  float step = (float)(1.0 / Kc);
  float lin_end = (float)(1.0 - 1.0 / Kc);
  float lin_step = Kc > 1 ? lin_end / (float)(Kc - 1) : 0.f;
  sample_coarse_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(rays, jitter, n, Kc, lindisp, lin_end, lin_step,
                                                                     step, z);
  PNR_LAUNCHED();
  return PNR_OK;
}

// ---- gen_rays: rays of N pinhole cameras, src/util/util.py:118-148,243-281 ----------
// unproj = normalize((x-cx)/fx, -(y-cy)/fy, -1); dir = R * unproj; origin = t; near/far constants.
__global__ void gen_rays_kernel(const float* __restrict__ poses, int N, int W, int H, float fx, float fy, float cx,
                                float cy, float near, float far, float* __restrict__ rays) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H * W;
  if (i >= total) return;
  int x = (int)(i % W);
  long long t = i / W;
  int y = (int)(t % H);
  int n = (int)(t / H);
  const float* P = poses + (size_t)n * 16;
  float X = __fdiv_rn(__fsub_rn((float)x, cx), fx);
  float Y = -__fdiv_rn(__fsub_rn((float)y, cy), fy);
  float Z = -1.f;
  float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(X, X), __fmul_rn(Y, Y)), 1.f));
  X = __fdiv_rn(X, nrm);
  Y = __fdiv_rn(Y, nrm);
  Z = __fdiv_rn(Z, nrm);
  float4 a, b;
  a.x = P[3];
  a.y = P[7];
  a.z = P[11];
  a.w = P[0] * X + P[1] * Y + P[2] * Z;
  b.x = P[4] * X + P[5] * Y + P[6] * Z;
  b.y = P[8] * X + P[9] * Y + P[10] * Z;
  b.z = near;
  b.w = far;
  reinterpret_cast<float4*>(rays)[2 * i] = a;
  reinterpret_cast<float4*>(rays)[2 * i + 1] = b;
}

int launch_gen_rays(const float* poses, int N, int W, int H, float fx, float fy, float cx, float cy, float near,
                    float far, float* rays, cudaStream_t st) {
  long long total = (long long)N * H * W;
  if (total == 0) return PNR_OK;
  gen_rays_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(poses, N, W, H, fx, fy, cx, cy, near, far, rays);
  PNR_LAUNCHED();
  return PNR_OK;
}

// ---- output side: clamp + uint8 quantisation + squared-error accumulation on device ---
// u8 = trunc(clamp(rgb,0,1)*255) as eval/gen_video.py:226 does with numpy; sse += (clamp(rgb)-gt)^2 in fp64
// (eval/eval.py:278-300 computes PSNR from the clamped image), no host synchronisation.
__global__ void finalize_rgb_kernel(const float* __restrict__ rgb, const float* __restrict__ gt, long long n,
                                    uint8_t* __restrict__ u8, double* __restrict__ sse) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (i < n) {
    float v = fminf(fmaxf(rgb[i], 0.f), 1.f);
    if (u8) u8[i] = (uint8_t)(v * 255.f);
    if (gt) {
      float d = v - gt[i];
      e = (double)d * (double)d;
    }
  }
  if (sse) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(sse, e);
  }
}

int launch_finalize_rgb(const float* rgb, const float* gt, long long n, uint8_t* u8, double* sse, cudaStream_t st) {
  if (n == 0) return PNR_OK;
  finalize_rgb_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(rgb, gt, n, u8, sse);
  PNR_LAUNCHED();
  return PNR_OK;
}

// ---- per-view image metrics (eval/eval.py:314-343) --------------------------------------------------
// SSIM as skimage.measure.compare_ssim(a, b, multichannel=True, data_range=R) computes it (defaults:
// 7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03; per channel, mean over the image with
// the (win-1)/2 border cropped, then mean over channels), and the squared error for compare_psnr.
// Images are (NV,H,W,C) fp32; a is clamped to [0,1] first like the driver does.  One thread per
// interior pixel and channel, fp64 window sums (the variance is a difference of nearly equal
// numbers), block reduction, one atomicAdd per block into sums[view*2 + {0: ssim, 1: sq. error}].
__global__ void __launch_bounds__(256)
frame_metrics_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int C, int win,
                     double c1, double c2, double* __restrict__ sums) {
  const int view = blockIdx.y;
  const int pad = (win - 1) / 2;
  const int Hi = H - 2 * pad, Wi = W - 2 * pad;
  const long long n_in = (long long)(Hi > 0 ? Hi : 0) * (Wi > 0 ? Wi : 0) * C;
  const long long n_all = (long long)H * W * C;
  const float* av = a + (size_t)view * n_all;
  const float* bv = b + (size_t)view * n_all;
  double ssim = 0.0, se = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += (long long)gridDim.x * blockDim.x) {
    const float d = fminf(fmaxf(av[i], 0.f), 1.f) - bv[i];
    se += (double)d * (double)d;
    if (i < n_in) {
      const int c = (int)(i % C);
      const int x = (int)((i / C) % Wi), y = (int)(i / ((long long)C * Wi));
      double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
      for (int dy = 0; dy < win; ++dy) {
        const size_t row = ((size_t)(y + dy) * W + x) * C + c;
        for (int dx = 0; dx < win; ++dx) {
          const double p = (double)fminf(fmaxf(av[row + (size_t)dx * C], 0.f), 1.f), q = (double)bv[row + (size_t)dx * C];
          sx += p; sy += q; sxx += p * p; syy += q * q; sxy += p * q;
        }
      }
      const double np = (double)win * win, cov = np / (np - 1.0);
      const double ux = sx / np, uy = sy / np;
      const double vx = cov * (sxx / np - ux * ux), vy = cov * (syy / np - uy * uy), vxy = cov * (sxy / np - ux * uy);
      ssim += ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    }
  }
  __shared__ double sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
    se += __shfl_xor_sync(0xffffffffu, se, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ssim; sh[1][threadIdx.x >> 5] = se; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    atomicAdd(&sums[(size_t)view * 2 + threadIdx.x], t);
  }
}

int launch_frame_metrics(const float* a, const float* b, int NV, int H, int W, int C, int win, float data_range,
                         double* sums, cudaStream_t st) {
  if (NV == 0) return PNR_OK;
  const long long n_all = (long long)H * W * C;
  unsigned bx = (unsigned)ceil_div_ll(n_all, 256);
  if (bx > 2048) bx = 2048;
  const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
  frame_metrics_kernel<<<dim3(bx, (unsigned)NV), 256, 0, st>>>(a, b, H, W, C, win, c1, c2, sums);
  PNR_LAUNCHED();
  return PNR_OK;
}

// ---- composite --------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
composite_kernel(const float* __restrict__ rays, const float* __restrict__ z, const float4* __restrict__ out4, int B,
                 int K, int white, float* __restrict__ weights, float* __restrict__ rgb, float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B) return;
  const float far = rays[(size_t)r * 8 + 7];
  const float* zr = z + (size_t)r * K;
  const float4* o4 = out4 + (size_t)r * K;
  float carry = 1.f;  // T at the start of this 32-sample chunk
  float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f;
  for (int k0 = 0; k0 < K; k0 += 32) {
    int k = k0 + lane;
    bool ok = k < K;
    float zk = ok ? zr[k] : 0.f;
    float zn = (k + 1 < K) ? zr[k + 1] : far;  // last delta = far - z_K      nerf.py:181
    float4 o = ok ? o4[k] : make_float4(0.f, 0.f, 0.f, 0.f);
    float delta = __fsub_rn(zn, zk);
    // alpha = 1 - exp(-delta * relu(sigma))                                   nerf.py:228
    float alpha = ok ? __fsub_rn(1.f, expf(-__fmul_rn(delta, fmaxf(o.w, 0.f)))) : 0.f;
    float f = ok ? __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.f;  // 1 - alpha + 1e-10   :230
    // inclusive product scan over the chunk
    float p = f;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      float q = __shfl_up_sync(0xffffffffu, p, off);
      if (lane >= off) p = __fmul_rn(p, q);
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.f;
    float T = __fmul_rn(carry, excl);  // cumprod([1, f_0, f_1, ...])[k]           :231-234
    float w = __fmul_rn(alpha, T);     // weights = alpha * T[:-1]                  :235
    carry = __fmul_rn(carry, __shfl_sync(0xffffffffu, p, 31));
    if (ok) {
      if (weights) weights[(size_t)r * K + k] = w;
      ar += w * o.x;
      ag += w * o.y;
      ab += w * o.z;
      ad += w * zk;
      aw += w;
    }
  }
  ar = warp_sum(ar);
  ag = warp_sum(ag);
  ab = warp_sum(ab);
  ad = warp_sum(ad);
  aw = warp_sum(aw);
  if (lane == 0) {
    if (white) {  // rgb + 1 - pix_alpha                                          :241-244
      ar = ar + 1.f - aw;
      ag = ag + 1.f - aw;
      ab = ab + 1.f - aw;
    }
    rgb[(size_t)r * 3 + 0] = ar;
    rgb[(size_t)r * 3 + 1] = ag;
    rgb[(size_t)r * 3 + 2] = ab;
    depth[r] = ad;
  }
}

int launch_composite(const float* rays, const float* z, const float* rgb_sigma, int B, int K, int white,
                     float* weights, float* rgb, float* depth, cudaStream_t st) {
  if (B == 0) return PNR_OK;
  PNR_CHECK_ARG(K >= 1, "composite: K must be >= 1");
  composite_kernel<<<ceil_div(B, 8), 256, 0, st>>>(rays, z, (const float4*)rgb_sigma, B, K, white, weights, rgb,
                                                   depth);
  PNR_LAUNCHED();
  return PNR_OK;
}

// ---- inverse-CDF bin lookup ---------------------------------------------------------------------
// searchsorted(cdf, u, right=True) - 1, clamped below at 0 only (nerf.py:138-139):
// number of cdf entries <= u, minus one.  n entries, ascending.
__device__ __forceinline__ float cdf_bin(const float* cdf, int n, float u) {
  int lo = 0, hi = n;  // first index with cdf[idx] > u
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  float ind = (float)lo - 1.f;
  return fmaxf(ind, 0.f);
}

__global__ void fine_indices_kernel(const float* __restrict__ cdf, const float* __restrict__ u, long long n, int Kc,
                                    int Kf, float* __restrict__ inds) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long r = i / Kf;
  inds[i] = cdf_bin(cdf + r * (Kc + 1), Kc + 1, u[i]);
}

int launch_fine_indices(const float* cdf, const float* u, int B, int Kc, int Kf, float* inds, cudaStream_t st) {
  long long n = (long long)B * Kf;
  if (n == 0) return PNR_OK;
  fine_indices_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(cdf, u, n, Kc, Kf, inds);
  PNR_LAUNCHED();
  return PNR_OK;
}

// ---- sample_fine + sample_fine_depth + cat + sort -------------------------------------------------
// One warp per ray.  smem per warp: cdf[Kc+1] followed by the sort buffer [Kpad] (Kpad = next pow2 >= Kc + n_fine).
//   * cdf: the divisions pdf = (w + 1e-5) / sum run on all lanes; the running sum itself stays ONE sequential fp32
//     chain (c_k = c_{k-1} + pdf_k, the order of torch.cumsum on a CPU row), executed by lane 0 on registers --
//     the bin search depends on the last bit of every cdf entry;
//   * sort: the coarse samples arrive ascending (stratified), so only the n_fine new samples are sorted (one per
//     lane, a 15-step shuffle bitonic network) and the two sorted lists are merged by rank: position = own index +
//     number of elements of the other list before it (ties: coarse first).  If a rounding inversion left the
//     coarse list unsorted, or n_fine > 32, the general shared-memory bitonic sort of all Kpad values runs instead.
__device__ __forceinline__ float warp_sort32(float v, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, stride);
      const bool up = (lane & size) == 0;          // ascending block
      const bool lower = (lane & stride) == 0;     // this lane keeps the smaller value of the pair when ascending
      const float mn = fminf(v, o), mx = fmaxf(v, o);
      v = (lower == up) ? mn : mx;
    }
  }
  return v;
}

__global__ void __launch_bounds__(256)
sample_fine_sorted_kernel(const float* __restrict__ rays, const float* __restrict__ z_coarse,
                          const float* __restrict__ weights, const float* __restrict__ depth,
                          const float* __restrict__ fine_u, const float* __restrict__ fine_jit,
                          const float* __restrict__ depth_nrm, int B, int Kc, int n_imp, int n_dep, float depth_std,
                          int lindisp, int Kpad, int Kbuf, float* __restrict__ z_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= B) return;  // whole warp exits together
  float* cdf = smem + (size_t)warp * (Kc + 1 + Kbuf);  // Kbuf = max(Kpad, Kc + 32) floats of sort / merge buffer
  float* buf = cdf + Kc + 1;
  const float near = rays[(size_t)r * 8 + 6], far = rays[(size_t)r * 8 + 7];
  const int n_fine = n_imp + n_dep, K = Kc + n_fine;
  const bool fast = n_fine <= 32;
  // fast path: this lane's random inputs are requested now, so that their DRAM latency overlaps the cdf work
  float in_a = 0.f, in_b = 0.f;
  if (fast && lane < n_fine) {
    if (lane < n_imp) {
      in_a = __ldg(fine_u + (size_t)r * n_imp + lane);
      in_b = __ldg(fine_jit + (size_t)r * n_imp + lane);
    } else {
      in_a = __ldg(depth_nrm + (size_t)r * n_dep + (lane - n_imp));
      in_b = __ldg(depth + r);
    }
  }
  // coarse samples go into the sort buffer; padding is +inf
  bool sorted = true;
  for (int k = lane; k < Kpad; k += 32) {
    const float zk = k < Kc ? z_coarse[(size_t)r * Kc + k] : CUDART_INF_F;
    buf[k] = zk;
  }
  if (n_imp > 0) {
    // pdf = (w + 1e-5) / sum ; cdf = [0, cumsum(pdf)]                          nerf.py:129-133
    float s = 0.f;
    for (int k = lane; k < Kc; k += 32) {
      float w = __fadd_rn(weights[(size_t)r * Kc + k], 1e-5f);
      cdf[k + 1] = w;
      s += w;
    }
    s = warp_sum(s);
    for (int k = lane; k < Kc; k += 32) cdf[k + 1] = __fdiv_rn(cdf[k + 1], s);
    __syncwarp();
    if (lane == 0) {  // sequential running sum, like torch.cumsum on one row
      float c = 0.f;
      cdf[0] = 0.f;
      int k = 1;
      for (; k + 7 <= Kc; k += 8) {  // loads in flight together, adds in order
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = cdf[k + i];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          c = __fadd_rn(c, p[i]);
          cdf[k + i] = c;
        }
      }
      for (; k <= Kc; ++k) {
        c = __fadd_rn(c, cdf[k]);
        cdf[k] = c;
      }
    }
    __syncwarp();
  }
  // the n_fine new samples: lane j (j, j+32, ...) computes sample j
  float mine = CUDART_INF_F;
  for (int j = lane; j < n_fine; j += 32) {
    float v;
    if (j < n_imp) {
      float ind = cdf_bin(cdf, Kc + 1, fast ? in_a : fine_u[(size_t)r * n_imp + j]);
      // z_steps = (inds + rand) / n_coarse                                     nerf.py:141
      float zs = __fdiv_rn(__fadd_rn(ind, fast ? in_b : fine_jit[(size_t)r * n_imp + j]), (float)Kc);
      v = lerp_depth(near, far, zs, lindisp);
    } else {
      // clamp(depth + randn*depth_std, near, far)                              nerf.py:157-160
      float zz = __fadd_rn(fast ? in_b : depth[r], __fmul_rn(fast ? in_a : depth_nrm[(size_t)r * n_dep + (j - n_imp)], depth_std));
      v = fmaxf(fminf(zz, far), near);
    }
    if (fast) mine = v; else buf[Kc + j] = v;
  }
  __syncwarp();
  for (int k = lane; k + 1 < Kc; k += 32) sorted = sorted && (buf[k] <= buf[k + 1]);
  sorted = __all_sync(0xffffffffu, sorted);
  float* zr = z_out + (size_t)r * K;
  if (fast && sorted) {
    // ---- merge by rank: coarse list (Kc, ascending, in buf[0..Kc)) with the sorted new samples (one per lane)
    const float f = warp_sort32(mine, lane);
    float* fs = buf + Kc;           // sorted new samples, +inf padded
    fs[lane] = f;
    __syncwarp();
    if (lane < n_fine) {            // position = lane + #(coarse <= f)
      int lo = 0, hi = Kc;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (buf[mid] <= f) lo = mid + 1; else hi = mid;
      }
      zr[lane + lo] = f;
    }
    for (int k = lane; k < Kc; k += 32) {  // position = k + #(new < z_k)
      const float zk = buf[k];
      int lo = 0, hi = n_fine;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (fs[mid] < zk) lo = mid + 1; else hi = mid;
      }
      zr[k + lo] = zk;
    }
    return;
  }
  if (fast) buf[Kc + lane] = mine;  // (lanes >= n_fine hold +inf: equals the padding)
  __syncwarp();
  // general path: bitonic sort of Kpad values in shared memory, ascending
  for (int size = 2; size <= Kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (Kpad >> 1); t += 32) {
        int i = 2 * t - (t & (stride - 1));  // index with bit 'stride' clear
        int j = i + stride;
        bool up = ((i & size) == 0);
        float a = buf[i], b = buf[j];
        if ((a > b) == up) {
          buf[i] = b;
          buf[j] = a;
        }
      }
      __syncwarp();
    }
  }
  for (int k = lane; k < K; k += 32) zr[k] = buf[k];
}

int launch_sample_fine_sorted(const float* rays, const float* z_coarse, const float* weights, const float* depth,
                              const float* fine_u, const float* fine_jitter, const float* depth_normal, int B,
                              int Kc, int n_fine, int n_fine_depth, float depth_std, int lindisp, float* z_out,
                              cudaStream_t st) {
  if (B == 0) return PNR_OK;
  int n_dep = n_fine_depth, n_imp = n_fine - n_fine_depth;
  PNR_CHECK_ARG(n_imp >= 0 && n_dep >= 0, "sample_fine: n_fine (%d) must include n_fine_depth (%d)", n_fine, n_dep);
  int K = Kc + n_fine;
  int Kpad = 2;
  while (Kpad < K) Kpad <<= 1;
  PNR_UNSUPPORTED(Kpad > 2048, "sample_fine: more than 2048 samples per ray");
  const int Kbuf = Kpad > Kc + 32 ? Kpad : Kc + 32;
  int wpb = 8;
  size_t smem = (size_t)wpb * (Kc + 1 + Kbuf) * sizeof(float);
  while (smem > 48 * 1024 && wpb > 1) {
    wpb >>= 1;
    smem = (size_t)wpb * (Kc + 1 + Kbuf) * sizeof(float);
  }
  PNR_UNSUPPORTED(smem > 48 * 1024, "sample_fine: per-ray sample count too large for shared memory");
  sample_fine_sorted_kernel<<<ceil_div(B, wpb), wpb * 32, smem, st>>>(rays, z_coarse, weights, depth, fine_u,
                                                                     fine_jitter, depth_normal, B, Kc, n_imp, n_dep,
                                                                     depth_std, lindisp, Kpad, Kbuf, z_out);
  PNR_LAUNCHED();
  return PNR_OK;
}

}  // namespace pnr
