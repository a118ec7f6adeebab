// fp32 VALIDATION path of kernel (b): ResnetFC (src/model/resnetfc.py:173-236) with true fp32
// multiplicands and fp32 accumulation on the SIMT FMA pipe (no TF32, no tensor cores), used to
// check the bf16/tcgen05 production path and to meet the <=1e-4 fp32 parity mode.
//
//   x = lin_in(code)                                   resnetfc.py:199
//   for b < n_blocks:
//     if b == combine_layer: x = mean over NS views     resnetfc.py:204-224, util.py:466-476
//     if b <  combine_layer: x += lin_z[b](z)           resnetfc.py:226-232
//     x = x + fc_1(relu(fc_0(relu(x))))                 resnetfc.py:53-62,234
//   out = lin_out(relu(x))                              resnetfc.py:235
#include "common.cuh"

namespace pnr {

// C[M,N] (ldc) = (ACC ? C : 0) + act(A[M,K] (lda)) @ W[N,K]^T + bias[N]
// 128x128x8 block tile, 256 threads, 8x8 outputs per thread (two 4-wide strips per dimension).
template <bool RELU_IN, bool ACC>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ C, int ldc, long long M, int N, int K) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // loader mapping: thread -> (row = tid/2, k-quad = tid%2)
  const int lr = tid >> 1, lk = (tid & 1) * 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const long long arow = m0 + lr;
  const int wrow = n0 + lr;
  const float* Ap = A + arow * (long long)lda;
  const float* Wp = W + (long long)wrow * K;
  float ra[4], rw[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = k0 + lk + i;
      float a = (arow < M && k < K) ? Ap[k] : 0.f;
      if (RELU_IN) a = fmaxf(a, 0.f);
      ra[i] = a;
      rw[i] = (wrow < N && k < K) ? Wp[k] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = ra[i];
      Ws[lk + i][lr] = rw[i];
    }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      *(float4*)&a[0] = *(const float4*)&As[k][ty * 4];
      *(float4*)&a[4] = *(const float4*)&As[k][64 + ty * 4];
      *(float4*)&b[0] = *(const float4*)&Ws[k][tx * 4];
      *(float4*)&b[4] = *(const float4*)&Ws[k][64 + tx * 4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int n = n0 + jh * 64 + tx * 4;
      float* cp = C + m * (long long)ldc + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < N) {
          float v = acc[i][jh * 4 + j] + bias[n + j];
          if (ACC) v += cp[j];
          cp[j] = v;
        }
      }
    }
  }
}

template <bool RELU_IN, bool ACC>
static int launch_linear(const float* A, int lda, const float* W, const float* bias, float* C, int ldc,
                         long long M, int N, int K, cudaStream_t st) {
  if (M == 0) return PNR_OK;
  dim3 grid((unsigned)ceil_div_ll(M, 128), (unsigned)ceil_div(N, 128));
  linear_f32_kernel<RELU_IN, ACC><<<grid, 256, 0, st>>>(A, lda, W, bias, C, ldc, M, N, K);
  PNR_LAUNCHED();
  return PNR_OK;
}

// x_pooled[(sb*P + p)] = mean_v x[((sb*NS + v)*P + p)]      (util.combine_interleaved, average)
__global__ void view_mean_kernel(const float* __restrict__ x, float* __restrict__ y, int SB, int NS, long long P,
                                 int D) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over SB*P*D/4 float4
  long long total = (long long)SB * P * (D / 4);
  if (i >= total) return;
  int d4 = (int)(i % (D / 4));
  long long sp = i / (D / 4);
  int sb = (int)(sp / P);
  long long p = sp - (long long)sb * P;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = 0; v < NS; ++v) {
    float4 t = *(const float4*)(x + (((long long)sb * NS + v) * P + p) * D + d4 * 4);
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  float inv = (float)NS;
  s.x /= inv; s.y /= inv; s.z /= inv; s.w /= inv;
  *(float4*)(y + sp * D + d4 * 4) = s;
}

// out[m][0..d_out) = lin_out(relu(x[m])); optional head: sigmoid on 0..2, relu on 3
// (models.py.backup2:274-281).  One warp per row.
__global__ void lin_out_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                               float* __restrict__ out, long long M, int D, int d_out, int apply_head) {
  long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (m >= M) return;
  const float* xr = x + m * D;
  for (int o = 0; o < d_out; ++o) {
    float s = 0.f;
    for (int k = lane; k < D; k += 32) s = fmaf(fmaxf(xr[k], 0.f), W[o * D + k], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
      s += b[o];
      if (apply_head) s = (o < 3) ? 1.f / (1.f + expf(-s)) : fmaxf(s, 0.f);
      out[m * d_out + o] = s;
    }
  }
}

size_t mlp_f32_workspace(const pnr_mlp& m, long long rows_pre, long long rows_post) {
  Arena a(nullptr, 0);
  a.take<float>((size_t)rows_pre * m.d_hidden);  // x
  a.take<float>((size_t)rows_pre * m.d_hidden);  // net
  if (rows_post != rows_pre) a.take<float>((size_t)rows_post * m.d_hidden);
  return a.off + 256;
}

int mlp_forward_f32(const pnr_mlp& m, const float* zx, int SB, int NS, int P, float* out, bool apply_head,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  PNR_UNSUPPORTED(m.combine_type != 0, "ResnetFC combine_type other than 'average' is not supported natively");
  PNR_CHECK_ARG(m.d_hidden % 4 == 0, "d_hidden must be a multiple of 4");
  const int D = m.d_hidden, L = m.d_latent, ld = m.d_latent + m.d_in;
  const bool pools = (m.combine_layer < m.n_blocks) && NS > 1;
  long long rows_pre = (long long)SB * NS * P, rows_post = pools ? (long long)SB * P : rows_pre;
  Arena a(ws, ws_bytes);
  float* x = a.take<float>((size_t)rows_pre * D);
  float* net = a.take<float>((size_t)rows_pre * D);
  float* xp = pools ? a.take<float>((size_t)rows_post * D) : nullptr;
  if (!a.ok()) {
    set_err("mlp_forward_f32: workspace too small (%zu < %zu)", ws_bytes, a.off);
    return PNR_ERR_WORKSPACE;
  }
  long long M = rows_pre;
  PNR_CHECK_ARG(m.d_in > 0, "d_in == 0 is not supported");
  PNR_TRY((launch_linear<false, false>(zx + L, ld, m.lin_in_w, m.lin_in_b, x, D, M, D, m.d_in, st)));
  for (int b = 0; b < m.n_blocks; ++b) {
    if (b == m.combine_layer && pools) {
      long long total = (long long)SB * P * (D / 4);
      view_mean_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(x, xp, SB, NS, P, D);
      PNR_LAUNCHED();
      x = xp;
      M = rows_post;
    }
    if (L > 0 && b < m.combine_layer && b < m.n_lin_z)
      PNR_TRY((launch_linear<false, true>(zx, ld, m.lin_z_w[b], m.lin_z_b[b], x, D, M, D, L, st)));
    PNR_TRY((launch_linear<true, false>(x, D, m.fc0_w[b], m.fc0_b[b], net, D, M, D, D, st)));
    PNR_TRY((launch_linear<true, true>(net, D, m.fc1_w[b], m.fc1_b[b], x, D, M, D, D, st)));
  }
  if (M > 0) {
    lin_out_kernel<<<(unsigned)ceil_div_ll(M, 8), 256, 0, st>>>(x, m.lin_out_w, m.lin_out_b, out, M, D, m.d_out,
                                                                apply_head ? 1 : 0);
    PNR_LAUNCHED();
  }
  return PNR_OK;
}

}  // namespace pnr
