// Kernels (a)+(b), production path: fused point-feature gather + ResnetFC
// (src/model/models.py.backup2:155-282, src/model/resnetfc.py:173-236) as a chain of tcgen05
// (UTCHMMA.2CTA) GEMMs with bf16 operands and fp32 accumulation in tensor memory -- ONE persistent
// kernel per network evaluation (mlp_fused_kernel).
//
// Design ("pair-64"): a cluster of 2 CTAs owns a tile of 128 rows, 64 rows per CTA, and issues
// cta_group::2 MMAs with M=128, N=256, K=16.  Per CTA:
//   TMEM  512 columns = X (64 rows x 512 fp32 residual stream, accumulated IN PLACE by the lin_z and
//         fc_1 GEMMs, so the residual add costs nothing and never leaves fp32) + NET (fc_0 output)
//   smem  S_x = relu(x) bf16 (64 KB) and H = relu(net) bf16 (64 KB) as K-major SWIZZLE_NONE UMMA
//         operands + one ring of 6 x 16 KB slots carrying weight chunks and [latent|code] operand
//         slices, filled by 2-D tiled TMA (UTMALDG.2D.2CTA) whose bytes complete on the leader CTA's
//         mbarrier.  Weights are pre-packed into the exact shared-memory image, so a box of
//         rows-of-256-bytes is a plain linear copy.
//   warps 0 TMA producer | 1 MMA issuer (leader CTA only) | 2-3 gather (operand image of the NEXT tile;
//         warp 2 also owns the TMEM allocation) | 4-11 epilogue (TMEM -> +bias, relu -> bf16 operand
//         panels; view mean-pool; lin_out head).  Producer and issuer run warp-uniform, electing one
//         lane only around the instruction, so descriptors live in uniform registers.
// A pair walks its tiles in GROUPS: nA = 64 / (64 / NS) "A tiles" (rows = points x views: gather,
//         lin_in+lin_z[0], blocks 0..combine_layer-1 with the next block's lin_z GEMM scheduled between
//         fc_0 and fc_1, view mean-pool = util.combine_interleaved), whose pooled fp32 rows (64/NS points
//         per CTA and tile) are parked in a 128 KB per-CTA staging buffer that is rewritten every group
//         and therefore stays in L2, followed by one "B tile" (rows = the group's pooled points, 64 per
//         CTA): remaining blocks, lin_out on CUDA cores from TMEM, sigmoid/relu head
//         (models.py.backup2:274-281).  Nothing but rays/samples in and (rgb, sigma) out touches DRAM.
// Every mbarrier wait is wall-clock bounded; a protocol fault writes its tag to pinned host memory
// and traps (the next call on that device reports it) instead of hanging the GPU.
#include <cstdlib>
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "features.cuh"
#include "tc_ptx.cuh"

namespace pnr {
using namespace ptx;

static thread_local unsigned long long* g_stats_ptr = nullptr;  // debug cycle counters (host pointer holder)
void tc_set_stats(unsigned long long* p) { g_stats_ptr = p; }

#ifndef PNR_TC_STATS
#define PNR_TC_STATS 0  // 1: per-role cycle counters (tools/tc_stats.py); costs ~10% of the MMA issue rate
#endif

namespace tc {
constexpr int DH = 512;                 // hidden width the tensor-core path is specialised for
constexpr int ROWS = 64;                // rows per CTA
constexpr int KS = 64;                  // K elements per slice
constexpr int B_CHUNK = 128 * KS * 2;   // 16 KB: [8 k-groups][128 n-rows][8 bf16]
constexpr int A_SLICE = ROWS * KS * 2;  // 8 KB : [8 k-groups][64 rows][8 bf16]
constexpr int NB_ST = 6;                // unified operand ring: weight chunks (16 KB) and [latent|code] slices (8 KB)
#ifndef PNR_RING_A
#define PNR_RING_A NB_ST                // experiment knob: ring slots actually used (<= NB_ST); 4/5/6 -> 632k/654k/665k rays/s on C2
#endif
#ifndef PNR_B_SPLIT
#define PNR_B_SPLIT 2
#endif
#ifndef PNR_A_SPLIT
#define PNR_A_SPLIT 1
#endif
constexpr int B_SPLIT = PNR_B_SPLIT;    // TMA boxes per weight chunk (8/4/2/1 -> 667/671/673/673 k rays/s on C2)
constexpr int A_SPLIT = PNR_A_SPLIT;    // TMA boxes per operand slice
constexpr int OFF_SX = 0;
constexpr int OFF_H = OFF_SX + ROWS * DH * 2;
constexpr int OFF_BRING = OFF_H + ROWS * DH * 2;
constexpr int OFF_BARS = OFF_BRING + NB_ST * B_CHUNK;
constexpr int SMEM_BYTES = OFF_BARS + 512;
constexpr int THREADS = 384;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

// barrier indices (uint64 each)
constexpr int B_FULL = 0, B_EMPTY = B_FULL + NB_ST,
              X_READY = B_EMPTY + NB_ST, NET_READY = X_READY + 2, SX_READY = NET_READY + 2, H_READY = SX_READY + 8,
              XP_DONE = H_READY + 8, ZC_READY = XP_DONE + 1, ZC_TAKEN = ZC_READY + 1, N_BARS = ZC_TAKEN + 1;

struct Params {
  const uint8_t* w;                    // packed image base
  uint32_t off_g1[PNR_MAX_BLOCKS];     // lin_in+lin_z[0] (b=0), lin_z[b] (b>0)
  uint32_t off_g2[PNR_MAX_BLOCKS];     // fc_0
  uint32_t off_g3[PNR_MAX_BLOCKS];     // fc_1
  uint32_t off_biasA, off_biasB, off_bias0, off_lin_out;  // fp32 tables
  int n_pre, n_post;                   // blocks before / after the view pool
  int nks_z, nks_c;                    // K slices of the latent / code part of an input row
  int ns, ppw;                         // views per point, points per CTA (= 64 / ns; rows pl*ns + v)
  int nA;                              // A tiles per group (= 64 / ppw): their pooled points fill one B tile
  long long P;                         // points
  int tilesA;
  const uint8_t* zc;                   // [tileA][cta][slice][A_SLICE]
  float* stage;                        // pooled residual rows of the current group: [CTA][nb][h][col/4][row (64)][4] fp32
  float* out;                          // (P,4)
  int apply_head;
  int* err;
  unsigned long long* stats;           // optional [pairs][16] cycle counters (debug)
  // in-kernel gather (warps 2-3 produce the operand image of the next A tile one tile ahead of the MMAs)
  int fused_gather;                    // 0: zc was written by rows_to_operand_kernel (pnr_mlp_forward)
  int zc_ring;                         // > 0 (fused gather): zc holds zc_ring (2) tiles per cluster pair, reused round-robin,
                                       // so the image stays in L2 and is never written back to DRAM; 0: one slot per tile
  const float *xyz, *viewdirs, *rays, *zsamp;
  int K;
  pnr_scene sc;
  CUtensorMap tm_w;                    // packed weights as rows of 256 B, box = 16 rows (4 KB)
  CUtensorMap tm_zc;                   // operand image as rows of 256 B, box = 32 rows (8 KB)
};
}  // namespace tc
using namespace tc;

// ---------------------------------------------------------------------------------------------
// row <-> point mapping of an A tile: CTA c, row r (0..63): point-in-CTA pl = r / ns, view
// v = r % ns; rows with pl >= ppc (= 64 / ns) are padding.  A point's rows are adjacent TMEM lanes; for
// ns that do not divide 32 one point straddles the two 32-lane groups (handled in the pool epilogue).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long tileA_point(int tile, int cta, int row, int ns, int ppc, int& v, bool& valid) {
  int pl = row / ns;
  v = row - pl * ns;
  valid = pl < ppc;
  return (long long)(tile * 2 + cta) * ppc + pl;
}

// =============================================================================================
// weight packing
// =============================================================================================
struct PackJob {
  const float* w;   // (512, K) row-major (nn.Linear weight)
  int K;            // valid K columns of this source
  int k_slices;     // 64-wide slices this source occupies
  uint32_t dst_off; // byte offset of its first chunk
  int fmt;          // 0 = bf16, 1 = f16
};

__global__ void pack_weights_kernel(PackJob job, uint8_t* __restrict__ dst) {
  // one thread per 16-byte group: index = ((((s*2+nb)*2+c)*8+g)*128+i)
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)job.k_slices * 4 * 8 * 128;
  if (idx >= total) return;
  int i = (int)(idx & 127);
  int g = (int)((idx >> 7) & 7);
  int c = (int)((idx >> 10) & 1);
  int nb = (int)((idx >> 11) & 1);
  int s = (int)(idx >> 12);
  int n = nb * 256 + c * 128 + i;
  int k0 = s * 64 + g * 8;
  uint32_t v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float lo = (k0 + 2 * e < job.K) ? job.w[(size_t)n * job.K + k0 + 2 * e] : 0.f;
    float hi = (k0 + 2 * e + 1 < job.K) ? job.w[(size_t)n * job.K + k0 + 2 * e + 1] : 0.f;
    v[e] = job.fmt == 1 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
  }
  *reinterpret_cast<uint4*>(dst + job.dst_off + idx * 16) = make_uint4(v[0], v[1], v[2], v[3]);
}

__global__ void pack_bias_kernel(const float* a, const float* b, const float* c, const float* prev, float* dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = prev ? prev[i] : 0.f;
  if (a) v += a[i];
  if (b) v += b[i];
  if (c) v += c[i];
  dst[i] = v;
}

struct Layout {
  uint32_t off_g1[PNR_MAX_BLOCKS], off_g2[PNR_MAX_BLOCKS], off_g3[PNR_MAX_BLOCKS];
  uint32_t off_biasA, off_biasB, off_bias0, off_lin_out;
  int n_pre, n_post, nks_z, nks_c;
  size_t total;
};

static int tc_supported(const pnr_mlp& m) {
  PNR_UNSUPPORTED(m.d_hidden != DH, "tensor-core MLP is specialised for d_hidden=512 (got %d); use precision fp32", m.d_hidden);
  PNR_UNSUPPORTED(m.d_latent <= 0 || m.d_latent % 64 != 0, "tensor-core MLP needs d_latent %% 64 == 0 (got %d)", m.d_latent);
  PNR_UNSUPPORTED(m.d_in <= 0 || m.d_in > 128, "tensor-core MLP needs 0 < d_in <= 128 (got %d)", m.d_in);
  PNR_UNSUPPORTED(m.d_out != 4, "tensor-core MLP needs d_out == 4");
  PNR_UNSUPPORTED(m.n_lin_z != (m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks),
                  "unexpected number of lin_z layers");
  PNR_UNSUPPORTED(m.n_lin_z < 1, "tensor-core MLP needs at least one latent injection (combine_layer >= 1)");
  return PNR_OK;
}

static Layout make_layout(const pnr_mlp& m) {
  Layout L;
  memset(&L, 0, sizeof(L));
  L.n_pre = m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks;
  L.n_post = m.n_blocks - L.n_pre;
  L.nks_z = m.d_latent / 64;
  L.nks_c = (m.d_in + 63) / 64;
  size_t off = 0;
  auto group = [&](int slices) {
    uint32_t o = (uint32_t)off;
    off += (size_t)slices * 4 * B_CHUNK;
    return o;
  };
  for (int b = 0; b < m.n_blocks; ++b) {
    if (b < L.n_pre) L.off_g1[b] = group(b == 0 ? L.nks_z + L.nks_c : L.nks_z);
    L.off_g2[b] = group(DH / 64);
    L.off_g3[b] = group(DH / 64);
  }
  L.off_biasA = (uint32_t)off;  off += (size_t)(L.n_pre + 1) * DH * 4;
  L.off_biasB = (uint32_t)off;  off += (size_t)(L.n_post + 1) * DH * 4;
  L.off_bias0 = (uint32_t)off;  off += (size_t)m.n_blocks * DH * 4;
  L.off_lin_out = (uint32_t)off; off += (size_t)(4 * DH + 4) * 4;
  L.total = align_up(off, 256);
  return L;
}

size_t mlp_tc_packed_bytes(const pnr_mlp& m) {
  if (tc_supported(m) != PNR_OK) return 256;
  return make_layout(m).total;
}

int mlp_tc_pack(const pnr_mlp& m, void* dst, size_t dst_bytes, int fmt, cudaStream_t st) {
  PNR_TRY(tc_supported(m));
  Layout L = make_layout(m);
  PNR_CHECK_ARG(dst_bytes >= L.total, "mlp_pack: destination too small (%zu < %zu)", dst_bytes, L.total);
  PNR_CHECK_ARG(((uintptr_t)dst & 15) == 0, "mlp_pack: destination must be 16-byte aligned");
  uint8_t* d = (uint8_t*)dst;
  auto pack = [&](const float* w, int K, int slices, uint32_t off) -> int {
    PackJob j{w, K, slices, off, fmt};
    long long total = (long long)slices * 4 * 8 * 128;
    pack_weights_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(j, d);
    PNR_LAUNCHED();
    return PNR_OK;
  };
  for (int b = 0; b < m.n_blocks; ++b) {
    if (b < L.n_pre) {
      PNR_TRY(pack(m.lin_z_w[b], m.d_latent, L.nks_z, L.off_g1[b]));
      if (b == 0) PNR_TRY(pack(m.lin_in_w, m.d_in, L.nks_c, L.off_g1[0] + (uint32_t)L.nks_z * 4 * B_CHUNK));
    }
    PNR_TRY(pack(m.fc0_w[b], DH, DH / 64, L.off_g2[b]));
    PNR_TRY(pack(m.fc1_w[b], DH, DH / 64, L.off_g3[b]));
  }
  // bias tables.  biasA[b]: constant to add to the TMEM residual when it is read before block b of
  // the A tiles (the GEMMs accumulate without biases); biasA[n_pre] is the pooled output's constant.
  float* biasA = (float*)(d + L.off_biasA);
  float* biasB = (float*)(d + L.off_biasB);
  float* bias0 = (float*)(d + L.off_bias0);
  auto bias = [&](const float* a, const float* b, const float* c, const float* prev, float* out) -> int {
    pack_bias_kernel<<<2, 256, 0, st>>>(a, b, c, prev, out, DH);
    PNR_LAUNCHED();
    return PNR_OK;
  };
  PNR_TRY(bias(m.lin_in_b, m.lin_z_b[0], nullptr, nullptr, biasA));
  for (int b = 1; b <= L.n_pre; ++b)
    PNR_TRY(bias(m.fc1_b[b - 1], b < L.n_pre ? m.lin_z_b[b] : nullptr, nullptr, biasA + (size_t)(b - 1) * DH,
                 biasA + (size_t)b * DH));
  PNR_TRY(bias(nullptr, nullptr, nullptr, nullptr, biasB));  // pooled x is stored exactly
  for (int j = 1; j <= L.n_post; ++j)
    PNR_TRY(bias(m.fc1_b[L.n_pre + j - 1], nullptr, nullptr, biasB + (size_t)(j - 1) * DH, biasB + (size_t)j * DH));
  for (int b = 0; b < m.n_blocks; ++b) PNR_TRY(bias(m.fc0_b[b], nullptr, nullptr, nullptr, bias0 + (size_t)b * DH));
  PNR_CUDA(cudaMemcpyAsync(d + L.off_lin_out, m.lin_out_w, (size_t)4 * DH * 4, cudaMemcpyDeviceToDevice, st));
  PNR_CUDA(cudaMemcpyAsync(d + L.off_lin_out + (size_t)4 * DH * 4, m.lin_out_b, 16, cudaMemcpyDeviceToDevice, st));
  return PNR_OK;
}

// =============================================================================================
// kernel (a), bf16 operand variant (gather warps of the fused kernel): [latent | code] straight into the
// A-tile operand image  zc[slot][cta][k-group][row][8 bf16]
// =============================================================================================
template <int FMT>
__device__ __forceinline__ void x8_to_float(const uint4& u, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t;
    if constexpr (FMT == 1) t = __half22float2(reinterpret_cast<const __half2*>(&u)[i]);
    else t = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(&u)[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// select one of three registers without dynamic indexing (which would go through local memory)
__device__ __forceinline__ float sel3(float a, float b, float c, int i) { return i == 0 ? a : (i == 1 ? b : c); }

// Production variant of the positional code: same layout as code_entry() (features.cuh), hardware
// sine (the rounding to bf16 hides its ~1e-4 error at |arg| <= 300) and constant-divisor index math.
struct CodeCtx {
  int dz, db, coded, d_in, use_code, include_input, use_xyz, normalize_z;
  float freq_factor;
};
__device__ __forceinline__ float code_entry_fast(const CodeCtx& c, const PointCam& pc, int j) {
  if (j >= c.d_in) return 0.f;
  auto base = [&](int i) -> float {
    if (i < c.dz) {
      if (c.use_xyz) return c.normalize_z ? sel3(pc.xr[0], pc.xr[1], pc.xr[2], i) : sel3(pc.xc[0], pc.xc[1], pc.xc[2], i);
      return c.normalize_z ? -pc.xr[2] : -pc.xc[2];
    }
    return sel3(pc.vd[0], pc.vd[1], pc.vd[2], i - c.dz);
  };
  if (j >= c.coded) return sel3(pc.vd[0], pc.vd[1], pc.vd[2], j - c.coded);
  if (!c.use_code) return base(j);
  if (c.include_input) {
    if (j < c.db) return base(j);
    j -= c.db;
  }
  int g;
  switch (c.db) {  // constant divisors -> multiply-shift
    case 1: g = j; break;
    case 3: g = j / 3; break;
    case 4: g = j / 4; break;
    case 6: g = j / 6; break;
    default: g = j / c.db; break;
  }
  const int i = j - g * c.db;
  const float freq = c.freq_factor * (float)(1 << (g >> 1));
  const float phase = (g & 1) ? 1.57079637050628662109375f : 0.f;
  return __sinf(fmaf(base(i), freq, phase));
}

// bilinear taps without the normalise/unnormalise round trip (identical up to ~1e-5 texel)
__device__ __forceinline__ Taps make_taps_fast(float u, float v, int H, int W, float kx, float ky) {
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float ix = fminf(fmaxf(u * kx, 0.f), wm1), iy = fminf(fmaxf(v * ky, 0.f), hm1);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float fx1 = ix - x0f, fy1 = iy - y0f, fx0 = 1.f - fx1, fy0 = 1.f - fy1;
  const bool xin = x0 + 1 <= W - 1, yin = y0 + 1 <= H - 1;
  const int x1 = xin ? x0 + 1 : x0, y1 = yin ? y0 + 1 : y0;
  Taps t;
  t.o00 = y0 * W + x0;
  t.o01 = y0 * W + x1;
  t.o10 = y1 * W + x0;
  t.o11 = y1 * W + x1;
  t.w00 = fx0 * fy0;
  t.w01 = xin ? fx1 * fy0 : 0.f;
  t.w10 = yin ? fx0 * fy1 : 0.f;
  t.w11 = (xin && yin) ? fx1 * fy1 : 0.f;
  return t;
}

// Gather warps of the fused kernel: lane = one row of the tile; the lane loops
// over the 8-channel groups (4 scattered 128-bit tap loads each; adjacent groups share 32-B sectors
// through L1) and over the code groups, and writes 16-byte operand entries -- a warp stores 512
// contiguous bytes per group.
template <int FMT>
__device__ __forceinline__ void gather_row_to_zc(const pnr_scene& sc, const float* __restrict__ xyz,
                                                 const float* __restrict__ viewdirs, const float* __restrict__ rays,
                                                 const float* __restrict__ z, int K, long long gp, int v, bool valid,
                                                 int nks_z, int nks_c, uint8_t* __restrict__ base /* + kgroup*1024 */) {
  const int nsl = nks_z + nks_c;
  if (!valid) {
    for (int g = 0; g < nsl * 8; ++g) *reinterpret_cast<uint4*>(base + (size_t)g * 1024) = make_uint4(0, 0, 0, 0);
    return;
  }
  float X[3], D[3];
  if (rays != nullptr) {
    const unsigned gpu = (unsigned)gp, r = gpu / (unsigned)K;
    const float4 ra = __ldg(reinterpret_cast<const float4*>(rays) + 2 * (size_t)r);
    const float4 rb = __ldg(reinterpret_cast<const float4*>(rays) + 2 * (size_t)r + 1);
    const float t = __ldg(z + gpu);
    X[0] = ra.x + t * ra.w;
    X[1] = ra.y + t * rb.x;
    X[2] = ra.z + t * rb.y;
    D[0] = ra.w;
    D[1] = rb.x;
    D[2] = rb.y;
  } else {
    load_point(xyz, viewdirs, nullptr, nullptr, 0, gp, X, D);
  }
  PointCam pc;
  camera_project(sc.cams + v * 16, X, D, pc);
  for (int l = 0; l < sc.n_levels; ++l) {
    const int C8 = sc.C[l] >> 3, H = sc.H[l], W = sc.W[l];
    const Taps t = make_taps_fast(pc.u, pc.v, H, W, sc.kx[l], sc.ky[l]);
    const uint4* f = reinterpret_cast<const uint4*>(sc.level[l]) + (size_t)v * H * W * C8;
    const uint4 *p00 = f + (size_t)t.o00 * C8, *p01 = f + (size_t)t.o01 * C8, *p10 = f + (size_t)t.o10 * C8,
                *p11 = f + (size_t)t.o11 * C8;
    uint8_t* dst = base + (size_t)(sc.ch_off[l] >> 3) * 1024;
#pragma unroll 4
    for (int g = 0; g < C8; ++g) {
      const uint4 q00 = __ldg(p00 + g), q01 = __ldg(p01 + g), q10 = __ldg(p10 + g), q11 = __ldg(p11 + g);
      float a[8], b[8], c[8], d[8];
      x8_to_float<FMT>(q00, a);
      x8_to_float<FMT>(q01, b);
      x8_to_float<FMT>(q10, c);
      x8_to_float<FMT>(q11, d);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = a[2 * e] * t.w00 + b[2 * e] * t.w01 + c[2 * e] * t.w10 + d[2 * e] * t.w11;
        float hi = a[2 * e + 1] * t.w00 + b[2 * e + 1] * t.w01 + c[2 * e + 1] * t.w10 + d[2 * e + 1] * t.w11;
        o[e] = pack2<FMT>(lo, hi);
      }
      *reinterpret_cast<uint4*>(dst + (size_t)g * 1024) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  CodeCtx cc;
  cc.dz = sc.use_xyz ? 3 : 1;
  cc.db = cc.dz + ((sc.use_viewdirs && sc.use_code && sc.use_code_viewdirs) ? 3 : 0);
  cc.coded = sc.use_code ? (sc.num_freqs * 2 * cc.db + (sc.include_input ? cc.db : 0)) : cc.db;
  cc.d_in = sc.d_in;
  cc.use_code = sc.use_code;
  cc.include_input = sc.include_input;
  cc.use_xyz = sc.use_xyz;
  cc.normalize_z = sc.normalize_z;
  cc.freq_factor = sc.freq_factor;
  for (int g = 0; g < nks_c * 8; ++g) {
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = pack2<FMT>(code_entry_fast(cc, pc, g * 8 + 2 * e), code_entry_fast(cc, pc, g * 8 + 2 * e + 1));
    *reinterpret_cast<uint4*>(base + (size_t)(nks_z * 8 + g) * 1024) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// fp32 rows in reference order (sb, ns, p) -> operand image (used by pnr_mlp_forward in bf16 mode)
__global__ void __launch_bounds__(256)
rows_to_operand_kernel(const float* __restrict__ zx, int d_latent, int d_in, int SB, int NS, long long Pper, int ppw,
                       int tilesA, int nks_z, int nks_c, int fmt, uint8_t* __restrict__ zc) {
  const int lane = threadIdx.x & 31;
  const long long wrow = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wrow >= (long long)tilesA * 128) return;
  const int tile = (int)(wrow >> 7), cta = (int)((wrow >> 6) & 1), row = (int)(wrow & 63);
  int v;
  bool valid;
  long long gp = tileA_point(tile, cta, row, NS, ppw, v, valid);
  valid = valid && gp < (long long)SB * Pper;
  const int nsl = nks_z + nks_c;
  uint8_t* base = zc + ((size_t)(tile * 2 + cta) * nsl) * A_SLICE + (size_t)row * 16;
  const int width = d_latent + d_in;
  const float* src = nullptr;
  if (valid) {
    long long sb = gp / Pper, p = gp - sb * Pper;
    src = zx + ((sb * NS + v) * Pper + p) * width;
  }
  for (int g = lane; g < nsl * 8; g += 32) {
    uint32_t o[4] = {0, 0, 0, 0};
    if (valid) {
      bool is_code = g >= nks_z * 8;
      int k0 = is_code ? (g - nks_z * 8) * 8 : g * 8;
      int lim = is_code ? d_in : d_latent;
      const float* s = src + (is_code ? d_latent : 0);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = (k0 + 2 * e < lim) ? s[k0 + 2 * e] : 0.f;
        float hi = (k0 + 2 * e + 1 < lim) ? s[k0 + 2 * e + 1] : 0.f;
        o[e] = fmt == 1 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
      }
    }
    *reinterpret_cast<uint4*>(base + (size_t)g * 1024) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// =============================================================================================
// the fused MLP kernels
// =============================================================================================
// Producer and MMA-issuer roles: run by ONE elected thread for the whole role (PNR_SOLO_ROLES=1) instead of
// the whole warp electing a lane around every instruction: the per-chunk control path (BRA.DIV / BSSY /
// ELECT / R2UR) was the bound of both kernels -- with a quarter of the tensor work AND a quarter of the
// weight traffic a tile still took 96 % of its cycles.
#ifndef PNR_SOLO_PROD
#define PNR_SOLO_PROD 0  // the producer measured 0.3-0.9 % slower solo (A/B on C2 and C4)
#endif
#if PNR_SOLO_PROD
#define PROD_ELECT() true
#define PROD_SYNC() ((void)0)
#define PROD_ENTER() elect_one()
#else
#define PROD_ELECT() elect_one()
#define PROD_SYNC() __syncwarp()
#define PROD_ENTER() true
#endif
// The issuer's mode is a template parameter (SM) of its helpers and of the fused kernel: solo is 1.5 % faster
// where the fc GEMMs dominate (L = 256: c3 458 vs 451 k rays/s) and 0.8 % slower on the multi-scale L = 512 schema
// (c4 384 vs 387 k), measured A/B on one box with the fused kernel; run_fused picks by latent width.
template <bool SM> __device__ __forceinline__ bool mma_elect() { if constexpr (SM) return true; else return elect_one(); }
template <bool SM> __device__ __forceinline__ void mma_sync() { if constexpr (!SM) __syncwarp(); }
template <bool SM> __device__ __forceinline__ bool mma_enter() { if constexpr (SM) return elect_one(); else return true; }

struct Ring {
  uint32_t full, empty;  // smem addresses of barrier arrays
  int n, idx;
  uint32_t phase;
  __device__ __forceinline__ void init(uint32_t f, uint32_t e, int n_) { full = f; empty = e; n = n_; idx = 0; phase = 0; }
  __device__ __forceinline__ void advance() { if (++idx == n) { idx = 0; phase ^= 1; } }
  __device__ __forceinline__ uint32_t full_bar() const { return full + idx * 8; }
  __device__ __forceinline__ uint32_t empty_bar() const { return empty + idx * 8; }
};

struct Ctx {
  uint32_t smem;   // shared::cta base address
  uint32_t bars;   // address of barrier 0
  uint32_t tmem;
  uint32_t rank;
  int* err;
  uint32_t idesc;  // tcgen05 instruction descriptor (operand format, M=128, N=256)
  long long w[6];  // cycles spent waiting, by class (debug statistics)
  __device__ __forceinline__ uint32_t bar(int i) const { return bars + i * 8; }
};

__device__ __forceinline__ void twait(Ctx& cx, int cls, uint32_t bar, uint32_t parity, int tag) {
#if PNR_TC_STATS
  long long t0 = clock64();
  mbar_wait(bar, parity, cx.err, tag);
  cx.w[cls] += clock64() - t0;
#else
  mbar_wait(bar, parity, cx.err, tag);
#endif
}

// Group waits: the 8 epilogue warps (and the 2 gather warps) always wait for the same barrier phase at the same
// point of their programs.  Every waiting warp polls shared memory (a suspended try_wait comes back a few hundred
// cycles later), and those polls compete with the TMA fills and the tensor core's operand reads for the
// shared-memory pipe (spinning instead of suspending costs 14 % of the kernel).  So ONE warp of the group waits on
// the mbarrier and the others block on a named hardware barrier, which does not poll.
__device__ __forceinline__ void epi_group_wait(Ctx& cx, int cls, uint32_t bar, uint32_t parity, int tag) {
  if ((threadIdx.x >> 5) == 4) twait(cx, cls, bar, parity, tag);
  asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void gather_group_wait(Ctx& cx, int cls, uint32_t bar, uint32_t parity, int tag) {
  if ((threadIdx.x >> 5) == 2) twait(cx, cls, bar, parity, tag);
  asm volatile("bar.sync 6, 64;" ::: "memory");
}

// ---- producer side helpers ----------------------------------------------------------------------
__device__ __forceinline__ long long* ts_slot(int which, int idx) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  return reinterpret_cast<long long*>(smem_raw + OFF_BARS + 320) + which * NB_ST + idx;
}
// Weight chunk / operand slice loads: 2-D tiled TMA whose transaction bytes complete on the LEADER
// CTA's full barrier (no peer->leader relay hop).  The leader expects both CTAs' bytes.
__device__ __forceinline__ void load_b(Ctx& cx, Ring& rb, const CUtensorMap* tm, uint32_t byte_off) {
  twait(cx, 0, rb.empty_bar(), rb.phase ^ 1, 201);
#if PNR_TC_STATS
  {  // commit -> producer-observed-empty latency
    long long tc = *(volatile long long*)ts_slot(1, rb.idx);
    if (tc != 0) { cx.w[2] += clock64() - tc; cx.w[3] += 1; }
  }
#endif
  const int row0 = (int)(byte_off >> 8);
  const uint32_t dst = cx.smem + OFF_BRING + rb.idx * B_CHUNK, fb = rb.full_bar();
  if (PROD_ELECT()) {  // operands computed in warp-uniform code -> uniform registers, no R2UR waterfall
#ifdef PNR_EXP_NOLOAD  // timing experiment: a quarter of the weight traffic (results are garbage)
    if (cx.rank == 0) mbar_expect_tx(fb, 2 * (B_CHUNK / 4));
    for (int q = 0; q < B_SPLIT / 4; ++q)
      tma_load_2d_pair(dst + q * (B_CHUNK / B_SPLIT), tm, 0, row0 + q * (B_CHUNK / B_SPLIT / 256), fb);
#else
    if (cx.rank == 0) mbar_expect_tx(fb, 2 * B_CHUNK);
#pragma unroll
    for (int q = 0; q < B_SPLIT; ++q)
      tma_load_2d_pair(dst + q * (B_CHUNK / B_SPLIT), tm, 0, row0 + q * (B_CHUNK / B_SPLIT / 256), fb);
#endif
#if PNR_TC_STATS
    *(volatile long long*)ts_slot(0, rb.idx) = clock64();
#endif
  }
  PROD_SYNC();
  rb.advance();
}
__device__ __forceinline__ void load_a(Ctx& cx, Ring& ra, const CUtensorMap* tm, size_t byte_off) {
  twait(cx, 1, ra.empty_bar(), ra.phase ^ 1, 202);
  const uint32_t dst = cx.smem + OFF_BRING + ra.idx * B_CHUNK, fb = ra.full_bar();  // an 8 KB slice in a 16 KB slot
  if (PROD_ELECT()) {
    if (cx.rank == 0) mbar_expect_tx(fb, 2 * A_SLICE);
#pragma unroll
    for (int q = 0; q < A_SPLIT; ++q)
      tma_load_2d_pair(dst + q * (A_SLICE / A_SPLIT), tm, 0, (int)(byte_off >> 8) + q * (A_SLICE / A_SPLIT / 256), fb);
  }
  PROD_SYNC();
  ra.advance();
}
// weight chunk (slice s, column block nb) of a GEMM group for this CTA
__device__ __forceinline__ uint32_t wchunk(uint32_t off, int s, int nb, uint32_t rank) {
  return off + (uint32_t)((s * 2 + nb) * 2 + rank) * B_CHUNK;
}

// ---- MMA side helpers -----------------------------------------------------------------------------
// Executed by ALL lanes of warp 1 in warp-uniform control flow (descriptors stay in uniform
// registers; only the tcgen05 / arrive instructions themselves are issued by one elected lane).
// Only the leader CTA runs this role: waits until both CTAs' copies of the ring slot have landed
// (both complete on its barrier), issues 4 MMAs (K=64) and releases the slot in both CTAs.
template <bool SM>
__device__ __forceinline__ void mma_step_b(Ctx& cx, Ring& rb, uint32_t a_addr, uint32_t d_col, bool first) {
  // (cx.idesc: instruction descriptor of the kernel's operand format, M=128 N=256)
  twait(cx, 0, rb.full_bar(), rb.phase, 301);
  if (cx.rank == 0) {
#if PNR_TC_STATS
    cx.w[5] += clock64() - *(volatile long long*)ts_slot(0, rb.idx);  // load issue -> MMA-observed-full
#endif
    tc_fence_after();
#ifdef PNR_EXP_N64  // timing experiment: quarter-size MMAs (results are garbage)
    const uint32_t idesc = idesc_bf16_f32(128, 64);
#else
    const uint32_t idesc = cx.idesc;
#endif
    const uint32_t b_addr = cx.smem + OFF_BRING + rb.idx * B_CHUNK;
    const uint64_t da0 = smem_desc(a_addr, ROWS * 16, 128);
    const uint64_t db0 = smem_desc(b_addr, 128 * 16, 128);
    const uint32_t dcol = cx.tmem + d_col;
    const uint32_t ebar = rb.empty_bar();
    if (mma_elect<SM>()) {
#pragma unroll
#ifdef PNR_EXP_ONE_MMA  // timing experiment: a quarter of the MMA instructions (results are garbage)
      for (int kk = 0; kk < 1; ++kk) {
#else
      for (int kk = 0; kk < 4; ++kk) {
#endif
        // advancing K by 16 elements = 2 core-matrix panels: add to the (addr>>4) field only
        mma_bf16<2>(dcol, da0 + (uint64_t)(kk * 2 * (ROWS * 16) >> 4), db0 + (uint64_t)(kk * 2 * (128 * 16) >> 4), idesc,
                    (first && kk == 0) ? 0u : 1u);
      }
      mma_commit<2>(ebar, 0x3);
#if PNR_TC_STATS
      *(volatile long long*)ts_slot(1, rb.idx) = clock64();
#endif
    }
    mma_sync<SM>();
  }
  rb.advance();
}
template <bool SM>
__device__ __forceinline__ void signal(const Ctx& cx, int bar_idx) {
  if (mma_elect<SM>()) mma_commit<2>(cx.bar(bar_idx), 0x3);
  mma_sync<SM>();
}

// x[:, all 512] (+)= A @ W^T over `nks` slices (k-outer) with A streamed through the ring: per slice the
// ring carries [A slice][W block 0][W block 1]; the A slot is released after both column blocks.
template <bool SM>
__device__ __forceinline__ void gemm_from_ring(Ctx& cx, Ring& rb, int nks, uint32_t xcol, bool overwrite) {
  for (int s = 0; s < nks; ++s) {
    twait(cx, 1, rb.full_bar(), rb.phase, 302);
    const uint32_t a_addr = cx.smem + OFF_BRING + rb.idx * B_CHUNK;
    const uint32_t a_empty = rb.empty_bar();
    rb.advance();
    for (int nb = 0; nb < 2; ++nb) mma_step_b<SM>(cx, rb, a_addr, xcol + nb * 128, overwrite && s == 0);
    if (mma_elect<SM>()) mma_commit<2>(a_empty, 0x3);
    mma_sync<SM>();
  }
}
// NET = S_x @ W0^T, n-outer; waits for operand slices as the epilogue publishes them
template <bool SM>
__device__ __forceinline__ void gemm_fc0(Ctx& cx, Ring& rb, uint32_t netcol, uint32_t sx_phase) {
  for (int nb = 0; nb < 2; ++nb) {
    for (int s = 0; s < DH / KS; ++s) {
      if (nb == 0 && cx.rank == 0) twait(cx, 2, cx.bar(SX_READY + s), sx_phase, 310 + s);
      mma_step_b<SM>(cx, rb, cx.smem + OFF_SX + s * A_SLICE, netcol + nb * 128, s == 0);
    }
    signal<SM>(cx, NET_READY + nb);
  }
}
// X += H @ W1^T in four quarters (s 0-3 | nb 0,1), (s 4-7 | nb 0,1): the first half of X's columns
// is final one quarter before the end, so its epilogue overlaps the last quarter, and the second
// half of H is only needed from the third quarter on.  The producer streams chunks in this order.
template <bool SM>
__device__ __forceinline__ void gemm_fc1(Ctx& cx, Ring& rb, uint32_t xcol, uint32_t h_phase) {
  for (int sh = 0; sh < 2; ++sh)
    for (int nb = 0; nb < 2; ++nb) {
      for (int s = sh * 4; s < sh * 4 + 4; ++s) {
        if (nb == 0) twait(cx, 3, cx.bar(H_READY + s), h_phase, 320 + s);
        mma_step_b<SM>(cx, rb, cx.smem + OFF_H + s * A_SLICE, xcol + nb * 128, false);
      }
      if (sh == 1) signal<SM>(cx, X_READY + nb);
    }
}

// ---- epilogue helpers -------------------------------------------------------------------------------
struct Epi {
  int q, cs, h, row, lane;  // TMEM quadrant, column split, column half of the 2x2 layout, tile row
  uint32_t lane_addr;       // (q*32) << 16
};
__device__ __forceinline__ Epi make_epi(int warp, int lane) {
  Epi e;
  e.q = warp & 3;
  e.cs = (warp - 4) >> 2;
  e.h = e.q >> 1;
  e.row = (e.q & 1) * 32 + lane;
  e.lane = lane;
  e.lane_addr = (uint32_t)(e.q * 32) << 16;
  return e;
}
// feature index of column i of the 32-column group (nb, half) held by this thread
__device__ __forceinline__ int feat0(const Epi& e, int nb, int half) { return nb * 256 + e.h * 128 + e.cs * 64 + half * 32; }

// Biases of the 64 features (2 halves x 32 columns) one epilogue thread handles in column block nb.
// With ~225 KB of the SM given to shared memory there is almost no L1 left, so a bias load is an L2
// round trip (~700 cycles): the loads are issued BEFORE the mbarrier wait that precedes the stage.
struct BiasRegs {
  float4 v[16];
};
__device__ __forceinline__ void prefetch_bias(BiasRegs& b, const Epi& e, const float* __restrict__ bias, int nb) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float4* b4 = reinterpret_cast<const float4*>(bias + feat0(e, nb, half));
#pragma unroll
    for (int j = 0; j < 8; ++j) b.v[half * 8 + j] = __ldg(b4 + j);
  }
}

// TMEM (64 columns of block nb) -> relu(v + bias) -> 16-bit operand panels in `dst_off`; publishes slice
template <int FMT>
__device__ __forceinline__ void epi_to_operand(const Ctx& cx, const Epi& e, uint32_t col, int nb, const BiasRegs& bias,
                                               uint32_t dst_off, int ready_bar0) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld32(cx.tmem + e.lane_addr + col + nb * 128 + e.cs * 64 + half * 32, r);
    tmem_ld_wait();
    const int f0 = feat0(e, nb, half);
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 8 features -> one 16-byte k-group entry
      const float4 ba = bias.v[half * 8 + 2 * j], bb = bias.v[half * 8 + 2 * j + 1];
      float v0 = fmaxf(__uint_as_float(r[8 * j + 0]) + ba.x, 0.f), v1 = fmaxf(__uint_as_float(r[8 * j + 1]) + ba.y, 0.f);
      float v2 = fmaxf(__uint_as_float(r[8 * j + 2]) + ba.z, 0.f), v3 = fmaxf(__uint_as_float(r[8 * j + 3]) + ba.w, 0.f);
      float v4 = fmaxf(__uint_as_float(r[8 * j + 4]) + bb.x, 0.f), v5 = fmaxf(__uint_as_float(r[8 * j + 5]) + bb.y, 0.f);
      float v6 = fmaxf(__uint_as_float(r[8 * j + 6]) + bb.z, 0.f), v7 = fmaxf(__uint_as_float(r[8 * j + 7]) + bb.w, 0.f);
      uint32_t kg = (uint32_t)(f0 >> 3) + j;
      uint32_t addr = cx.smem + dst_off + kg * (ROWS * 16) + e.row * 16;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2<FMT>(v0, v1)),
                   "r"(pack2<FMT>(v2, v3)), "r"(pack2<FMT>(v4, v5)), "r"(pack2<FMT>(v6, v7))
                   : "memory");
    }
  }
  tc_fence_before();
  fence_proxy_async_smem();
  __syncwarp();
  if (e.lane == 0) mbar_arrive_cluster(cx.bar(ready_bar0 + nb * 4 + e.h * 2 + e.cs), 0);
}

__device__ __forceinline__ void setup_barriers(const Ctx& cx) {
  for (int i = 0; i < NB_ST; ++i) { mbar_init(cx.bar(B_FULL + i), 1); mbar_init(cx.bar(B_EMPTY + i), 1); }
  mbar_init(cx.bar(X_READY), 1);
  mbar_init(cx.bar(X_READY + 1), 1);
  mbar_init(cx.bar(NET_READY), 1);
  mbar_init(cx.bar(NET_READY + 1), 1);
  for (int i = 0; i < 8; ++i) { mbar_init(cx.bar(SX_READY + i), 4); mbar_init(cx.bar(H_READY + i), 4); }
  mbar_init(cx.bar(XP_DONE), 16);
  mbar_init(cx.bar(ZC_READY), 2);   // the two gather warps of THIS CTA
  mbar_init(cx.bar(ZC_TAKEN), 1);   // this CTA's producer
  for (int i = 0; i < 2 * NB_ST; ++i) *ts_slot(i / NB_ST, i % NB_ST) = 0;
  fence_mbar_init();
}

// View mean-pool of one A tile's residual stream (TMEM X) -> pooled x (fp32) rows krow0.. of this CTA's
// staging buffer (the group's B tile reads them back).
// Runs once per tile, i.e. with cold instruction cache lines every time (the kernel is far larger
// than the I-cache and the other roles keep running): one compact out-of-line body per view count,
// selected once, instead of code specialised on p.ns inline (which spread the executed path over
// ~100 KB of SASS and cost ~10K cycles per tile in instruction fetch alone).  NS = 0: any view count.
template <int NS>
__device__ __noinline__ long long pool_tile(const Params& p, Ctx cx, const Epi e, int tile, int krow0, uint32_t xcol,
                                            uint32_t xph, uint32_t it, uint8_t* smem_raw) {
  const int ns = NS ? NS : p.ns;
  const int lane = e.lane;
  const float* biasA = reinterpret_cast<const float*>(p.w + p.off_biasA);
  int v;
  bool valid;
  long long gp = tileA_point(tile, (int)cx.rank, e.row, ns, p.ppw, v, valid);
  valid = valid && v == 0 && gp < p.P;
  const float* bP = biasA + (size_t)p.n_pre * DH;
  const float inv = 1.0f / (float)ns;
  const int rb_ = krow0 + e.row / ns;  // staging row of this lane's point
  float4* stage4 = reinterpret_cast<float4*>(p.stage) + (size_t)blockIdx.x * (ROWS * DH / 4);
  // A point whose NS rows straddle lanes 31|32 (NS not a divisor of 32) has its first `rem` rows in
  // the lower warp and the other NS-rem in the partner warp (same columns, next TMEM quadrant):
  // the upper warp pre-sums its rows and hands one value per column over through the idle S_x.
  const int rem = 32 % ns, nup = ns - rem;
  const bool straddle = rem != 0;
  const bool upper = (e.q & 1) != 0;
  const bool takes_spill = straddle && !upper && lane == 32 - rem;
  const int pair_id = e.h * 2 + e.cs;                 // warps (q even, q odd) with equal h, cs
  float4* spill = reinterpret_cast<float4*>(smem_raw + OFF_SX) + (size_t)pair_id * (2 * 2 * 8);
  long long work = 0;
#pragma unroll 1
  for (int nb = 0; nb < 2; ++nb) {
    BiasRegs bp;
    prefetch_bias(bp, e, bP, nb);
    epi_group_wait(cx, 0, cx.bar(X_READY + nb), xph, 405 + nb * 1000 + (int)it * 10000);
    tc_fence_after();
    const long long tq0 = clock64();
#pragma unroll
    for (int half = 0; half < 2; ++half) {  // (unrolled: bp.v is indexed by half)
      uint32_t r[32];
      tmem_ld32(cx.tmem + e.lane_addr + xcol + nb * 128 + e.cs * 64 + half * 32, r);
      tmem_ld_wait();
      float4* sp = spill + (nb * 2 + half) * 8;  // 32 columns
      if (straddle) {
        if (upper) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float u[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float x = __uint_as_float(r[4 * j + i]);
              float a = __shfl_sync(0xffffffffu, x, 0);
              if (NS) {
#pragma unroll
                for (int k = 1; k < (NS ? NS - 32 % (NS ? NS : 1) : 1); ++k) a += __shfl_sync(0xffffffffu, x, k);
              } else {
#pragma unroll 1
                for (int k = 1; k < nup; ++k) a += __shfl_sync(0xffffffffu, x, k);
              }
              u[i] = a;
            }
            if (lane == 0) sp[j] = make_float4(u[0], u[1], u[2], u[3]);
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + pair_id) : "memory");
      }
      // stage[cta][nb][h][jq (32)][row (64)][4]
      float4* dst = stage4 + ((nb * 2 + e.h) * 32 + (e.cs * 16 + half * 8)) * 64 + rb_;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = __uint_as_float(r[4 * j + i]);
          float s = x;
          if (NS) {
#pragma unroll
            for (int k = 1; k < NS; ++k) {
              const float y = __shfl_down_sync(0xffffffffu, x, k);
              s += (lane + k <= 31) ? y : 0.f;
            }
          } else {
#pragma unroll 1
            for (int k = 1; k < ns; ++k) {
              const float y = __shfl_down_sync(0xffffffffu, x, k);
              s += (lane + k <= 31) ? y : 0.f;
            }
          }
          s4[i] = s;
        }
        if (takes_spill) {
          const float4 u = sp[j];
          s4[0] += u.x; s4[1] += u.y; s4[2] += u.z; s4[3] += u.w;
        }
        if (valid) {
          const float4 bb = bp.v[half * 8 + j];
          dst[(size_t)j * 64] = make_float4(s4[0] * inv + bb.x, s4[1] * inv + bb.y, s4[2] * inv + bb.z, s4[3] * inv + bb.w);
        }
      }
    }
    work += clock64() - tq0;
  }
  return work;
}

// =============================================================================================
// B-tile pieces of the epilogue warps (executed once per group: kept out of line so that the A-tile
// loop, which runs nA times as often, keeps a small instruction footprint)
// =============================================================================================
constexpr int OFF_WOUT = OFF_SX;                       // lin_out weights [4][512] + bias [4] (fp32), during the head
constexpr int OFF_PART = OFF_SX + 8704;                // [4 outputs][4 column parts][64 rows] partial sums

// pooled rows of this CTA's staging buffer -> TMEM residual at column `xc` (fp32) and, with `to_sx`,
// relu -> bf16 operand in S_x (publishing the 8 K slices)
template <int FMT>
__device__ __noinline__ void load_x_tile(const Params& p, const Ctx cx, const Epi e, uint32_t xc, bool valid, bool to_sx) {
  const float4* stage4 = reinterpret_cast<const float4*>(p.stage) + (size_t)blockIdx.x * (ROWS * DH / 4);
#pragma unroll 1
  for (int nb = 0; nb < 2; ++nb) {
    const float4* src = stage4 + ((nb * 2 + e.h) * 32 + e.cs * 16) * 64 + e.row;
    float4 t[16];  // both 32-column halves in flight: the loads are L2 round trips (.cg: written by this CTA)
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = valid ? __ldcg(src + (size_t)j * 64) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[4 * j] = __float_as_uint(t[half * 8 + j].x);
        r[4 * j + 1] = __float_as_uint(t[half * 8 + j].y);
        r[4 * j + 2] = __float_as_uint(t[half * 8 + j].z);
        r[4 * j + 3] = __float_as_uint(t[half * 8 + j].w);
      }
      tmem_st32(cx.tmem + e.lane_addr + xc + nb * 128 + e.cs * 64 + half * 32, r);
      if (to_sx) {
        const int f0 = feat0(e, nb, half);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t kg = (uint32_t)(f0 >> 3) + j;
          uint32_t addr = cx.smem + OFF_SX + kg * (ROWS * 16) + e.row * 16;
          auto rl = [&](int i) { return fmaxf(__uint_as_float(r[8 * j + i]), 0.f); };
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2<FMT>(rl(0), rl(1))),
                       "r"(pack2<FMT>(rl(2), rl(3))), "r"(pack2<FMT>(rl(4), rl(5))), "r"(pack2<FMT>(rl(6), rl(7)))
                       : "memory");
        }
      }
    }
    tmem_st_wait();
    tc_fence_before();
    fence_proxy_async_smem();
    __syncwarp();
    if (to_sx && e.lane == 0) mbar_arrive_cluster(cx.bar(SX_READY + nb * 4 + e.h * 2 + e.cs), 0);
  }
}

// lin_out(relu(x)) on CUDA cores from the fp32 residual in TMEM + sigmoid/relu head
// (src/model/resnetfc.py:234-235, models.py.backup2:274-281); `gp` = this thread's output point (< 0: none)
__device__ __noinline__ void head_tile(const Params& p, Ctx cx, const Epi e, uint32_t xcol, uint32_t xph, bool wait_x,
                                       long long gp_row, uint8_t* smem_raw) {
  float* s_wout = reinterpret_cast<float*>(smem_raw + OFF_WOUT);
  float* s_part = reinterpret_cast<float*>(smem_raw + OFF_PART);
  const float* biasB = reinterpret_cast<const float*>(p.w + p.off_biasB);
  const float* bO = biasB + (size_t)p.n_post * DH;
  const int warp = threadIdx.x >> 5, lane = e.lane;
  {  // S_x is idle (the last fc_0 has completed): stage the head weights there
    const float4* wo = reinterpret_cast<const float4*>(p.w + p.off_lin_out);
    const int t = threadIdx.x - 128;  // 256 epilogue threads
    for (int i = t; i < (4 * DH + 4) / 4; i += 256) reinterpret_cast<float4*>(s_wout)[i] = __ldg(wo + i);
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int nb = 0; nb < 2; ++nb) {
    BiasRegs bo;
    prefetch_bias(bo, e, bO, nb);
    if (wait_x) {
      epi_group_wait(cx, 0, cx.bar(X_READY + nb), xph, 505);
      tc_fence_after();
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
      tmem_ld32(cx.tmem + e.lane_addr + xcol + nb * 128 + e.cs * 64 + half * 32, r);
      tmem_ld_wait();
      if (nb == 1 && half == 1) {  // X is in registers: the next tile may reuse its TMEM half
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(cx.bar(XP_DONE), 0);
      }
      const int f0 = feat0(e, nb, half);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = bo.v[half * 8 + j];
        const float xs[4] = {fmaxf(__uint_as_float(r[4 * j + 0]) + bb.x, 0.f), fmaxf(__uint_as_float(r[4 * j + 1]) + bb.y, 0.f),
                             fmaxf(__uint_as_float(r[4 * j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(r[4 * j + 3]) + bb.w, 0.f)};
#pragma unroll
        for (int o = 0; o < 4; ++o) {  // one 128-bit broadcast read per (output, 4 features)
          const float4 w4 = *reinterpret_cast<const float4*>(s_wout + o * DH + f0 + 4 * j);
          part[o] = fmaf(xs[0], w4.x, part[o]);
          part[o] = fmaf(xs[1], w4.y, part[o]);
          part[o] = fmaf(xs[2], w4.z, part[o]);
          part[o] = fmaf(xs[3], w4.w, part[o]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) s_part[(o * 4 + e.h * 2 + e.cs) * 64 + e.row] = part[o];
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (warp < 6) {  // 64 threads: one per row (warps 4, 5 hold rows 0..63 in e.row order)
    const int row = (warp - 4) * 32 + lane;
    if (gp_row >= 0) {
      float o4[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float s = s_wout[4 * DH + o];
#pragma unroll
        for (int q = 0; q < 4; ++q) s += s_part[(o * 4 + q) * 64 + row];
        if (p.apply_head) s = (o < 3) ? 1.f / (1.f + __expf(-s)) : fmaxf(s, 0.f);
        o4[o] = s;
      }
      // streaming store: consumed once by the compositing kernel, must not displace the L2-resident rings
      __stcs(reinterpret_cast<float4*>(p.out) + gp_row, make_float4(o4[0], o4[1], o4[2], o4[3]));
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");  // s_wout / s_part (= S_x) may be overwritten from here on
}

// =============================================================================================
// The fused kernel
// =============================================================================================
template <bool SM, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) mlp_fused_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Ctx cx;
  cx.smem = smem_u32(smem_raw);
  cx.bars = cx.smem + OFF_BARS;
  cx.rank = cluster_ctarank();
  cx.err = p.err;
  cx.idesc = FMT == 1 ? idesc_f16_f32(128, 256) : idesc_bf16_f32(128, 256);
  for (int i = 0; i < 6; ++i) cx.w[i] = 0;
  const long long t_begin = clock64();
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + OFF_BARS + N_BARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef PNR_EXP_NOLOAD
  for (int i = threadIdx.x; i < NB_ST * B_CHUNK / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem_raw + OFF_BRING)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
#endif
  if (threadIdx.x == 0) setup_barriers(cx);
  if (warp == 2) {
    tmem_alloc<2>(smem_u32(tmem_slot), 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  cx.tmem = *tmem_slot;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nsl = p.nks_z + p.nks_c;
  // A tile `tile` closes its group when it is the nA-th of the group or the pair's last tile
#define PNR_GROUP_END(k_, tile_) ((k_) == p.nA || (tile_) + npairs >= p.tilesA)

  if (warp == 0) {
    // ===================== producer =====================
    if (PROD_ENTER()) {
      Ring rb;
      rb.init(cx.bar(B_FULL), cx.bar(B_EMPTY), PNR_RING_A);
      uint32_t git = 0;
      int k = 0;
      for (int tile = pair; tile < p.tilesA; tile += npairs, ++git) {
        if (p.fused_gather) {  // this CTA's rows of the tile have been gathered (and are visible to TMA)
          twait(cx, 1, cx.bar(ZC_READY), git & 1, 203);
          if (PROD_ELECT()) mbar_arrive(cx.bar(ZC_TAKEN));
          PROD_SYNC();
        }
        const size_t zslot = p.zc_ring ? (size_t)pair * p.zc_ring + git % p.zc_ring : (size_t)tile;
        const size_t zt = ((zslot * 2 + cx.rank) * nsl) * A_SLICE;
        // (loops deliberately not unrolled: instruction footprint, see mbar_wait_slow)
        for (int s = 0; s < nsl; ++s) {
          load_a(cx, rb, &p.tm_zc, zt + (size_t)s * A_SLICE);
          for (int nb = 0; nb < 2; ++nb) load_b(cx, rb, &p.tm_w, wchunk(p.off_g1[0], s, nb, cx.rank));
        }
        for (int b = 0; b < p.n_pre; ++b) {
          for (int c = 0; c < 2 * (DH / KS); ++c)  // nb-outer, s-inner
            load_b(cx, rb, &p.tm_w, wchunk(p.off_g2[b], c % (DH / KS), c / (DH / KS), cx.rank));
          if (b + 1 < p.n_pre) {
            for (int s = 0; s < p.nks_z; ++s) {
              load_a(cx, rb, &p.tm_zc, zt + (size_t)s * A_SLICE);
              for (int nb = 0; nb < 2; ++nb) load_b(cx, rb, &p.tm_w, wchunk(p.off_g1[b + 1], s, nb, cx.rank));
            }
          }
          for (int c = 0; c < 16; ++c)  // same quartered order as gemm_fc1: (sh, nb, s4) = (c/8, (c/4)%2, c%4)
            load_b(cx, rb, &p.tm_w, wchunk(p.off_g3[b], (c >> 3) * 4 + (c & 3), (c >> 2) & 1, cx.rank));
        }
        if (++k, PNR_GROUP_END(k, tile)) {  // ---- B tile of the group
          k = 0;
          for (int j = 0; j < p.n_post; ++j) {
            const int b = p.n_pre + j;
#pragma unroll 1
            for (int c = 0; c < 2 * (DH / KS); ++c)
              load_b(cx, rb, &p.tm_w, wchunk(p.off_g2[b], c % (DH / KS), c / (DH / KS), cx.rank));
#pragma unroll 1
            for (int c = 0; c < 16; ++c)
              load_b(cx, rb, &p.tm_w, wchunk(p.off_g3[b], (c >> 3) * 4 + (c & 3), (c >> 2) & 1, cx.rank));
          }
        }
      }
      if (p.stats && cx.rank == 0 && lane == 0) {
        unsigned long long* st = p.stats + (size_t)pair * 16;
        st[6] = clock64() - t_begin;
        st[7] = cx.w[0];
        st[8] = cx.w[1];
        st[15] = cx.w[3] ? cx.w[2] / cx.w[3] : 0;  // mean commit -> empty-observed latency
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA only ======
    if (cx.rank == 0 && mma_enter<SM>()) {
      Ring rb;
      rb.init(cx.bar(B_FULL), cx.bar(B_EMPTY), PNR_RING_A);
      uint32_t use = 0;  // block counter: parity source of the SX/H barriers
      uint32_t tc = 0;   // tile counter (A and B tiles): X/NET halves swap every tile; XP_DONE parity
      int k = 0;
      for (int tile = pair; tile < p.tilesA; tile += npairs) {
        {
          const uint32_t xcol = (tc & 1) ? 256u : 0u, netcol = 256u - xcol;
          gemm_from_ring<SM>(cx, rb, nsl, xcol, true);
          // The previous tile's pool / head must have consumed its X_READY completions before they are
          // signalled again (a waiter that misses one completion of a 1-count mbarrier waits for
          // ever), and it must have finished reading the old X before fc_0 reuses it as NET.
          if (tc > 0) twait(cx, 4, cx.bar(XP_DONE), (tc - 1) & 1, 330);
          signal<SM>(cx, X_READY);
          signal<SM>(cx, X_READY + 1);
          for (int b = 0; b < p.n_pre; ++b, ++use) {
            gemm_fc0<SM>(cx, rb, netcol, use & 1);
            if (b + 1 < p.n_pre) gemm_from_ring<SM>(cx, rb, p.nks_z, xcol, false);
            gemm_fc1<SM>(cx, rb, xcol, use & 1);
          }
          ++tc;
        }
        if (++k, PNR_GROUP_END(k, tile)) {  // ---- B tile: X was loaded by the epilogue warps (SX_READY)
          k = 0;
          const uint32_t xcol = (tc & 1) ? 256u : 0u, netcol = 256u - xcol;
          twait(cx, 4, cx.bar(XP_DONE), (tc - 1) & 1, 530);  // the pool has read the A tile's X (= this NET half)
          for (int j = 0; j < p.n_post; ++j, ++use) {
            gemm_fc0<SM>(cx, rb, netcol, use & 1);
            gemm_fc1<SM>(cx, rb, xcol, use & 1);
          }
          ++tc;
        }
      }
      if (p.stats && cx.rank == 0 && lane == 0) {
        unsigned long long* st = p.stats + (size_t)pair * 16;
        st[0] = clock64() - t_begin;
        for (int i = 0; i < 5; ++i) st[1 + i] = cx.w[i];
      }
    }
  } else if (warp < 4) {
    // ===================== gather warps (2, 3): operand image of the NEXT tiles ==================
    if (p.fused_gather) {
      uint32_t git = 0;
      for (int tile = pair; tile < p.tilesA; tile += npairs, ++git) {
        const int row = (warp - 2) * 32 + lane;
        int v;
        bool valid;
        long long gp = tileA_point(tile, (int)cx.rank, row, p.ns, p.ppw, v, valid);
        valid = valid && gp < p.P;
        // 2-deep ring per pair: slot git % 2 last held tile git-2.  The producer "takes" tile git-1 only after it has
        // issued every load of tile git-2, and the ring slots those loads used have been recycled since (so they
        // have landed): once ZC_TAKEN(git-1) completed the slot is free.  The gather therefore runs exactly one
        // tile ahead of the MMAs, and the image (2 x 2 x nsl x 8 KB per pair) stays resident in L2.
        if (git >= 1) gather_group_wait(cx, 2, cx.bar(ZC_TAKEN), (git - 1) & 1, 204);
        const size_t zslot = p.zc_ring ? (size_t)pair * p.zc_ring + git % p.zc_ring : (size_t)tile;
        uint8_t* base = const_cast<uint8_t*>(p.zc) + ((zslot * 2 + cx.rank) * nsl) * A_SLICE + (size_t)row * 16;
        gather_row_to_zc<FMT>(p.sc, p.xyz, p.viewdirs, p.rays, p.zsamp, p.K, gp, v, valid, p.nks_z, p.nks_c, base);
        // generic-proxy global writes -> visible to the async proxy (TMA) of this SM before the signal.  (The wait
        // above also keeps ZC_READY from completing twice before the producer looks.)
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(cx.bar(ZC_READY));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const Epi e = make_epi(warp, lane);
    const float* biasA = reinterpret_cast<const float*>(p.w + p.off_biasA);
    const float* biasB = reinterpret_cast<const float*>(p.w + p.off_biasB);
    const float* bias0 = reinterpret_cast<const float*>(p.w + p.off_bias0);
    uint32_t xph = 0, nph = 0, tc = 0;
    int k = 0, group_tile0 = pair;
    for (int tile = pair; tile < p.tilesA; tile += npairs) {
      {
        const uint32_t xcol = (tc & 1) ? 256u : 0u, netcol = 256u - xcol;
        for (int b = 0; b < p.n_pre; ++b) {
          long long t0 = 0;
          BiasRegs br;
          for (int nb = 0; nb < 2; ++nb) {
            prefetch_bias(br, e, biasA + (size_t)b * DH, nb);
            epi_group_wait(cx, 0, cx.bar(X_READY + nb), xph, 401 + nb * 1000 + (int)tc * 10000 + (int)cx.rank * 100000000);
            tc_fence_after();
            t0 = clock64();
            epi_to_operand<FMT>(cx, e, xcol, nb, br, OFF_SX, SX_READY);
            cx.w[2] += clock64() - t0;
          }
          xph ^= 1;
          for (int nb = 0; nb < 2; ++nb) {
            prefetch_bias(br, e, bias0 + (size_t)b * DH, nb);
            epi_group_wait(cx, 1, cx.bar(NET_READY + nb), nph, 402 + nb);
            tc_fence_after();
            t0 = clock64();
            epi_to_operand<FMT>(cx, e, netcol, nb, br, OFF_H, H_READY);
            cx.w[3] += clock64() - t0;
          }
          nph ^= 1;
        }
        // ---- view mean-pool of the residual stream -> pooled x (fp32) rows k*ppw.. of the staging buffer ----
        long long tp;
        const int krow0 = k * p.ppw;
        switch (p.ns) {
          case 1: tp = pool_tile<1>(p, cx, e, tile, krow0, xcol, xph, tc, smem_raw); break;
          case 2: tp = pool_tile<2>(p, cx, e, tile, krow0, xcol, xph, tc, smem_raw); break;
          case 3: tp = pool_tile<3>(p, cx, e, tile, krow0, xcol, xph, tc, smem_raw); break;
          case 4: tp = pool_tile<4>(p, cx, e, tile, krow0, xcol, xph, tc, smem_raw); break;
          default: tp = pool_tile<0>(p, cx, e, tile, krow0, xcol, xph, tc, smem_raw); break;
        }
        cx.w[4] += tp;  // pool work only (waits excluded)
        xph ^= 1;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(cx.bar(XP_DONE), 0);
        ++tc;
      }
      if (++k, PNR_GROUP_END(k, tile)) {
        // ===== B tile: the group's pooled points (row r = kk*ppw + pl <-> A tile group_tile0 + kk*npairs) =====
        const uint32_t xcol = (tc & 1) ? 256u : 0u, netcol = 256u - xcol;
        const int kk = e.row / p.ppw, pl = e.row - kk * p.ppw;
        const long long gp = (long long)((group_tile0 + kk * npairs) * 2 + (int)cx.rank) * p.ppw + pl;
        const bool valid = kk < k && gp < p.P;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // all pooled rows of this CTA are written
        load_x_tile<FMT>(p, cx, e, xcol, valid, p.n_post > 0);
        for (int j = 0; j < p.n_post; ++j) {
          const int b = p.n_pre + j;
          BiasRegs br;
          for (int nb = 0; nb < 2; ++nb) {
            prefetch_bias(br, e, bias0 + (size_t)b * DH, nb);
            epi_group_wait(cx, 1, cx.bar(NET_READY + nb), nph, 502 + nb);
            tc_fence_after();
            epi_to_operand<FMT>(cx, e, netcol, nb, br, OFF_H, H_READY);
          }
          nph ^= 1;
          if (j + 1 < p.n_post) {
            for (int nb = 0; nb < 2; ++nb) {
              prefetch_bias(br, e, biasB + (size_t)(j + 1) * DH, nb);
              epi_group_wait(cx, 0, cx.bar(X_READY + nb), xph, 501);
              tc_fence_after();
              epi_to_operand<FMT>(cx, e, xcol, nb, br, OFF_SX, SX_READY);
            }
            xph ^= 1;
          }
        }
        // output row of the 64 row-owner threads (warps 4, 5; their e.row = (warp-4)*32 + lane)
        head_tile(p, cx, e, xcol, xph, p.n_post > 0, valid ? gp : -1, smem_raw);
        if (p.n_post > 0) xph ^= 1;
        ++tc;
        k = 0;
        group_tile0 = tile + npairs;
      }
    }
    if (p.stats && cx.rank == 0 && warp == 4 && lane == 0) {
      unsigned long long* st = p.stats + (size_t)pair * 16;
      st[9] = clock64() - t_begin;
      for (int i = 0; i < 5; ++i) st[10 + i] = cx.w[i];
    }
  }
#undef PNR_GROUP_END
  __syncwarp();
  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(cx.tmem, 512);
}

// =============================================================================================
// host side
// =============================================================================================
// Per-device state: the barrier-fault word (mapped pinned host memory, so the tag survives a trap), the SM
// count, and whether the kernels' dynamic shared-memory attribute has been raised.  Guarded by one mutex;
// everything else in this file is stateless or thread-local.
struct DeviceState {
  int* err_host = nullptr;  // [0] fault tag (0 = none), [1] wait timeout in ms (0 = never give up)
  int* err_dev = nullptr;
  int sms = 0;
  bool attr_set[4] = {false, false, false, false};
};
static std::mutex g_dev_mutex;
static DeviceState g_dev[PNR_MAX_DEVICES];

static int wait_timeout_ms() {
  static const int ms = [] {  // PNR_WAIT_TIMEOUT_MS: 0 disables the trap (debuggers, compute-sanitizer, time-slicing)
    const char* e = getenv("PNR_WAIT_TIMEOUT_MS");
    return e ? atoi(e) : 2000;
  }();
  return ms < 0 ? 0 : ms;
}

static int device_state(DeviceState** out) {
  int dev = 0;
  PNR_CUDA(cudaGetDevice(&dev));
  PNR_CHECK_ARG(dev >= 0 && dev < PNR_MAX_DEVICES, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  DeviceState& d = g_dev[dev];
  if (!d.err_host) {
    int* h = nullptr;
    PNR_CUDA(cudaHostAlloc((void**)&h, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    h[0] = 0;
    h[1] = wait_timeout_ms();
    PNR_CUDA(cudaHostGetDevicePointer((void**)&d.err_dev, h, 0));
    PNR_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    if (d.sms <= 0) d.sms = 148;
    d.err_host = h;
  }
  *out = &d;
  return PNR_OK;
}

// a fault recorded by an earlier launch on this device is reported (once) by the next call
static int take_fault(DeviceState& d, const char* when) {
  int v = *(volatile int*)d.err_host;
  if (v == 0) return PNR_OK;
  d.err_host[0] = 0;
  set_err("tensor-core MLP pipeline: a barrier wait timed out in an earlier launch on this device (tag %d; %s); "
          "its results are invalid", v, when);
  return PNR_ERR_CUDA;
}

constexpr int ZC_RING = 2;  // operand-image tiles per pair (see the gather warps)
struct Plan {
  int ppw, ptile, tilesA, nA, nsl, pairs, zc_slots;
  uint8_t* zc;
  float* stage;
  size_t total;
};

static int pairs_for(int sms, int tiles) {
  int pairs = sms / 2;
  static int cap = [] { const char* e = getenv("PNR_MAX_PAIRS"); return e ? atoi(e) : 0; }();  // experiment knob
  if (cap > 0 && cap < pairs) pairs = cap;
  return tiles < pairs ? (tiles > 0 ? tiles : 1) : pairs;
}

// `ring`: the operand image is produced inside the kernel (gather warps) and needs only ZC_RING slots per pair
static Plan make_plan(const Layout& L, int sms, int ns, long long P, bool ring, void* ws, size_t ws_bytes) {
  Plan pl;
  pl.ppw = 64 / ns;
  pl.ptile = 2 * pl.ppw;
  pl.nA = 64 / pl.ppw;
  pl.tilesA = (int)ceil_div_ll(P, pl.ptile);
  pl.nsl = L.nks_z + L.nks_c;
  pl.pairs = pairs_for(sms, pl.tilesA);
  pl.zc_slots = (ring && pl.tilesA > ZC_RING * pl.pairs) ? ZC_RING * pl.pairs : pl.tilesA;
  Arena a(ws, ws_bytes);
  pl.zc = a.take<uint8_t>((size_t)pl.zc_slots * 2 * pl.nsl * A_SLICE);
  pl.stage = a.take<float>((size_t)pl.pairs * 2 * ROWS * DH);
  pl.total = a.off + 256;
  return pl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// A linear byte range viewed as rows of 256 B (128 x u16); a box of `box_rows` rows is a contiguous
// copy of box_rows*256 bytes, so the pre-packed operand images load exactly as laid out.
static int encode_rows256(CUtensorMap* tm, const void* base, size_t bytes, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PNR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) {
      set_err("cuTensorMapEncodeTiled is not available in this driver");
      return PNR_ERR_CUDA;
    }
    fn = (EncodeTiledFn)p;
  }
  PNR_CHECK_ARG(((uintptr_t)base & 15) == 0, "TMA source must be 16-byte aligned");
  cuuint64_t rows = (cuuint64_t)((bytes + 255) / 256);
  if (rows < (cuuint64_t)box_rows) rows = box_rows;
  cuuint64_t dims[2] = {128, rows};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PNR_ERR_CUDA;
  }
  return PNR_OK;
}

// The two tensor maps of a launch depend only on (base, bytes, box): encoded once per distinct operand
// image / scratch range and thread, not per launch.
struct TmapKey {
  const void* base;
  size_t bytes;
  int box_rows;
};
static int cached_tmap(CUtensorMap* out, const void* base, size_t bytes, int box_rows) {
  struct Entry {
    TmapKey k;
    CUtensorMap tm;
  };
  static thread_local Entry cache[8];
  static thread_local int used = 0, next = 0;
  for (int i = 0; i < used; ++i)
    if (cache[i].k.base == base && cache[i].k.bytes == bytes && cache[i].k.box_rows == box_rows) {
      *out = cache[i].tm;
      return PNR_OK;
    }
  PNR_TRY(encode_rows256(out, base, bytes, box_rows));
  Entry& e = cache[next];
  e.k = TmapKey{base, bytes, box_rows};
  e.tm = *out;
  next = (next + 1) % 8;
  if (used < 8) ++used;
  return PNR_OK;
}

static int launch_cluster(DeviceState& d, bool solo, int fmt, int pairs, const Params& p, cudaStream_t st) {
  void (*kern)(const Params) = fmt == 1 ? (solo ? mlp_fused_kernel<true, 1> : mlp_fused_kernel<false, 1>)
                                        : (solo ? mlp_fused_kernel<true, 0> : mlp_fused_kernel<false, 0>);
  {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    const int slot = (fmt == 1 ? 2 : 0) + (solo ? 1 : 0);
    if (!d.attr_set[slot]) {
      PNR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      d.attr_set[slot] = true;
    }
  }
  cudaLaunchConfig_t cfg;
  memset((void*)&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  PNR_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  launch_counter()++;
  return PNR_OK;
}

struct GatherArgs {
  const pnr_scene* sc;
  const float *xyz, *viewdirs, *rays, *z;
  int K;
};

static int run_fused(DeviceState& d, const pnr_mlp& m, const Layout& L, const Plan& pl, int ns, long long P, float* out,
                     int head, int fmt, cudaStream_t st, const GatherArgs* ga = nullptr) {
  PNR_TRY(take_fault(d, "reported before the next launch"));
  // algorithmic work (2*MACs of the nn.Linear layers, unpadded)
  const double mac_pre = (double)m.d_in * DH + (double)L.n_pre * m.d_latent * DH + 2.0 * L.n_pre * DH * DH;
  const double mac_post = 2.0 * L.n_post * DH * DH + (double)DH * m.d_out;
  Params p;
  memset(&p, 0, sizeof(p));
  p.w = (const uint8_t*)m.packed;
  for (int b = 0; b < PNR_MAX_BLOCKS; ++b) {
    p.off_g1[b] = L.off_g1[b];
    p.off_g2[b] = L.off_g2[b];
    p.off_g3[b] = L.off_g3[b];
  }
  p.off_biasA = L.off_biasA;
  p.off_biasB = L.off_biasB;
  p.off_bias0 = L.off_bias0;
  p.off_lin_out = L.off_lin_out;
  p.n_pre = L.n_pre;
  p.n_post = L.n_post;
  p.nks_z = L.nks_z;
  p.nks_c = L.nks_c;
  p.ns = ns;
  p.ppw = pl.ppw;
  p.nA = pl.nA;
  p.P = P;
  p.tilesA = pl.tilesA;
  p.zc = pl.zc;
  p.stage = pl.stage;
  p.out = out;
  p.apply_head = head;
  if (ga) {
    p.fused_gather = 1;
    p.zc_ring = (pl.zc_slots < pl.tilesA) ? ZC_RING : 0;  // ring per pair unless one slot per tile is smaller
    p.sc = *ga->sc;
    p.xyz = ga->xyz;
    p.viewdirs = ga->viewdirs;
    p.rays = ga->rays;
    p.zsamp = ga->z;
    p.K = ga->K;
  }
  p.err = d.err_dev;
  p.stats = g_stats_ptr;
  PNR_TRY(cached_tmap(&p.tm_w, m.packed, L.total, B_CHUNK / B_SPLIT / 256));
  PNR_TRY(cached_tmap(&p.tm_zc, pl.zc, (size_t)pl.zc_slots * 2 * pl.nsl * A_SLICE, A_SLICE / A_SPLIT / 256));
  {
    ProfScope ps(PROF_FUSED_MLP, 2.0 * (mac_pre * ns + mac_post) * (double)P, 0.0, st);
    // issuer mode by latent width (see mma_elect): PNR_SOLO_MMA=0/1 forces one of them
    static const int force = [] { const char* e = getenv("PNR_SOLO_MMA"); return e ? atoi(e) : -1; }();
    const bool solo = force >= 0 ? force != 0 : L.nks_z <= 4;
    PNR_TRY(launch_cluster(d, solo, fmt, pl.pairs, p, st));
  }
  return PNR_OK;
}

int tc_check(cudaStream_t st) {
  cudaError_t e = cudaStreamSynchronize(st);
  DeviceState* d = nullptr;
  int v = 0;
  if (device_state(&d) == PNR_OK) {
    v = *(volatile int*)d->err_host;
    d->err_host[0] = 0;
  }
  if (v != 0) {
    set_err("tensor-core MLP pipeline: barrier wait timed out (tag %d)%s", v,
            e != cudaSuccess ? "; the kernel trapped and the CUDA context is lost" : "");
    return PNR_ERR_CUDA;
  }
  if (e != cudaSuccess) {
    set_err("cudaStreamSynchronize -> %s", cudaGetErrorString(e));
    return PNR_ERR_CUDA;
  }
  return PNR_OK;
}

size_t net_tc_workspace(const pnr_scene& sc, const pnr_mlp& m, int SB, long long P) {
  if (tc_supported(m) != PNR_OK) return 256;
  DeviceState* d = nullptr;
  if (device_state(&d) != PNR_OK) return 256;
  Layout L = make_layout(m);
  return make_plan(L, d->sms, sc.ns, (long long)SB * P, true, nullptr, 0).total;
}

int net_forward_tc(const pnr_scene& sc, const pnr_mlp& m, const float* xyz, const float* viewdirs, const float* rays,
                   const float* z, int K, int SB, long long P, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  PNR_TRY(tc_supported(m));
  PNR_CHECK_ARG(SB == 1, "net_forward_tc: one object per call");
  PNR_UNSUPPORTED(sc.ns > 32, "more than 32 source views per object");
  PNR_UNSUPPORTED(sc.feat_dtype != PNR_BF16 && sc.feat_dtype != PNR_FP16, "tensor-core path needs a bf16- or f16-packed pyramid");
  PNR_CHECK_ARG(m.packed_dtype == sc.feat_dtype, "mlp.packed was built for dtype %d, the feature pyramid for %d", m.packed_dtype,
                sc.feat_dtype);
  for (int l = 0; l < sc.n_levels; ++l)
    PNR_UNSUPPORTED(sc.C[l] % 8 != 0 || sc.ch_off[l] % 8 != 0, "bf16 gather needs channel counts that are multiples of 8");
  PNR_CHECK_ARG(m.packed_bytes >= make_layout(m).total, "mlp.packed image too small");
  PNR_CHECK_ARG(((uintptr_t)out & 15) == 0, "net_forward: out must be 16-byte aligned");
  PNR_CHECK_ARG(P > 0 && P < (1LL << 31), "net_forward_tc: point count out of range");
  DeviceState* d = nullptr;
  PNR_TRY(device_state(&d));
  Layout L = make_layout(m);
  Plan pl = make_plan(L, d->sms, sc.ns, P, true, ws, ws_bytes);
  if (pl.total > ws_bytes + 256 || !ws) {
    set_err("net_forward_tc: workspace too small (%zu < %zu)", ws_bytes, pl.total);
    return PNR_ERR_WORKSPACE;
  }
  GatherArgs ga{&sc, xyz, viewdirs, rays, z, K};
  return run_fused(*d, m, L, pl, sc.ns, P, out, 1, sc.feat_dtype == PNR_FP16 ? 1 : 0, st, &ga);
}

size_t mlp_tc_rows_workspace(const pnr_mlp& m, int SB, int NS, int P) {
  if (tc_supported(m) != PNR_OK) return 256;
  DeviceState* d = nullptr;
  if (device_state(&d) != PNR_OK) return 256;
  Layout L = make_layout(m);
  return make_plan(L, d->sms, NS, (long long)SB * P, false, nullptr, 0).total;
}

int mlp_forward_tc_rows(const pnr_mlp& m, const float* zx, int SB, int NS, int P, float* out, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  PNR_TRY(tc_supported(m));
  PNR_UNSUPPORTED(NS > 32, "more than 32 source views per object");
  Layout L = make_layout(m);
  PNR_CHECK_ARG(m.packed_bytes >= L.total, "mlp.packed image too small");
  PNR_CHECK_ARG(((uintptr_t)out & 15) == 0, "mlp_forward: out must be 16-byte aligned");
  DeviceState* d = nullptr;
  PNR_TRY(device_state(&d));
  long long pts = (long long)SB * P;
  Plan pl = make_plan(L, d->sms, NS, pts, false, ws, ws_bytes);
  if (pl.total > ws_bytes + 256 || !ws) {
    set_err("mlp_forward_tc: workspace too small (%zu < %zu)", ws_bytes, pl.total);
    return PNR_ERR_WORKSPACE;
  }
  long long wrows = (long long)pl.tilesA * 128;
  const int fmt = m.packed_dtype == PNR_FP16 ? 1 : 0;
  rows_to_operand_kernel<<<(unsigned)ceil_div_ll(wrows, 8), 256, 0, st>>>(zx, m.d_latent, m.d_in, SB, NS, P, pl.ppw,
                                                                        pl.tilesA, L.nks_z, L.nks_c, fmt, pl.zc);
  PNR_LAUNCHED();
  return run_fused(*d, m, L, pl, NS, pts, out, 0, fmt, st);
}

}  // namespace pnr
