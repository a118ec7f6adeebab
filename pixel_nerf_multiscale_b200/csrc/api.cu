// extern "C" surface declared in include/pixelnerf_b200.h + host-side orchestration of one
// PixelNeRFNet.forward (models.py.backup2:155-282) and one NeRFRenderer.forward (nerf.py:251-303).
#include <cstdlib>
#include <stdarg.h>

#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace pnr {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
}
int64_t& launch_counter() {
  static thread_local int64_t n = 0;
  return n;
}

// NVTX ranges named after the reference's torch.autograd.profiler.record_function scopes (SURVEY.md section 5:
// renderer_forward nerf.py:264, renderer_composite :175, model_inference models.py.backup2:165), so a timeline
// of this library reads like one of the reference.  Free when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ---- profiling -------------------------------------------------------------------------------
struct ProfRec {
  int kind;
  double flops, bytes;
  cudaEvent_t a, b;
};
struct ProfState {
  bool on = false;
  std::vector<ProfRec*> recs;
};
static ProfState& prof() {
  static thread_local ProfState p;
  return p;
}
ProfScope::ProfScope(int kind_, double flops, double bytes, cudaStream_t st_) : kind(kind_), st(st_), rec(nullptr) {
  if (!prof().on) return;
  ProfRec* r = new ProfRec();
  r->kind = kind;
  r->flops = flops;
  r->bytes = bytes;
  cudaEventCreate(&r->a);
  cudaEventCreate(&r->b);
  cudaEventRecord(r->a, st);
  rec = r;
}
ProfScope::~ProfScope() {
  if (!rec) return;
  ProfRec* r = (ProfRec*)rec;
  cudaEventRecord(r->b, st);
  prof().recs.push_back(r);
}

int launch_pack_level(const float* src, int n_views, int C, int H, int W, void* dst, int dtype, cudaStream_t st);
int mlp_forward_tc_rows(const pnr_mlp& m, const float* zx, int SB, int NS, int P, float* out, void* ws,
                        size_t ws_bytes, cudaStream_t st);
size_t mlp_tc_rows_workspace(const pnr_mlp& m, int SB, int NS, int P);
int tc_check(cudaStream_t st);
void tc_set_stats(unsigned long long* p);

// pre-pool rows evaluated per internal chunk (rays are never split).  fp32 validation path: bounds the
// scratch.  bf16 path: the fused kernel keeps all intermediates on chip / in two small L2-resident rings, so
// a whole pass is ONE launch; the cap only keeps 32-bit point indices valid.
static const long long kChunkRowsF32 = 1LL << 18;
static const long long kChunkRowsBF16 = 1LL << 30;

static int validate_scene(const pnr_scene* sc) {
  PNR_CHECK_ARG(sc != nullptr, "scene is NULL");
  PNR_CHECK_ARG(sc->n_levels >= 1 && sc->n_levels <= PNR_MAX_LEVELS, "scene.n_levels=%d out of range", sc->n_levels);
  PNR_CHECK_ARG(sc->ns >= 1 && sc->n_views >= sc->ns && sc->n_views % sc->ns == 0, "scene.n_views=%d / ns=%d inconsistent",
                sc->n_views, sc->ns);
  PNR_CHECK_ARG(sc->cams != nullptr, "scene.cams is NULL");
  int off = 0;
  for (int l = 0; l < sc->n_levels; ++l) {
    PNR_CHECK_ARG(sc->level[l] != nullptr, "scene.level[%d] is NULL", l);
    PNR_CHECK_ARG(sc->C[l] > 0 && sc->H[l] > 1 && sc->W[l] > 1, "scene level %d has degenerate shape", l);
    PNR_CHECK_ARG(sc->ch_off[l] == off, "scene.ch_off[%d]=%d, expected %d", l, sc->ch_off[l], off);
    off += sc->C[l];
  }
  PNR_CHECK_ARG(off == sc->d_latent, "scene.d_latent=%d but levels sum to %d", sc->d_latent, off);
  PNR_CHECK_ARG(sc->feat_dtype == PNR_FP32 || sc->feat_dtype == PNR_BF16 || sc->feat_dtype == PNR_FP16, "scene.feat_dtype invalid");
  int dz = sc->use_xyz ? 3 : 1;
  int db = dz + ((sc->use_viewdirs && sc->use_code && sc->use_code_viewdirs) ? 3 : 0);
  int d = sc->use_code ? sc->num_freqs * 2 * db + (sc->include_input ? db : 0) : db;
  if (sc->use_viewdirs && !(sc->use_code && sc->use_code_viewdirs)) d += 3;
  PNR_CHECK_ARG(d == sc->d_in, "scene.d_in=%d inconsistent with code settings (expected %d)", sc->d_in, d);
  return PNR_OK;
}

static int validate_mlp(const pnr_mlp* m, const pnr_scene* sc, int precision) {
  PNR_CHECK_ARG(m != nullptr, "mlp is NULL");
  PNR_CHECK_ARG(m->n_blocks >= 1 && m->n_blocks <= PNR_MAX_BLOCKS, "mlp.n_blocks=%d out of range", m->n_blocks);
  PNR_CHECK_ARG(m->d_out == 4, "mlp.d_out must be 4");
  PNR_UNSUPPORTED(m->combine_type != 0, "combine_type other than average is not supported natively");
  if (sc) {
    // resnetfc.py:190-191: assert zx.size(-1) == d_latent + d_in
    PNR_CHECK_ARG(m->d_in == sc->d_in && m->d_latent == sc->d_latent,
                  "Input size %d != d_latent (%d) + d_in (%d)", sc->d_latent + sc->d_in, m->d_latent, m->d_in);
    PNR_UNSUPPORTED(sc->ns > 1 && m->combine_layer >= m->n_blocks,
                    "multi-view input with combine_layer >= n_blocks is not supported");
  }
  PNR_CHECK_ARG(m->lin_in_w && m->lin_in_b && m->lin_out_w && m->lin_out_b, "mlp linear pointers missing");
  PNR_CHECK_ARG(precision == PNR_FP32 || precision == PNR_BF16 || precision == PNR_FP16, "precision %d invalid", precision);
  if (precision != PNR_FP32) {
    PNR_CHECK_ARG(m->packed != nullptr, "mlp.packed is NULL (call pnr_mlp_pack)");
    PNR_CHECK_ARG(m->packed_dtype == precision, "mlp.packed was built for dtype %d, the call asks for %d", m->packed_dtype, precision);
    if (sc) PNR_CHECK_ARG(sc->feat_dtype == precision, "scene.feat_dtype %d does not match precision %d", sc->feat_dtype, precision);
  }
  return PNR_OK;
}

static pnr_scene object_scene(const pnr_scene& sc, int sb) {
  pnr_scene s = sc;
  s.n_views = sc.ns;
  s.cams = sc.cams + (size_t)sb * sc.ns * 16;
  size_t esz = sc.feat_dtype == PNR_FP32 ? 4 : 2;
  for (int l = 0; l < sc.n_levels; ++l)
    s.level[l] = (const char*)sc.level[l] + (size_t)sb * sc.ns * sc.H[l] * sc.W[l] * sc.C[l] * esz;
  return s;
}

static long long bf16_chunk_rows() {
  static long long rows = [] {  // PNR_CHUNK_ROWS_LOG2: experiment knob (scratch grows with it)
    const char* e = getenv("PNR_CHUNK_ROWS_LOG2");
    int l = e ? atoi(e) : 0;
    return (l >= 14 && l <= 30) ? (1LL << l) : kChunkRowsBF16;
  }();
  return rows;
}
static long long chunk_points(const pnr_scene& sc, int precision, int K, long long P) {
  long long rows = precision == PNR_FP32 ? kChunkRowsF32 : bf16_chunk_rows();
  long long pts = rows / sc.ns;
  if (K > 0) pts = (pts / K) * K;  // whole rays only
  if (pts < (K > 0 ? K : 1)) pts = (K > 0 ? K : 1);
  return pts < P ? pts : P;
}

static size_t net_chunk_workspace(const pnr_scene& sc, const pnr_mlp& m, int precision, long long Pc) {
  if (precision == PNR_FP32) {
    Arena a(nullptr, 0);
    a.take<float>((size_t)sc.ns * Pc * (sc.d_latent + sc.d_in));
    return a.off + 256 + mlp_f32_workspace(m, (long long)sc.ns * Pc, Pc);
  }
  return net_tc_workspace(sc, m, 1, Pc);
}

// evaluates the point network on SB x P points given either explicit xyz(+viewdirs) or rays+z
static int net_eval(const pnr_scene& sc, const pnr_mlp& m, int precision, const float* xyz, const float* vd,
                    const float* rays, const float* z, int K, int SB, long long P, float* out, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  NvtxRange nvtx("model_inference");
  long long Pc = chunk_points(sc, precision, K, P);
  for (int sb = 0; sb < SB; ++sb) {
    pnr_scene os = object_scene(sc, sb);
    for (long long p0 = 0; p0 < P; p0 += Pc) {
      long long n = (P - p0 < Pc) ? (P - p0) : Pc;
      long long g0 = (long long)sb * P + p0;
      const float* cx = xyz ? xyz + g0 * 3 : nullptr;
      const float* cv = vd ? vd + g0 * 3 : nullptr;
      const float* cr = rays ? rays + (g0 / K) * 8 : nullptr;
      const float* cz = z ? z + g0 : nullptr;
      float* co = out + g0 * 4;
      if (precision == PNR_FP32) {
        Arena a(ws, ws_bytes);
        float* zx = a.take<float>((size_t)sc.ns * n * (sc.d_latent + sc.d_in));
        size_t used = align_up(a.off, 256);
        if (used > ws_bytes) {
          set_err("net_eval: workspace too small");
          return PNR_ERR_WORKSPACE;
        }
        PNR_TRY(launch_point_features_f32(os, cx, cv, cr, cz, K, 1, (int)n, zx, st));
        PNR_TRY(mlp_forward_f32(m, zx, 1, sc.ns, (int)n, co, true, (char*)ws + used, ws_bytes - used, st));
      } else {
        PNR_TRY(net_forward_tc(os, m, cx, cv, cr, cz, K, 1, n, co, ws, ws_bytes, st));
      }
    }
  }
  return PNR_OK;
}

}  // namespace pnr

using namespace pnr;

extern "C" {

int pnr_abi_version(void) { return 2; }
const char* pnr_last_error(void) { return err_buf(); }
int64_t pnr_launch_count(int reset) {
  int64_t v = launch_counter();
  if (reset) launch_counter() = 0;
  return v;
}

int pnr_profile_begin(void) {
  for (ProfRec* r : prof().recs) {
    cudaEventDestroy(r->a);
    cudaEventDestroy(r->b);
    delete r;
  }
  prof().recs.clear();
  prof().on = true;
  return PNR_OK;
}

int pnr_profile_end(double* ms, int64_t* launches, double* flops, double* bytes) {
  prof().on = false;
  for (int k = 0; k < PROF_KINDS; ++k) {
    ms[k] = 0;
    launches[k] = 0;
    flops[k] = 0;
    bytes[k] = 0;
  }
  PNR_CUDA(cudaDeviceSynchronize());
  for (ProfRec* r : prof().recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r->a, r->b);
    ms[r->kind] += t;
    launches[r->kind] += 1;
    flops[r->kind] += r->flops;
    bytes[r->kind] += r->bytes;
    cudaEventDestroy(r->a);
    cudaEventDestroy(r->b);
    delete r;
  }
  prof().recs.clear();
  return PNR_OK;
}

// debug: device buffer of [SMs/2][16] uint64 cycle counters filled by the next fused-MLP launches of this
// thread (NULL = off)
int pnr_tc_debug_stats(void* device_buffer) {
  tc_set_stats((unsigned long long*)device_buffer);
  return PNR_OK;
}

int pnr_tc_check(pnr_stream stream) { return tc_check((cudaStream_t)stream); }

int pnr_pack_level(const float* src, int n_views, int C, int H, int W, void* dst, int dst_dtype, pnr_stream stream) {
  PNR_CHECK_ARG(src && dst, "pack_level: NULL pointer");
  PNR_CHECK_ARG(n_views > 0 && C > 0 && H > 0 && W > 0, "pack_level: bad shape");
  PNR_CHECK_ARG(dst_dtype == PNR_FP32 || dst_dtype == PNR_BF16 || dst_dtype == PNR_FP16, "pack_level: bad dtype");
  PNR_CHECK_ARG((long long)n_views * H <= 65535LL, "pack_level: n_views * H = %lld exceeds 65535 (grid z limit)",
                (long long)n_views * H);
  return launch_pack_level(src, n_views, C, H, W, dst, dst_dtype, (cudaStream_t)stream);
}

size_t pnr_mlp_packed_bytes(const pnr_mlp* mlp) { return mlp ? mlp_tc_packed_bytes(*mlp) : 0; }

int pnr_mlp_pack(const pnr_mlp* mlp, void* dst, size_t dst_bytes, int dtype, pnr_stream stream) {
  PNR_TRY(validate_mlp(mlp, nullptr, PNR_FP32));
  PNR_CHECK_ARG(dst != nullptr, "mlp_pack: dst is NULL");
  PNR_CHECK_ARG(dtype == PNR_BF16 || dtype == PNR_FP16, "mlp_pack: dtype must be PNR_BF16 or PNR_FP16");
  return mlp_tc_pack(*mlp, dst, dst_bytes, dtype == PNR_FP16 ? 1 : 0, (cudaStream_t)stream);
}
int pnr_mlp_pack_bf16(const pnr_mlp* mlp, void* dst, size_t dst_bytes, pnr_stream stream) {
  return pnr_mlp_pack(mlp, dst, dst_bytes, PNR_BF16, stream);
}

int pnr_point_features_f32(const pnr_scene* scene, const float* xyz, const float* viewdirs, int SB, int P, float* zx,
                           pnr_stream stream) {
  PNR_TRY(validate_scene(scene));
  PNR_CHECK_ARG(xyz && zx, "point_features: NULL pointer");
  PNR_CHECK_ARG(!scene->use_viewdirs || viewdirs, "point_features: viewdirs required (use_viewdirs)");
  PNR_CHECK_ARG(SB * scene->ns == scene->n_views, "point_features: SB*NS != n_views");
  return launch_point_features_f32(*scene, xyz, viewdirs, nullptr, nullptr, 0, SB, P, zx, (cudaStream_t)stream);
}

size_t pnr_net_forward_workspace(const pnr_scene* scene, const pnr_mlp* mlp, int SB, int P, int precision) {
  if (!scene || !mlp) return 0;
  long long Pc = chunk_points(*scene, precision, 0, P);
  return net_chunk_workspace(*scene, *mlp, precision, Pc) + 1024;
}

int pnr_net_forward(const pnr_scene* scene, const pnr_mlp* mlp, const float* xyz, const float* viewdirs, int SB,
                    int P, int precision, float* out, void* workspace, size_t workspace_bytes, pnr_stream stream) {
  PNR_TRY(validate_scene(scene));
  PNR_TRY(validate_mlp(mlp, scene, precision));
  PNR_CHECK_ARG(xyz && out, "net_forward: NULL pointer");
  PNR_CHECK_ARG(((uintptr_t)out & 15) == 0, "net_forward: out must be 16-byte aligned");
  PNR_CHECK_ARG(!scene->use_viewdirs || viewdirs, "net_forward: viewdirs required (use_viewdirs)");
  PNR_CHECK_ARG(SB * scene->ns == scene->n_views, "net_forward: SB*NS != n_views");
  if (SB == 0 || P == 0) return PNR_OK;
  return net_eval(*scene, *mlp, precision, xyz, viewdirs, nullptr, nullptr, 0, SB, P, out, workspace, workspace_bytes,
                  (cudaStream_t)stream);
}

int pnr_mlp_forward(const pnr_mlp* mlp, const float* zx, int SB, int NS, int P, int precision, float* out,
                    void* workspace, size_t workspace_bytes, pnr_stream stream) {
  PNR_TRY(validate_mlp(mlp, nullptr, precision));
  PNR_CHECK_ARG(zx && out, "mlp_forward: NULL pointer");
  PNR_CHECK_ARG(((uintptr_t)out & 15) == 0, "mlp_forward: out must be 16-byte aligned");
  PNR_UNSUPPORTED(NS > 1 && mlp->combine_layer >= mlp->n_blocks, "multi-view rows need combine_layer < n_blocks");
  if (precision == PNR_FP32)
    return mlp_forward_f32(*mlp, zx, SB, NS, P, out, false, workspace, workspace_bytes, (cudaStream_t)stream);
  return mlp_forward_tc_rows(*mlp, zx, SB, NS, P, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t pnr_mlp_forward_workspace(const pnr_mlp* mlp, int SB, int NS, int P, int precision) {
  if (!mlp) return 0;
  if (precision == PNR_FP32) return mlp_f32_workspace(*mlp, (long long)SB * NS * P, (long long)SB * P) + 1024;
  return mlp_tc_rows_workspace(*mlp, SB, NS, P) + 1024;
}

int pnr_gen_rays(const float* poses_c2w, int N, int W, int H, float fx, float fy, float cx, float cy, float z_near,
                 float z_far, float* rays, pnr_stream stream) {
  PNR_CHECK_ARG(poses_c2w && rays, "gen_rays: NULL pointer");
  PNR_CHECK_ARG(N >= 0 && W > 0 && H > 0 && fx != 0.f && fy != 0.f, "gen_rays: bad camera");
  PNR_CHECK_ARG(((uintptr_t)rays & 15) == 0, "gen_rays: rays must be 16-byte aligned");
  return launch_gen_rays(poses_c2w, N, W, H, fx, fy, cx, cy, z_near, z_far, rays, (cudaStream_t)stream);
}

int pnr_finalize_rgb(const float* rgb, const float* gt, int64_t n, uint8_t* u8, double* sse, pnr_stream stream) {
  PNR_CHECK_ARG(rgb != nullptr && n >= 0, "finalize_rgb: bad arguments");
  PNR_CHECK_ARG(!(gt && !sse), "finalize_rgb: gt given without an sse accumulator");
  return launch_finalize_rgb(rgb, gt, n, u8, sse, (cudaStream_t)stream);
}

int pnr_frame_metrics(const float* rgb, const float* gt, int NV, int H, int W, int C, int win, float data_range,
                      double* sums, pnr_stream stream) {
  PNR_CHECK_ARG(rgb && gt && sums, "frame_metrics: NULL pointer");
  PNR_CHECK_ARG(NV >= 0 && H >= 1 && W >= 1 && C >= 1, "frame_metrics: bad image shape (%d,%d,%d,%d)", NV, H, W, C);
  PNR_CHECK_ARG(win >= 3 && (win & 1) == 1 && win <= H && win <= W, "frame_metrics: win=%d must be odd, >= 3 and fit the image", win);
  PNR_CHECK_ARG(data_range > 0.f, "frame_metrics: data_range must be positive");
  return launch_frame_metrics(rgb, gt, NV, H, W, C, win, data_range, sums, (cudaStream_t)stream);
}

int pnr_sample_coarse(const float* rays, const float* jitter, int B, int Kc, int lindisp, float* z, pnr_stream stream) {
  PNR_CHECK_ARG(rays && jitter && z, "sample_coarse: NULL pointer");
  PNR_CHECK_ARG(Kc >= 1, "sample_coarse: Kc must be >= 1");
  return launch_sample_coarse(rays, jitter, B, Kc, lindisp, z, (cudaStream_t)stream);
}

int pnr_composite(const float* rays, const float* z, const float* rgb_sigma, int B, int K, int white_bkgd,
                  float* weights, float* rgb, float* depth, pnr_stream stream) {
  PNR_CHECK_ARG(rays && z && rgb_sigma && rgb && depth, "composite: NULL pointer");
  PNR_CHECK_ARG(((uintptr_t)rgb_sigma & 15) == 0, "composite: rgb_sigma must be 16-byte aligned");
  return launch_composite(rays, z, rgb_sigma, B, K, white_bkgd, weights, rgb, depth, (cudaStream_t)stream);
}

int pnr_fine_indices(const float* cdf, const float* u, int B, int Kc, int Kf, float* inds, pnr_stream stream) {
  PNR_CHECK_ARG(cdf && u && inds, "fine_indices: NULL pointer");
  return launch_fine_indices(cdf, u, B, Kc, Kf, inds, (cudaStream_t)stream);
}

int pnr_sample_fine_sorted(const float* rays, const float* z_coarse, const float* weights, const float* depth,
                           const float* fine_u, const float* fine_jitter, const float* depth_normal, int B, int Kc,
                           int n_fine, int n_fine_depth, float depth_std, int lindisp, float* z_out,
                           pnr_stream stream) {
  PNR_CHECK_ARG(rays && z_coarse && z_out, "sample_fine: NULL pointer");
  PNR_CHECK_ARG(n_fine - n_fine_depth <= 0 || (weights && fine_u && fine_jitter), "sample_fine: importance inputs missing");
  PNR_CHECK_ARG(n_fine_depth <= 0 || (depth && depth_normal), "sample_fine: depth inputs missing");
  return launch_sample_fine_sorted(rays, z_coarse, weights, depth, fine_u, fine_jitter, depth_normal, B, Kc, n_fine,
                                   n_fine_depth, depth_std, lindisp, z_out, (cudaStream_t)stream);
}

// ---- whole path ---------------------------------------------------------------------------------
struct RenderPlan {
  float *z_c, *out_c, *w_c, *rgb_c, *d_c, *z_f, *out_f, *w_f;
  void* net_ws;
  size_t net_ws_bytes;
  size_t total;
};

static RenderPlan plan_render(const pnr_scene& sc, const pnr_mlp& mc, const pnr_mlp* mf, const pnr_render_cfg& cfg,
                              int SB, int B, void* ws, size_t ws_bytes) {
  RenderPlan p;
  memset(&p, 0, sizeof(p));
  Arena a(ws, ws_bytes);
  long long R = (long long)SB * B;
  int Kc = cfg.n_coarse, K = cfg.n_coarse + cfg.n_fine;
  p.z_c = a.take<float>((size_t)R * Kc);
  p.out_c = a.take<float>((size_t)R * Kc * 4);
  p.w_c = a.take<float>((size_t)R * Kc);
  p.rgb_c = a.take<float>((size_t)R * 3);
  p.d_c = a.take<float>((size_t)R);
  if (cfg.n_fine > 0) {
    p.z_f = a.take<float>((size_t)R * K);
    p.out_f = a.take<float>((size_t)R * K * 4);
    p.w_f = a.take<float>((size_t)R * K);
  }
  long long Pc_c = chunk_points(sc, cfg.precision, Kc, (long long)B * Kc);
  size_t nw = net_chunk_workspace(sc, mc, cfg.precision, Pc_c);
  if (cfg.n_fine > 0) {
    long long Pc_f = chunk_points(sc, cfg.precision, K, (long long)B * K);
    size_t nf = net_chunk_workspace(sc, mf ? *mf : mc, cfg.precision, Pc_f);
    if (nf > nw) nw = nf;
  }
  p.net_ws = a.take<char>(nw);
  p.net_ws_bytes = nw;
  p.total = a.off + 256;
  return p;
}

size_t pnr_render_workspace(const pnr_scene* scene, const pnr_mlp* coarse, const pnr_mlp* fine,
                            const pnr_render_cfg* cfg, int SB, int B) {
  if (!scene || !coarse || !cfg) return 0;
  return plan_render(*scene, *coarse, fine, *cfg, SB, B, nullptr, 0).total;
}

int pnr_render_rays(const pnr_scene* scene, const pnr_mlp* coarse, const pnr_mlp* fine, const pnr_render_cfg* cfg,
                    const float* rays, int SB, int B, const pnr_rng_tape* tape, const pnr_render_out* out,
                    void* workspace, size_t workspace_bytes, pnr_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PNR_CHECK_ARG(cfg && rays && tape && out, "render_rays: NULL argument");
  PNR_TRY(validate_scene(scene));
  PNR_TRY(validate_mlp(coarse, scene, cfg->precision));
  if (fine) PNR_TRY(validate_mlp(fine, scene, cfg->precision));
  PNR_CHECK_ARG(SB * scene->ns == scene->n_views, "render_rays: SB*NS != n_views");
  PNR_CHECK_ARG(cfg->n_coarse >= 1 && cfg->n_fine >= 0 && cfg->n_fine_depth >= 0 && cfg->n_fine_depth <= cfg->n_fine,
                "render_rays: bad sample counts c=%d f=%d fd=%d", cfg->n_coarse, cfg->n_fine, cfg->n_fine_depth);
  PNR_CHECK_ARG(tape->coarse_jitter, "render_rays: tape.coarse_jitter missing");
  PNR_CHECK_ARG(out->rgb_coarse && out->depth_coarse, "render_rays: coarse outputs missing");
  if (SB == 0 || B == 0) return PNR_OK;
  RenderPlan p = plan_render(*scene, *coarse, fine, *cfg, SB, B, workspace, workspace_bytes);
  if (p.total > workspace_bytes + 256 || !workspace) {
    set_err("render_rays: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    return PNR_ERR_WORKSPACE;
  }
  NvtxRange nvtx("renderer_forward");
  const int R = SB * B, Kc = cfg->n_coarse, K = cfg->n_coarse + cfg->n_fine;
  float* z_c = out->z_coarse ? out->z_coarse : p.z_c;
  float* w_c = out->weights_coarse ? out->weights_coarse : p.w_c;
  PNR_TRY(launch_sample_coarse(rays, tape->coarse_jitter, R, Kc, cfg->lindisp, z_c, st));
  {
    NvtxRange nc("renderer_composite");
    PNR_TRY(net_eval(*scene, *coarse, cfg->precision, nullptr, nullptr, rays, z_c, Kc, SB, (long long)B * Kc, p.out_c,
                     p.net_ws, p.net_ws_bytes, st));
    PNR_TRY(launch_composite(rays, z_c, p.out_c, R, Kc, cfg->white_bkgd, w_c, out->rgb_coarse, out->depth_coarse, st));
  }
  if (cfg->n_fine > 0) {
    PNR_CHECK_ARG(out->rgb_fine && out->depth_fine, "render_rays: fine outputs missing");
    int n_imp = cfg->n_fine - cfg->n_fine_depth;
    PNR_CHECK_ARG(n_imp == 0 || (tape->fine_u && tape->fine_jitter), "render_rays: tape.fine_u/fine_jitter missing");
    PNR_CHECK_ARG(cfg->n_fine_depth == 0 || tape->depth_normal, "render_rays: tape.depth_normal missing");
    float* z_f = out->z_fine ? out->z_fine : p.z_f;
    float* w_f = out->weights_fine ? out->weights_fine : p.w_f;
    PNR_TRY(launch_sample_fine_sorted(rays, z_c, w_c, out->depth_coarse, tape->fine_u, tape->fine_jitter,
                                      tape->depth_normal, R, Kc, cfg->n_fine, cfg->n_fine_depth, cfg->depth_std,
                                      cfg->lindisp, z_f, st));
    NvtxRange nc("renderer_composite");
    PNR_TRY(net_eval(*scene, fine ? *fine : *coarse, cfg->precision, nullptr, nullptr, rays, z_f, K, SB,
                     (long long)B * K, p.out_f, p.net_ws, p.net_ws_bytes, st));
    PNR_TRY(launch_composite(rays, z_f, p.out_f, R, K, cfg->white_bkgd, w_f, out->rgb_fine, out->depth_fine, st));
  }
  return PNR_OK;
}

}  // extern "C"
