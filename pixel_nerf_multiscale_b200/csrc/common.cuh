// Shared host/device helpers for the pixelnerf_b200 native library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pixelnerf_b200.h"

namespace pnr {

// ---- thread-local error + launch accounting ------------------------------------------------
char* err_buf();
void set_err(const char* fmt, ...);
int64_t& launch_counter();

#define PNR_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      pnr::set_err(__VA_ARGS__);            \
      return PNR_ERR_BAD_ARG;               \
    }                                       \
  } while (0)

#define PNR_UNSUPPORTED(cond, ...)          \
  do {                                      \
    if (cond) {                             \
      pnr::set_err(__VA_ARGS__);            \
      return PNR_ERR_UNSUPPORTED;           \
    }                                       \
  } while (0)

#define PNR_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      pnr::set_err("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return PNR_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors
#define PNR_LAUNCHED()                                                                     \
  do {                                                                                     \
    pnr::launch_counter()++;                                                               \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      pnr::set_err("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return PNR_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define PNR_TRY(expr)        \
  do {                       \
    int _s = (expr);         \
    if (_s != PNR_OK) return _s; \
  } while (0)

// optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the launching
// stream around the tracked kernels; thread-local, off by default.
enum ProfKind { PROF_FEATURES = 0, PROF_FUSED_MLP = 1, PROF_RESERVED = 2, PROF_KINDS = 3 };
#define PNR_MAX_DEVICES 64
struct ProfScope {
  int kind;
  cudaStream_t st;
  void* rec;
  ProfScope(int kind, double flops, double bytes, cudaStream_t st);
  ~ProfScope();
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap, off;
  bool dry;  // size-query mode: nothing is dereferenced
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0), dry(p == nullptr) {}
  template <typename T>
  T* take(size_t count) {
    size_t o = align_up(off, 256);
    off = o + count * sizeof(T);
    return dry ? (T*)nullptr : (T*)(base + o);
  }
  bool ok() const { return dry || off <= cap; }
};

// ---- internal cross-file entry points ---------------------------------------------------------
// features.cu
int launch_point_features_f32(const pnr_scene& sc, const float* xyz, const float* viewdirs,
                              const float* rays, const float* z, int K, int SB, int P, float* zx,
                              cudaStream_t st);
// mlp_f32.cu
size_t mlp_f32_workspace(const pnr_mlp& m, long long rows_pre, long long rows_post);
int mlp_forward_f32(const pnr_mlp& m, const float* zx, int SB, int NS, int P, float* out_raw,
                    bool apply_head, void* ws, size_t ws_bytes, cudaStream_t st);
// rays.cu
int launch_sample_coarse(const float* rays, const float* jitter, int B, int Kc, int lindisp, float* z,
                         cudaStream_t st);
int launch_gen_rays(const float* poses, int N, int W, int H, float fx, float fy, float cx, float cy, float near,
                    float far, float* rays, cudaStream_t st);
int launch_frame_metrics(const float* a, const float* b, int NV, int H, int W, int C, int win, float data_range,
                         double* sums, cudaStream_t st);
int launch_finalize_rgb(const float* rgb, const float* gt, long long n, uint8_t* u8, double* sse, cudaStream_t st);
int launch_composite(const float* rays, const float* z, const float* rgb_sigma, int B, int K, int white,
                     float* weights, float* rgb, float* depth, cudaStream_t st);
int launch_fine_indices(const float* cdf, const float* u, int B, int Kc, int Kf, float* inds, cudaStream_t st);
int launch_sample_fine_sorted(const float* rays, const float* z_coarse, const float* weights,
                              const float* depth, const float* fine_u, const float* fine_jitter,
                              const float* depth_normal, int B, int Kc, int n_fine, int n_fine_depth,
                              float depth_std, int lindisp, float* z_out, cudaStream_t st);
// mlp_tc.cu (tcgen05 path)
size_t mlp_tc_packed_bytes(const pnr_mlp& m);
int mlp_tc_pack(const pnr_mlp& m, void* dst, size_t dst_bytes, int fmt /* 0 bf16, 1 f16 */, cudaStream_t st);
size_t net_tc_workspace(const pnr_scene& sc, const pnr_mlp& m, int SB, long long P);
int net_forward_tc(const pnr_scene& sc, const pnr_mlp& m, const float* xyz, const float* viewdirs,
                   const float* rays, const float* z, int K, int SB, long long P, float* out, void* ws,
                   size_t ws_bytes, cudaStream_t st);

}  // namespace pnr
