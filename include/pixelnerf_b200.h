/*
 * pixelnerf_b200.h -- C ABI of the B200-native pixelNeRF ray-rendering hot path.
 *
 * The reference (Zxhh123/pixel-nerf-multiscale) is pure Python/PyTorch and has no FFI boundary
 * (SURVEY.md F1, section 8b); this header DEFINES the boundary underneath the reference's Python API.
 * Each entry point names the reference function it replaces (paths relative to the
 * reference root).  Conventions:
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (PyTorch in the shipped host code) unless stated otherwise;
 *   - every function enqueues on the caller's stream and never synchronises the device;
 *   - every function returns 0 on success or a negative pnr_status; pnr_last_error() gives a
 *     thread-local message.  Nothing throws, nothing aborts;
 *   - calls on different devices / streams / threads are independent.  Process-wide state is limited to
 *     (a) one lazily created record PER DEVICE (SM count, "kernel attribute set" flags and the
 *     barrier-fault word described at pnr_tc_check), guarded by a mutex, and (b) read-once
 *     environment knobs (PNR_WAIT_TIMEOUT_MS, PNR_MAX_PAIRS, PNR_SOLO_MMA, PNR_CHUNK_ROWS_LOG2);
 *     error message, launch counter, profile records and tensor-map cache are thread-local;
 *   - there is NO CPU implementation behind this ABI.
 */
#ifndef PIXELNERF_B200_H
#define PIXELNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pnr_stream; /* cudaStream_t */

enum pnr_status {
  PNR_OK = 0,
  PNR_ERR_BAD_ARG = -1,      /* shape / pointer / alignment violation            */
  PNR_ERR_UNSUPPORTED = -2,  /* configuration outside the native path            */
  PNR_ERR_CUDA = -3,         /* CUDA runtime error (message in pnr_last_error)   */
  PNR_ERR_WORKSPACE = -4     /* caller-provided workspace too small              */
};

enum pnr_precision {
  PNR_FP32 = 0, /* validation mode: fp32 SIMT FMA everywhere (<=1e-4 vs reference) */
  PNR_BF16 = 1, /* production mode: bf16 operands on tcgen05, fp32 accumulate      */
  PNR_FP16 = 2  /* same kernels and MMA rate with f16 operands (11 instead of 8
                   significand bits; values saturate at 65504), fp32 accumulate   */
};

#define PNR_MAX_LEVELS 8
#define PNR_MAX_BLOCKS 8

/*
 * Source-view state left behind by PixelNeRFNet.encode (src/model/models.py.backup2:98-153)
 * plus the static feature/code hyper-parameters (src/model/models.py.backup2:21-66).
 * Feature maps are the encoder's per-level outputs (src/model/encoder.py:117-136) re-laid
 * out channels-last by pnr_pack_level().
 */
typedef struct pnr_scene {
  int32_t n_views;   /* SB*NS rows of cams / images                                   */
  int32_t ns;        /* views per object (num_views_per_obj)                          */
  int32_t n_levels;  /* 1 (single-scale) .. PNR_MAX_LEVELS (use_multi_scale)          */
  int32_t d_latent;  /* sum of C[i]                                                   */
  int32_t feat_dtype; /* pnr_precision of the packed maps: fp32, bf16 or f16, NHWC    */
  int32_t C[PNR_MAX_LEVELS], H[PNR_MAX_LEVELS], W[PNR_MAX_LEVELS];
  int32_t ch_off[PNR_MAX_LEVELS]; /* first latent column of level i                   */
  float kx[PNR_MAX_LEVELS], ky[PNR_MAX_LEVELS]; /* pixel-uv -> texel scale; 1 = the
                                  fork's behaviour (encoder.py:162-176)   */
  const void* level[PNR_MAX_LEVELS]; /* [n_views][H][W][C]                            */
  const float* cams; /* [n_views][16] = R(9, row-major world->cam) t(3) fx fy cx cy,
                        fy already negated (models.py.backup2:139)                    */
  /* z_feature / positional code (models.py.backup2:176-209, src/model/code.py:30-47) */
  int32_t use_xyz, normalize_z, use_viewdirs, use_code, use_code_viewdirs;
  int32_t num_freqs, include_input;
  float freq_factor;
  int32_t d_in;      /* width of the code part of an MLP input row                    */
} pnr_scene;

/*
 * One ResnetFC (src/model/resnetfc.py:65-236): lin_in, lin_z[0..n_lin_z), n_blocks x
 * ResnetBlockFC(fc_0, fc_1), lin_out.  fp32 pointers are nn.Linear tensors as-is
 * (weight (out,in) row-major, bias (out)).  `packed` is the 16-bit operand image produced by
 * pnr_mlp_pack() (NULL in fp32 mode) and `packed_dtype` the pnr_precision it was packed for.
 */
typedef struct pnr_mlp {
  int32_t d_in, d_latent, d_hidden, d_out, n_blocks, combine_layer, n_lin_z;
  int32_t combine_type; /* 0 = average (only supported value; resnetfc.py:246)        */
  const float *lin_in_w, *lin_in_b, *lin_out_w, *lin_out_b;
  const float *lin_z_w[PNR_MAX_BLOCKS], *lin_z_b[PNR_MAX_BLOCKS];
  const float *fc0_w[PNR_MAX_BLOCKS], *fc0_b[PNR_MAX_BLOCKS];
  const float *fc1_w[PNR_MAX_BLOCKS], *fc1_b[PNR_MAX_BLOCKS];
  const void* packed;
  size_t packed_bytes;
  int32_t packed_dtype; /* PNR_BF16 | PNR_FP16 */
  int32_t reserved;
} pnr_mlp;

/* NeRFRenderer state (src/render/nerf.py:62-96). */
typedef struct pnr_render_cfg {
  int32_t n_coarse, n_fine, n_fine_depth; /* n_fine INCLUDES n_fine_depth (nerf.py:286-293) */
  int32_t white_bkgd, lindisp;
  float depth_std;
  int32_t precision;   /* pnr_precision */
  int32_t want_weights;
} pnr_render_cfg;

/* Random draws of one render call, in reference order/shape (nerf.py:111,135-141,158).
 * All (B, .) fp32 row-major, drawn by the caller (torch) so that a seeded run replays the
 * reference's own randoms. */
typedef struct pnr_rng_tape {
  const float* coarse_jitter; /* (B, n_coarse)              U[0,1)  */
  const float* fine_u;        /* (B, n_fine-n_fine_depth)   U[0,1)  */
  const float* fine_jitter;   /* (B, n_fine-n_fine_depth)   U[0,1)  */
  const float* depth_normal;  /* (B, n_fine_depth)          N(0,1)  */
} pnr_rng_tape;

typedef struct pnr_render_out {
  float *rgb_coarse, *depth_coarse, *weights_coarse; /* (B,3) (B) (B,Kc)        ; weights may be NULL */
  float *rgb_fine, *depth_fine, *weights_fine;       /* (B,3) (B) (B,Kc+Kf)     ; NULL when n_fine==0 */
  float *z_coarse, *z_fine;                          /* optional taps (B,Kc) (B,Kc+Kf), may be NULL   */
} pnr_render_out;

int pnr_abi_version(void);
const char* pnr_last_error(void);
/* Number of kernels this library has launched in the calling thread since the last reset
 * (bench.py's gpu_launches). */
int64_t pnr_launch_count(int reset);

/* Synchronises `stream` and reports (then clears) a pipeline fault recorded by the tensor-core
 * kernels on the CURRENT device: their mbarrier waits are wall-clock bounded (PNR_WAIT_TIMEOUT_MS,
 * default 2000, 0 = unbounded for debuggers / sanitizers), so a protocol error surfaces as
 * PNR_ERR_CUDA instead of hanging the GPU.  A fault is also reported, without synchronising, by
 * the next pnr_net_forward / pnr_mlp_forward / pnr_render_rays call on that device. */
int pnr_tc_check(pnr_stream stream);
/* Debug: device buffer of [SM count / 2][16] uint64 per-role cycle counters written by the fused-MLP
 * launches of the calling thread from now on (NULL = off; tools/tc_stats.py). */
int pnr_tc_debug_stats(void* device_buffer);

/* Optional per-kernel timing for the roofline report: between begin/end every tracked kernel is
 * bracketed by CUDA events on its launching stream.  Arrays have 3 entries:
 * [0] reserved (the gather runs inside [1]), [1] fused gather + ResnetFC kernel, [2] reserved.
 * flops/bytes are the ALGORITHMIC work of the tracked launches. */
int pnr_profile_begin(void);
int pnr_profile_end(double* ms, int64_t* launches, double* flops, double* bytes);

/* ---- packing ------------------------------------------------------------------------- */
/* NCHW fp32 feature map (encoder output, encoder.py:117-136) -> NHWC fp32|bf16|f16.      */
int pnr_pack_level(const float* src_nchw, int n_views, int C, int H, int W, void* dst_nhwc,
                   int dst_dtype, pnr_stream stream);
/* 16-bit operand image of one ResnetFC for the tcgen05 path (dtype PNR_BF16 | PNR_FP16; the caller
 * then sets mlp.packed / packed_bytes / packed_dtype).  pnr_mlp_pack_bf16 = dtype PNR_BF16.        */
size_t pnr_mlp_packed_bytes(const pnr_mlp* mlp);
int pnr_mlp_pack(const pnr_mlp* mlp, void* dst, size_t dst_bytes, int dtype, pnr_stream stream);
int pnr_mlp_pack_bf16(const pnr_mlp* mlp, void* dst, size_t dst_bytes, pnr_stream stream);

/* ---- kernel (a): per-(point,view) features ------------------------------------------- */
/* Rows of the ResnetFC input, fp32, reference order and layout:
 * zx[(sb*NS+ns)*P + p][d_latent + d_in] = [latent | code]   (models.py.backup2:166-243).
 * xyz (SB,P,3), viewdirs (SB,P,3) or NULL.                                               */
int pnr_point_features_f32(const pnr_scene* scene, const float* xyz, const float* viewdirs,
                           int SB, int P, float* zx, pnr_stream stream);

/* ---- kernels (a)+(b): PixelNeRFNet.forward (models.py.backup2:155-282) ---------------- */
size_t pnr_net_forward_workspace(const pnr_scene* scene, const pnr_mlp* mlp, int SB, int P,
                                 int precision);
/* out (SB,P,4) = [sigmoid(rgb), relu(sigma)].                                            */
int pnr_net_forward(const pnr_scene* scene, const pnr_mlp* mlp, const float* xyz,
                    const float* viewdirs, int SB, int P, int precision, float* out,
                    void* workspace, size_t workspace_bytes, pnr_stream stream);
/* ResnetFC.forward alone on caller-provided rows (resnetfc.py:173-236), fp32 rows in
 * reference order; NS/P are combine_inner_dims.  out (SB*P, d_out) raw (no sigmoid/relu). */
size_t pnr_mlp_forward_workspace(const pnr_mlp* mlp, int SB, int NS, int P, int precision);
int pnr_mlp_forward(const pnr_mlp* mlp, const float* zx, int SB, int NS, int P, int precision,
                    float* out, void* workspace, size_t workspace_bytes, pnr_stream stream);

/* ---- ray generation ("next" row 1) ------------------------------------------ */
/* util.gen_rays + unproj_map (src/util/util.py:118-148,243-281), ndc=False: N camera-to-world poses
 * (N,4,4) row-major -> rays (N,H,W,8) = [origin, unit direction, near, far], generated on the device. */
int pnr_gen_rays(const float* poses_c2w, int N, int W, int H, float fx, float fy, float cx, float cy,
                 float z_near, float z_far, float* rays, pnr_stream stream);

/* ---- output side ("next" row 3) ---------------------------------------------- */
/* n floats of rendered rgb: v = clamp(rgb,0,1); u8 (nullable) = trunc(v*255) (eval/gen_video.py:226);
 * *sse (nullable, fp64, caller-zeroed) += sum (v-gt)^2 for PSNR = -10 log10(sse/n) (eval/eval.py:278-300,
 * src/util/util.py:479-486) without the per-batch .cpu() sync of the reference. */
int pnr_finalize_rgb(const float* rgb, const float* gt, int64_t n, uint8_t* u8, double* sse, pnr_stream stream);

/* Per-view metrics of the eval driver (eval/eval.py:314-343) on (NV,H,W,C) fp32 images, rgb clamped to [0,1]
 * first (eval.py:290-292): sums[2*v] += SSIM map summed over the cropped interior and channels as
 * skimage.measure.compare_ssim(multichannel=True, data_range) defines it (third-party, absent from the
 * reference tree: scikit-image 0.16 structural_similarity, uniform win x win window, sample covariance,
 * K1 = 0.01, K2 = 0.03) -- divide by C*(H-win+1)*(W-win+1); sums[2*v+1] += sum of squared error for
 * compare_psnr = 10 log10(R^2 / mse) -- divide by H*W*C.  sums is fp64, caller-zeroed; no host sync. */
int pnr_frame_metrics(const float* rgb, const float* gt, int NV, int H, int W, int C, int win, float data_range,
                      double* sums, pnr_stream stream);

/* ---- kernel (c): per-ray sampling / compositing (src/render/nerf.py) ------------------ */
/* sample_coarse, nerf.py:98-118.  rays (B,8), jitter (B,Kc) -> z (B,Kc).                 */
int pnr_sample_coarse(const float* rays, const float* jitter, int B, int Kc, int lindisp,
                      float* z, pnr_stream stream);
/* composite after the model call, nerf.py:178-182,223-244.  sigma_rgb (B,K,4).           */
int pnr_composite(const float* rays, const float* z, const float* rgb_sigma, int B, int K,
                  int white_bkgd, float* weights /*nullable*/, float* rgb, float* depth,
                  pnr_stream stream);
/* inverse-CDF bin lookup on a GIVEN cdf, nerf.py:138-139: inds = max(0, #{cdf<=u} - 1).   */
int pnr_fine_indices(const float* cdf, const float* u, int B, int Kc, int Kf, float* inds,
                     pnr_stream stream);
/* sample_fine + sample_fine_depth + cat + sort, nerf.py:120-161,286-295.
 * z_out (B, Kc+Kf) ascending, Kf = n_fine (importance + depth samples).                  */
int pnr_sample_fine_sorted(const float* rays, const float* z_coarse, const float* weights,
                           const float* depth, const float* fine_u, const float* fine_jitter,
                           const float* depth_normal, int B, int Kc, int n_fine,
                           int n_fine_depth, float depth_std, int lindisp, float* z_out,
                           pnr_stream stream);

/* ---- whole path: NeRFRenderer.forward (nerf.py:251-303) on one ray batch -------------- */
/* rays (SB,B,8) flattened; one object per SB entry, views sb*NS.. in scene.              */
size_t pnr_render_workspace(const pnr_scene* scene, const pnr_mlp* coarse, const pnr_mlp* fine,
                            const pnr_render_cfg* cfg, int SB, int B);
int pnr_render_rays(const pnr_scene* scene, const pnr_mlp* coarse, const pnr_mlp* fine /*nullable*/,
                    const pnr_render_cfg* cfg, const float* rays, int SB, int B,
                    const pnr_rng_tape* tape, const pnr_render_out* out, void* workspace,
                    size_t workspace_bytes, pnr_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* PIXELNERF_B200_H */
