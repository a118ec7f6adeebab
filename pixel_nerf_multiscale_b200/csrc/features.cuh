// Device-side pieces of kernel (a): camera transform, projection, bilinear pyramid gather and
// positional code for one (point, source view) pair.  Shared by the fp32 validation kernel
// (features.cu) and the bf16 operand producer of the tcgen05 path (mlp_tc.cu).
//
// Reference semantics (paths relative to the reference root):
//   transform / projection   src/model/models.py.backup2:170-174, 215-221
//   gather                   src/model/encoder.py:138-205 (grid_sample bilinear/border/align_corners)
//   positional code          src/model/code.py:30-47
#pragma once
#include "common.cuh"

namespace pnr {

struct PointCam {
  float xr[3];  // R * X            (xyz_rot, models.py.backup2:171-173)
  float xc[3];  // R * X + t        (xyz,     models.py.backup2:174)
  float vd[3];  // R * viewdir      (models.py.backup2:197-202)
  float u, v;   // pixel coordinates in the source view (:215-221)
};

__device__ __forceinline__ void load_point(const float* __restrict__ xyz, const float* __restrict__ viewdirs,
                                           const float* __restrict__ rays, const float* __restrict__ z, int K,
                                           long long gp /* sb*P + p */, float X[3], float D[3]) {
  if (rays != nullptr) {
    long long r = gp / K;
    const float* ry = rays + r * 8;
    float t = z[gp];
    // points = o + z * d (src/render/nerf.py:185); viewdirs = d (nerf.py:203-205)
    X[0] = ry[0] + t * ry[3];
    X[1] = ry[1] + t * ry[4];
    X[2] = ry[2] + t * ry[5];
    D[0] = ry[3];
    D[1] = ry[4];
    D[2] = ry[5];
  } else {
    X[0] = xyz[gp * 3 + 0];
    X[1] = xyz[gp * 3 + 1];
    X[2] = xyz[gp * 3 + 2];
    if (viewdirs != nullptr) {
      D[0] = viewdirs[gp * 3 + 0];
      D[1] = viewdirs[gp * 3 + 1];
      D[2] = viewdirs[gp * 3 + 2];
    } else {
      D[0] = D[1] = D[2] = 0.f;
    }
  }
}

__device__ __forceinline__ void camera_project(const float* __restrict__ cam /*16 floats*/, const float X[3],
                                               const float D[3], PointCam& pc) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float r0 = cam[i * 3 + 0], r1 = cam[i * 3 + 1], r2 = cam[i * 3 + 2];
    pc.xr[i] = r0 * X[0] + r1 * X[1] + r2 * X[2];
    pc.xc[i] = pc.xr[i] + cam[9 + i];
    pc.vd[i] = r0 * D[0] + r1 * D[1] + r2 * D[2];
  }
  // uv = -xy / z * (fx, -fy) + c   (fy arrives already negated)
  float un = -pc.xc[0] / pc.xc[2];
  float vn = -pc.xc[1] / pc.xc[2];
  pc.u = un * cam[12] + cam[14];
  pc.v = vn * cam[13] + cam[15];
}

struct Taps {
  int o00, o01, o10, o11;  // element offsets of the 4 texels (x fastest; NOT multiplied by C)
  float w00, w01, w10, w11;
};

// grid_sample(bilinear, border, align_corners=True) of pixel coordinate (u,v) on an HxW map after
// the reference's uv/(W-1)*2-1 normalisation (encoder.py:174-176); the normalise/unnormalise round
// trip is kept literally so fp32 rounding follows the reference.
__device__ __forceinline__ Taps make_taps(float u, float v, int H, int W, float kx, float ky) {
  float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float gx = ((u * kx) / wm1) * 2.f - 1.f;
  float gy = ((v * ky) / hm1) * 2.f - 1.f;
  float ix = ((gx + 1.f) / 2.f) * wm1;
  float iy = ((gy + 1.f) / 2.f) * hm1;
  ix = fminf(fmaxf(ix, 0.f), wm1);
  iy = fminf(fmaxf(iy, 0.f), hm1);
  float x0f = floorf(ix), y0f = floorf(iy);
  int x0 = (int)x0f, y0 = (int)y0f;
  // weights exactly as grid_sample forms them: (x0+1 - ix), (ix - x0)
  float fx1 = ix - x0f, fy1 = iy - y0f;
  float fx0 = (x0f + 1.f) - ix, fy0 = (y0f + 1.f) - iy;
  // +1 taps that fall outside contribute zero (their weight is zero as well): clamp the
  // address and zero the weight.
  int x1 = x0 + 1, y1 = y0 + 1;
  bool xin = x1 <= W - 1, yin = y1 <= H - 1;
  if (!xin) x1 = W - 1;
  if (!yin) y1 = H - 1;
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

// j-th entry (0 <= j < d_in) of the code part of an MLP input row.
__device__ __forceinline__ float code_entry(const pnr_scene& sc, const PointCam& pc, int j) {
  // base vector b: xyz feature (3 or 1) [+ viewdirs when they are encoded too]
  const int dz = sc.use_xyz ? 3 : 1;
  const bool vd_in_code = sc.use_viewdirs && sc.use_code && sc.use_code_viewdirs;
  const int db = dz + (vd_in_code ? 3 : 0);
  auto base = [&](int i) -> float {
    if (i < dz) {
      if (sc.use_xyz) return sc.normalize_z ? pc.xr[i] : pc.xc[i];
      return sc.normalize_z ? -pc.xr[2] : -pc.xc[2];
    }
    return pc.vd[i - dz];
  };
  int coded = sc.use_code ? (sc.num_freqs * 2 * db + (sc.include_input ? db : 0)) : db;
  if (j >= coded) return pc.vd[j - coded];  // raw viewdirs appended after the code (:203-205)
  if (!sc.use_code) return base(j);
  if (sc.include_input) {
    if (j < db) return base(j);
    j -= db;
  }
  int g = j / db, i = j - g * db;
  float freq = ldexpf(sc.freq_factor, g >> 1);
  float phase = (g & 1) ? 1.57079637050628662109375f : 0.f;  // fl32(pi/2), code.py:26
  return sinf(fmaf(base(i), freq, phase));
}

}  // namespace pnr
