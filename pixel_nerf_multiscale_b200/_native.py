"""
ctypes binding of include/pixelnerf_b200.h (libpixelnerf_b200.so, built in-tree by
csrc/build.sh for sm_100a).  There is no fallback: if the library is missing or a call fails,
the caller gets a RuntimeError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PIXELNERF_B200_LIB: alternative build of the same library (A/B timing of kernel variants on one box).
LIB_PATH = os.environ.get("PIXELNERF_B200_LIB") or os.path.join(_HERE, "libpixelnerf_b200.so")

FP32, BF16, FP16 = 0, 1, 2
MAX_LEVELS, MAX_BLOCKS = 8, 8

_fp = C.c_void_p


class Scene(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("ns", C.c_int32), ("n_levels", C.c_int32), ("d_latent", C.c_int32),
        ("feat_dtype", C.c_int32),
        ("C", C.c_int32 * MAX_LEVELS), ("H", C.c_int32 * MAX_LEVELS), ("W", C.c_int32 * MAX_LEVELS),
        ("ch_off", C.c_int32 * MAX_LEVELS),
        ("kx", C.c_float * MAX_LEVELS), ("ky", C.c_float * MAX_LEVELS),
        ("level", _fp * MAX_LEVELS),
        ("cams", _fp),
        ("use_xyz", C.c_int32), ("normalize_z", C.c_int32), ("use_viewdirs", C.c_int32), ("use_code", C.c_int32),
        ("use_code_viewdirs", C.c_int32), ("num_freqs", C.c_int32), ("include_input", C.c_int32),
        ("freq_factor", C.c_float), ("d_in", C.c_int32),
    ]


class Mlp(C.Structure):
    _fields_ = [
        ("d_in", C.c_int32), ("d_latent", C.c_int32), ("d_hidden", C.c_int32), ("d_out", C.c_int32),
        ("n_blocks", C.c_int32), ("combine_layer", C.c_int32), ("n_lin_z", C.c_int32), ("combine_type", C.c_int32),
        ("lin_in_w", _fp), ("lin_in_b", _fp), ("lin_out_w", _fp), ("lin_out_b", _fp),
        ("lin_z_w", _fp * MAX_BLOCKS), ("lin_z_b", _fp * MAX_BLOCKS),
        ("fc0_w", _fp * MAX_BLOCKS), ("fc0_b", _fp * MAX_BLOCKS),
        ("fc1_w", _fp * MAX_BLOCKS), ("fc1_b", _fp * MAX_BLOCKS),
        ("packed", _fp), ("packed_bytes", C.c_size_t), ("packed_dtype", C.c_int32), ("reserved", C.c_int32),
    ]


class RenderCfg(C.Structure):
    _fields_ = [
        ("n_coarse", C.c_int32), ("n_fine", C.c_int32), ("n_fine_depth", C.c_int32), ("white_bkgd", C.c_int32),
        ("lindisp", C.c_int32), ("depth_std", C.c_float), ("precision", C.c_int32), ("want_weights", C.c_int32),
    ]


class RngTape(C.Structure):
    _fields_ = [("coarse_jitter", _fp), ("fine_u", _fp), ("fine_jitter", _fp), ("depth_normal", _fp)]


class RenderOut(C.Structure):
    _fields_ = [("rgb_coarse", _fp), ("depth_coarse", _fp), ("weights_coarse", _fp), ("rgb_fine", _fp),
                ("depth_fine", _fp), ("weights_fine", _fp), ("z_coarse", _fp), ("z_fine", _fp)]


# name -> (restype, argtypes); every symbol include/pixelnerf_b200.h declares
_i, _sz, _f = C.c_int, C.c_size_t, C.c_float
_PS, _PM, _PC, _PT, _PO = C.POINTER(Scene), C.POINTER(Mlp), C.POINTER(RenderCfg), C.POINTER(RngTape), C.POINTER(RenderOut)
EXPORTS = {
    "pnr_abi_version": (_i, []),
    "pnr_last_error": (C.c_char_p, []),
    "pnr_launch_count": (C.c_int64, [_i]),
    "pnr_tc_check": (_i, [_fp]),
    "pnr_tc_debug_stats": (_i, [_fp]),
    "pnr_profile_begin": (_i, []),
    "pnr_profile_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pnr_pack_level": (_i, [_fp, _i, _i, _i, _i, _fp, _i, _fp]),
    "pnr_mlp_packed_bytes": (_sz, [_PM]),
    "pnr_mlp_pack": (_i, [_PM, _fp, _sz, _i, _fp]),
    "pnr_mlp_pack_bf16": (_i, [_PM, _fp, _sz, _fp]),
    "pnr_point_features_f32": (_i, [_PS, _fp, _fp, _i, _i, _fp, _fp]),
    "pnr_net_forward_workspace": (_sz, [_PS, _PM, _i, _i, _i]),
    "pnr_net_forward": (_i, [_PS, _PM, _fp, _fp, _i, _i, _i, _fp, _fp, _sz, _fp]),
    "pnr_mlp_forward_workspace": (_sz, [_PM, _i, _i, _i, _i]),
    "pnr_mlp_forward": (_i, [_PM, _fp, _i, _i, _i, _i, _fp, _fp, _sz, _fp]),
    "pnr_gen_rays": (_i, [_fp, _i, _i, _i, _f, _f, _f, _f, _f, _f, _fp, _fp]),
    "pnr_finalize_rgb": (_i, [_fp, _fp, C.c_int64, _fp, _fp, _fp]),
    "pnr_frame_metrics": (_i, [_fp, _fp, _i, _i, _i, _i, _i, _f, _fp, _fp]),
    "pnr_sample_coarse": (_i, [_fp, _fp, _i, _i, _i, _fp, _fp]),
    "pnr_composite": (_i, [_fp, _fp, _fp, _i, _i, _i, _fp, _fp, _fp, _fp]),
    "pnr_fine_indices": (_i, [_fp, _fp, _i, _i, _i, _fp, _fp]),
    "pnr_sample_fine_sorted": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _i, _fp, _fp]),
    "pnr_render_workspace": (_sz, [_PS, _PM, _PM, _PC, _i, _i]),
    "pnr_render_rays": (_i, [_PS, _PM, _PM, _PC, _fp, _i, _i, _PT, _PO, _fp, _sz, _fp]),
}

_lib = None
_probe = None
PROBE_LIB_PATH = os.path.join(_HERE, "libpixelnerf_b200_probe.so")


def probe_lib():
    """Hardware probes of the tcgen05 building blocks (csrc/tc_probe.cu): a separate library used by
    tests/test_gpu_tc_probe.py and tools/probe_*.py only -- not part of the product ABI."""
    global _probe
    if _probe is None:
        _probe = C.CDLL(PROBE_LIB_PATH)
    return _probe



def lib():
    """The loaded shared library.  Raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "pixelnerf_b200 native library not found at %s -- build it with "
                "pixel_nerf_multiscale_b200/csrc/build.sh (or __graft_entry__.build()); there is no CPU or "
                "PyTorch fallback for the ray-rendering path" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().pnr_last_error()
        msg = msg.decode("utf-8", "replace") if msg else ""
        if status == -1:
            # the reference signals shape problems with assert (resnetfc.py:190, nerf.py:269)
            raise AssertionError("%s: %s" % (what, msg))
        if status == -2:
            raise NotImplementedError("%s: %s" % (what, msg))
        raise RuntimeError("%s failed (status %d): %s" % (what, status, msg))


def ptr(t):
    """Device pointer of a torch tensor (or NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
