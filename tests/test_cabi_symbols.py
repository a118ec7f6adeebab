"""The C-ABI shared library loads (no GPU needed) and exports every function that
include/pixelnerf_b200.h declares; the ctypes table binds exactly that set."""
import ctypes
import os
import re

from helpers import REPO


def _declared():
    text = open(os.path.join(REPO, "include", "pixelnerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pixel_nerf_multiscale_b200 import _native as N

    names = _declared()
    assert len(names) >= 18
    lib = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "declared in the header but not exported: %s" % n
    assert sorted(N.EXPORTS) == names, (sorted(set(names) - set(N.EXPORTS)), sorted(set(N.EXPORTS) - set(names)))


def test_abi_version_and_error_string_without_gpu():
    from pixel_nerf_multiscale_b200 import _native as N

    lib = N.lib()
    assert lib.pnr_abi_version() == 2
    assert isinstance(lib.pnr_last_error(), bytes)
    # argument validation happens before any CUDA call: NULL scene -> BAD_ARG, message set
    st = lib.pnr_point_features_f32(None, None, None, 1, 1, None, None)
    assert st == -1 and b"scene" in lib.pnr_last_error()


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the C structs (sizes implied by the header's field lists)."""
    from pixel_nerf_multiscale_b200 import _native as N

    assert ctypes.sizeof(N.Scene) == 5 * 4 + 4 * 8 * 4 + 2 * 8 * 4 + 4 + 8 * 8 + 8 + 9 * 4 + 4  # incl. padding
    assert ctypes.sizeof(N.Mlp) == 8 * 4 + 4 * 8 + 6 * 8 * 8 + 8 + 8 + 2 * 4
    assert ctypes.sizeof(N.RenderCfg) == 8 * 4
    assert ctypes.sizeof(N.RngTape) == 4 * 8 and ctypes.sizeof(N.RenderOut) == 8 * 8
