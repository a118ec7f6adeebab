"""
Hardware probes of the tcgen05 building blocks (pnr_tc_probe): UMMA shared-memory descriptors for
the SWIZZLE_NONE K-major panel layout, instruction descriptor, bulk-copy + mbarrier completion,
cross-CTA barrier relay, multicast commit, and the TMEM accumulator layout of cta_group::1 (M=128)
and cta_group::2 (M=128 = 64 rows per CTA) MMAs.  Integer-valued operands make the products exact.
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _panels(mat, rows_per_cta):
    """(R,K) -> per-CTA panel images [cta][K/8][rows][8] (bf16), flattened."""
    R, K = mat.shape
    n_cta = R // rows_per_cta
    t = mat.reshape(n_cta, rows_per_cta, K // 8, 8).permute(0, 2, 1, 3).contiguous()
    return t.to(torch.bfloat16).reshape(-1)


def _probe(mode, A, B, K):
    from pixel_nerf_multiscale_b200 import _native as N

    N.lib()
    fn = N.probe_lib().pnr_tc_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    ncols = 256 if mode == 1 else 128
    ncta = 1 if mode == 1 else 2
    mode_in, mode = mode, (2 if mode == 3 else mode)
    a_img = _panels(A, 128 if mode == 1 else 64).cuda()
    b_img = _panels(B, 256 if mode == 1 else 128).cuda()
    D = torch.full((ncta, 128, ncols), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    N.check(fn(mode_in, N.ptr(a_img), N.ptr(b_img), N.ptr(D), K, N.ptr(err), N.stream_ptr(D.device)), "pnr_tc_probe")
    torch.cuda.synchronize()
    return D.cpu(), int(err.item())


@pytest.mark.parametrize("K", [16, 64, 128])
def test_cta_group1_m128_n256(K):
    g = torch.Generator().manual_seed(K)
    A = torch.randint(-3, 4, (128, K), generator=g).float()
    B = torch.randint(-3, 4, (256, K), generator=g).float()
    D, err = _probe(1, A, B, K)
    assert err == 0, "barrier timeout tag %d" % err
    ref = A @ B.t()
    ok = torch.equal(D[0], ref)
    if not ok:
        print("mismatch fraction", (D[0] != ref).float().mean().item())
        print("D[0,:4,:8]", D[0, :4, :8], "ref", ref[:4, :8])
    assert ok


@pytest.mark.parametrize("K", [16, 64, 128])
def test_cta_group2_m128_n256(K):
    g = torch.Generator().manual_seed(100 + K)
    A = torch.randint(-3, 4, (128, K), generator=g).float()
    B = torch.randint(-3, 4, (256, K), generator=g).float()
    D, err = _probe(2, A, B, K)
    assert err == 0, "barrier timeout tag %d" % err
    ref = A @ B.t()  # (128 rows, 256 n)
    # expected "2x2" layout: CTA c, lane l, col j -> row c*64 + l%64, n = (l//64)*128 + j
    exp = torch.empty(2, 128, 128)
    for c in range(2):
        for h in range(2):
            exp[c, h * 64:(h + 1) * 64, :] = ref[c * 64:(c + 1) * 64, h * 128:(h + 1) * 128]
    ok = torch.equal(D, exp)
    if not ok:
        hyps = {}
        alt = torch.empty(2, 128, 128)  # hypothesis: rows interleaved 32-wise
        for c in range(2):
            for q in range(4):
                alt[c, q * 32:(q + 1) * 32, :] = ref[c * 64 + (q % 2) * 32:c * 64 + (q % 2) * 32 + 32,
                                                     (q // 2) * 128:(q // 2) * 128 + 128]
        hyps["2x2"] = (D == exp).float().mean().item()
        hyps["same-as-2x2-by-quadrant"] = (D == alt).float().mean().item()
        print("layout hypotheses match fractions:", hyps)
        print("D[0,0,:8]", D[0, 0, :8], "D[0,64,:8]", D[0, 64, :8], "D[1,0,:8]", D[1, 0, :8])
        print("ref[0,:8]", ref[0, :8], "ref[0,128:136]", ref[0, 128:136], "ref[64,:8]", ref[64, :8])
    assert ok
