"""Decode how tcgen05.mma (cta_group::2, M=128 = 64 rows per CTA) reads an A operand held in TENSOR MEMORY:
pnr_tc_probe mode 4 tags every A-region lane with its lane id, mode 5 with 2*column+half; B selects k == n,
so D[row][n] is what the hardware read as A[row][n].  Prints the decoded maps."""
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import torch
from pixel_nerf_multiscale_b200 import _native as N

def panels(mat, rows_per_cta):
    R, K = mat.shape
    t = mat.reshape(R // rows_per_cta, rows_per_cta, K // 8, 8).permute(0, 2, 1, 3).contiguous()
    return t.to(torch.bfloat16).reshape(-1)

N.lib(); fn = N.probe_lib().pnr_tc_probe
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
for K in (16, 32):
    B = torch.zeros(256, K)
    for k in range(K):
        B[k, k] = 1.0          # D[:, n] = A[:, n] for n < K   (columns held by CTA 0's half of B)
        B[128 + k, k] = 1.0    # and again in the second N half (CTA 1's B rows)
    for mode in (4, 5):
        D = torch.full((2, 128, 128), float("nan"), device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = fn(mode, None, N.ptr(panels(B, 128).cuda()), N.ptr(D), K, N.ptr(err), N.stream_ptr(D.device))
        torch.cuda.synchronize()
        D = D.cpu()
        print("=== K", K, "mode", mode, "rc", rc, "err", int(err.item()))
        for cta in range(2):
            for lanes in ((0, 4), (30, 34), (62, 66), (94, 98), (124, 128)):
                for L in range(*lanes):
                    print("cta", cta, "lane %3d" % L, "D[0:%d]" % K, [int(x) if x == x else -1 for x in D[cta, L, :K].tolist()],
                          "| D[64:64+4]", [int(x) if x == x else -1 for x in D[cta, L, 64:68].tolist()])
