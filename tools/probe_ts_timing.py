"""Cycles per tcgen05.mma (cta_group::2, M=128 = 64 rows per CTA, N=256, K=16, bf16) with the A operand in
shared memory (SS, mode 6) vs in tensor memory (TS, mode 7): 1024 back-to-back MMAs, clock64 on the issuing thread."""
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import torch
from pixel_nerf_multiscale_b200 import _native as N
N.lib(); fn = N.probe_lib().pnr_tc_probe
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
K = 16
Bimg = torch.zeros(256 * K, dtype=torch.bfloat16, device="cuda")
for rep in range(2):
    for mode, name in ((6, "SS (A in smem)"), (7, "TS (A in TMEM)")):
        D = torch.zeros(2, 128, 128, device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = fn(mode, None, N.ptr(Bimg), N.ptr(D), K, N.ptr(err), N.stream_ptr(D.device))
        torch.cuda.synchronize()
        d = D.flatten()[:2].tolist()
        print(name, "rc", rc, "err", int(err.item()), "issue cycles/MMA %.1f" % (d[0] / 1024), "complete cycles/MMA %.1f" % (d[1] / 1024))
