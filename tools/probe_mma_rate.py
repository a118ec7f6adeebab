"""Cycles per tcgen05.mma (K=16, bf16) for the shapes a 64-or-128-rows-per-SM design could use."""
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import torch
from pixel_nerf_multiscale_b200 import _native as N
N.lib(); fn = N.probe_lib().pnr_tc_rate_probe
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 6 + [C.c_void_p] * 3
out = torch.zeros(4, device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
IT = 2048
print("cta_group  M    N   A-from   cycles/MMA   rows/SM   MFLOP-per-SM-per-cycle(x1e-3)  frac of 8192 flop/clk/SM")
for cg, M, Nn, ts, alt in ((2, 128, 256, 0, 1), (2, 128, 256, 0, 2), (2, 128, 256, 0, 4), (2, 128, 256, 1, 1), (2, 256, 256, 0, 1), (2, 256, 256, 0, 2),
                           (2, 256, 256, 1, 1), (1, 128, 256, 0, 1), (1, 128, 256, 0, 2), (1, 128, 256, 1, 1),
                           (1, 64, 256, 0, 1), (1, 64, 256, 0, 2), (2, 128, 128, 0, 1), (2, 128, 128, 0, 2), (2, 128, 128, 0, 4), (2, 128, 64, 0, 1), (2, 128, 64, 0, 4)):
    rc = fn(cg, M, Nn, ts, IT, alt, N.ptr(out), N.ptr(err), N.stream_ptr(out.device))
    torch.cuda.synchronize()
    if rc != 0:
        print(cg, M, Nn, ts, "rc", rc, N.lib().pnr_last_error()); continue
    cyc = out[0].item() / IT
    rows = M // cg
    flop = rows * Nn * 16 * 2
    print("alt %d" % alt, "%5d    %4d %4d   %s   %8.1f   %6d   %8.1f   %.3f   err %d" % (cg, M, Nn, "TMEM" if ts else "smem", cyc, rows, flop / cyc, flop / cyc / 8192, err.item()))
