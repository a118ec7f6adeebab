"""Probe (negative result kept on purpose): a non-tensor cp.async.bulk cannot complete on the LEADER CTA's
mbarrier of a cta_group::2 pair -- the leader times out (tag 102); the tensor (TMA) form with .cta_group::2 can."""
import sys, os; sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
import torch
from test_gpu_tc_probe import _probe
for K in (16, 64, 128):
    g = torch.Generator().manual_seed(K)
    A = torch.randint(-3, 4, (128, K), generator=g).float(); B = torch.randint(-3, 4, (256, K), generator=g).float()
    D, err = _probe(3, A, B, K)
    ref = A @ B.t(); exp = torch.empty(2,128,128)
    for c in range(2):
        for h in range(2): exp[c, h*64:(h+1)*64] = ref[c*64:(c+1)*64, h*128:(h+1)*128]
    print("K", K, "err", err, "match", torch.equal(D, exp))
