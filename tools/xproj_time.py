"""Times the x-projection of OPNet's LSTM1 ([B*T, 90] x [90, 1024]) through opn_sgemm, cold L2 (a 256 MB write between
launches).  Run once with OPN_GEMM_PROJ=0 (tcgen05 path with its operand pre-passes) and once without."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from objectpermanence_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, N, K) in [(9600, 1024, 90), (9600, 2048, 75), (76800, 1024, 90)]:
    a = torch.rand(M, K, device=dev) - 0.5
    w = torch.rand(N, K, device=dev) - 0.5
    out = torch.empty(M, N, device=dev)
    times = []
    for it in range(12):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = _lib.launch_count()
        e0.record()
        ops.sgemm(a, w, out, trans_a=False, trans_b=True, M=M, N=N, K=K, lda=K, ldb=K, ldc=N)
        e1.record()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - before
        if it >= 2:
            times.append(e0.elapsed_time(e1) * 1e3)
    times.sort()
    err = (out.double() - a.double() @ w.double().t()).abs().max().item()
    print(f"OPN_GEMM_PROJ={os.environ.get('OPN_GEMM_PROJ', '1')} [{M},{K}]x[{K},{N}]: median {times[len(times) // 2]:.1f} us, "
          f"min {times[0]:.1f} us, {launches} launches, {4e-3 * M * N / times[len(times) // 2]:.0f} GB/s of output, max err {err:.2e}")
