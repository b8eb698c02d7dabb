"""Check + time the tcgen05 split-bf16 contraction against fp64 and the FFMA kernel."""
import os, sys, math, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.load()
def run(M, N, K, ta, tb, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand((K, M) if ta else (M, K), generator=g, dtype=torch.float64) * 2 - 1).float()
    b = (torch.rand((N, K) if tb else (K, N), generator=g, dtype=torch.float64) * 2 - 1).float()
    ref = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double())
    ad, bd = a.to(dev), b.to(dev)
    out = torch.full((M, N), float("nan"), device=dev)
    ws = lib.opn_sgemm_workspace_bytes(M, N, K)
    ops.sgemm(ad, bd, out, trans_a=ta, trans_b=tb, M=M, N=N, K=K, lda=a.shape[1], ldb=b.shape[1], ldc=N)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.sgemm(ad, bd, out, trans_a=ta, trans_b=tb, M=M, N=N, K=K, lda=a.shape[1], ldb=b.shape[1], ldc=N)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"M={M:5d} N={N:5d} K={K:5d} ta={int(ta)} tb={int(tb)} tc={'yes' if ws else 'no '} max|err|={err:.3e} (sqrtK*1e-5={1e-5*math.sqrt(K):.1e})  {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
for shape in [(128, 128, 8192), (2048, 512, 9599), (1024, 256, 9599), (9600, 1024, 90), (1024, 90, 9600), (9600, 2048, 512), (1000, 777, 555), (9600, 9600, 128)]:
    for ta, tb in [(False, True), (True, False), (False, False)]:
        try:
            run(*shape, ta, tb)
        except Exception as e:
            print("FAILED", shape, ta, tb, e, flush=True)
