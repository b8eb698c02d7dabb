"""GPU: each C-ABI kernel against a CPU fp64 restatement (oracle / plain torch CPU math)."""
import math
import os

import pytest
import torch

from objectpermanence_b200 import _lib, ops
from oracle import opnet_oracle as oracle

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * scale).float()


# ---- sgemm --------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (130, 70, 33), (257, 129, 64), (96, 4, 512), (300, 2048, 6),
                                    (2048, 6, 1200), (64, 512, 4)])
def test_sgemm_shapes(cuda_device, ta, tb, M, N, K):
    a = _rand((K, M) if ta else (M, K), 1)
    b = _rand((N, K) if tb else (K, N), 2)
    ref = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double())
    out = torch.full((M, N), float("nan"), device=cuda_device)
    ops.sgemm(a.to(cuda_device), b.to(cuda_device), out, trans_a=ta, trans_b=tb, M=M, N=N, K=K,
              lda=a.shape[1], ldb=b.shape[1], ldc=N)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 1e-5 * max(1.0, math.sqrt(K)), err


def test_sgemm_epilogue_and_split_k(cuda_device):
    M, N, K = 40, 24, 4096  # small grid + long K -> split-K with atomics
    a, b, bias, c0 = _rand((K, M), 3), _rand((K, N), 4), _rand((N,), 5), _rand((M, N), 6)
    out = c0.clone().to(cuda_device)
    ops.sgemm(a.to(cuda_device), b.to(cuda_device), out, trans_a=True, trans_b=False, M=M, N=N, K=K, lda=M, ldb=N,
              ldc=N, alpha=0.5, beta=1.0, bias=bias.to(cuda_device))
    ref = 0.5 * (a.double().t() @ b.double()) + c0.double() + bias.double()
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-4
    # relu + bias, no split
    x, w = _rand((50, 20), 7), _rand((30, 20), 8)
    out2 = torch.empty(50, 30, device=cuda_device)
    ops.sgemm(x.to(cuda_device), w.to(cuda_device), out2, trans_a=False, trans_b=True, M=50, N=30, K=20, lda=20,
              ldb=20, ldc=30, bias=_rand((30,), 9).to(cuda_device), relu=True)
    ref2 = torch.relu(x.double() @ w.double().t() + _rand((30,), 9).double())
    assert (out2.cpu().double() - ref2).abs().max().item() <= 1e-5


def test_sgemm_strided_rows_pair_gates_with_previous_hidden(cuda_device):
    """dW_hh = sum_{b,t>=1} dgates[b,t]^T hs[b,t-1] as a flat contraction minus the video-straddling pairs."""
    B, T, G, H = 3, 7, 24, 8
    dg, hs = _rand((B, T, G), 10), _rand((B, T, H), 11)
    ref = torch.einsum("btg,bth->gh", dg[:, 1:].double(), hs[:, :-1].double())
    out = torch.empty(G, H, device=cuda_device)
    dgd, hsd = dg.to(cuda_device), hs.to(cuda_device)
    ops.sgemm(dgd, hsd, out, trans_a=True, trans_b=False, M=G, N=H, K=B * T - 1, lda=G, ldb=H, ldc=H, a_off=G)
    ops.sgemm(dgd, hsd, out, trans_a=True, trans_b=False, M=G, N=H, K=B - 1, lda=T * G, ldb=T * H, ldc=H,
              alpha=-1.0, beta=1.0, a_off=T * G, b_off=(T - 1) * H)
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-5


# ---- tcgen05 split-bf16 path of opn_sgemm ---------------------------------------------------------
def _tc_expected(M, N, K):
    return _lib.load().opn_sgemm_workspace_bytes(M, N, K) > 0


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 8192), (200, 140, 5000), (1000, 777, 555), (2048, 512, 1279)])
def test_sgemm_tensor_core_path(cuda_device, ta, tb, M, N, K):
    """Shapes above the tensor-core threshold: odd sizes (padding), split-K (small grids), all layouts."""
    assert _tc_expected(M, N, K)
    a = _rand((K, M) if ta else (M, K), 21)
    b = _rand((N, K) if tb else (K, N), 22)
    ref = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double())
    out = torch.full((M, N), float("nan"), device=cuda_device)
    ops.sgemm(a.to(cuda_device), b.to(cuda_device), out, trans_a=ta, trans_b=tb, M=M, N=N, K=K, lda=a.shape[1],
              ldb=b.shape[1], ldc=N)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    # split-bf16 operands (2^-17) + truncating tensor-core accumulation: relative to the largest entry
    assert (got - ref).abs().max().item() <= 5e-5 * ref.abs().max().item()


def test_sgemm_tensor_core_epilogue(cuda_device):
    M, N, K = 384, 256, 2048
    assert _tc_expected(M, N, K)
    a, b, bias, c0 = _rand((M, K), 23), _rand((N, K), 24), _rand((N,), 25), _rand((M, N), 26)
    ad, bd = a.to(cuda_device), b.to(cuda_device)
    prod = a.double() @ b.double().t()
    # alpha, beta = 1, bias (split-K with atomics: 6 tiles only)
    out = c0.clone().to(cuda_device)
    ops.sgemm(ad, bd, out, trans_a=False, trans_b=True, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, alpha=-0.5, beta=1.0,
              bias=bias.to(cuda_device))
    ref = -0.5 * prod + c0.double() + bias.double()
    assert (out.cpu().double() - ref).abs().max().item() <= 5e-5 * prod.abs().max().item()
    # bias + ReLU (no split-K allowed with ReLU), strided output
    big = torch.full((M, N + 8), 7.0, device=cuda_device)
    ops.sgemm(ad, bd, big, trans_a=False, trans_b=True, M=M, N=N, K=K, lda=K, ldb=K, ldc=N + 8,
              bias=bias.to(cuda_device), relu=True)
    ref2 = torch.relu(prod + bias.double())
    assert (big[:, :N].cpu().double() - ref2).abs().max().item() <= 5e-5 * prod.abs().max().item()
    assert (big[:, N:] == 7.0).all()


# ---- short-K input projections (opn_gemm_proj.cu: x W_ih^T of an LSTM layer) -------------------------------
@pytest.mark.parametrize("M,N,K,pad", [
    (9600, 1024, 90, 0),      # OPNet LSTM1 at the headline shape (learned_models.py:29)
    (1100, 2048, 75, 0),      # baseline_lstm: odd K (scalar loads), ragged last row tile
    (2048, 136, 128, 2),      # ragged column tile, K at the upper end (one CTA per SM), padded leading dimensions
    (1024, 128, 16, 0),       # smallest K
    (4096, 512, 18, 4),       # K padded to 32 inside the kernel
])
def test_sgemm_short_k_projection(cuda_device, M, N, K, pad):
    A = torch.zeros(M, K + pad); A[:, :K] = _rand((M, K), 31, 1.0)
    W = torch.zeros(N, K + pad); W[:, :K] = _rand((N, K), 32, 1.0)
    A[:, K:] = 5.0; W[:, K:] = 5.0            # must not be read into the product
    C = torch.full((M, N + pad), 7.0, device=cuda_device)
    want = A[:, :K].double() @ W[:, :K].double().t()
    ops.sgemm(A.to(cuda_device), W.to(cuda_device), C, trans_a=False, trans_b=True, M=M, N=N, K=K, lda=K + pad,
              ldb=K + pad, ldc=N + pad)
    got = C.cpu()
    assert torch.isfinite(got).all()
    assert (got[:, :N].double() - want).abs().max().item() <= 5e-5 * want.abs().max().item()
    assert (got[:, N:] == 7.0).all()
    # the 1e-2 mode: one 16-bit product
    ops.set_precision("bf16")
    try:
        ops.sgemm(A.to(cuda_device), W.to(cuda_device), C, trans_a=False, trans_b=True, M=M, N=N, K=K, lda=K + pad,
                  ldb=K + pad, ldc=N + pad)
    finally:
        ops.set_precision("fp32")
    err = (C.cpu()[:, :N].double() - want).abs().max().item()
    assert err <= 1e-2 * want.abs().max().item()


# ---- skinny contractions (rowdot / colred / tinyk fast paths and their fall-backs) ---------------------------
@pytest.mark.parametrize("case", [
    # (ta, tb, M, N, K, lda_pad, ldb_pad, ldc_pad, alpha, beta)
    (False, True, 9600, 4, 512, 0, 0, 0, 1.0, 0.0),       # y = h2 Wo^T                      (rowdot)
    (False, False, 9600, 6, 2048, 0, 0, 0, 1.0, 0.0),     # d frames_boxes = dgates2 W_ih2   (rowdot)
    (False, False, 1000, 8, 256, 4, 2, 3, 0.5, 1.0),      # rowdot with padded rows, alpha, beta = 1
    (False, False, 1000, 3, 250, 0, 0, 0, 1.0, 0.0),      # K % 4 != 0 -> tiled kernel
    (False, False, 9600, 512, 4, 0, 0, 0, 1.0, 0.0),      # dh2 = dy Wo                       (tinyk)
    (False, True, 9600, 2048, 6, 0, 0, 0, 1.0, 0.0),      # xproj2 = frames_boxes W_ih2^T     (tinyk)
    (False, False, 700, 128, 7, 1, 4, 4, 2.0, 1.0),       # tinyk with strides, alpha, beta = 1
    (True, False, 4, 512, 9600, 0, 0, 0, 1.0, 0.0),       # dWo = dy^T h2                     (colred, M skinny)
    (True, False, 2048, 6, 9600, 0, 0, 0, 1.0, 0.0),      # dW_ih2 = dgates2^T frames_boxes   (colred, N skinny)
    (True, False, 256, 15, 9600, 0, 0, 0, 1.0, 0.0),      # dW_pred^T = hs1^T dlogits         (colred)
    (True, False, 300, 16, 1100, 3, 2, 5, -1.0, 1.0),     # colred with strides, alpha, beta = 1
])
def test_sgemm_skinny_shapes(cuda_device, case):
    ta, tb, M, N, K, pa, pb, pc, alpha, beta = case
    a_shape = (K, M) if ta else (M, K)
    b_shape = (N, K) if tb else (K, N)
    A = torch.zeros(a_shape[0], a_shape[1] + pa); A[:, :a_shape[1]] = _rand(a_shape, 11, 1.0)
    Bm = torch.zeros(b_shape[0], b_shape[1] + pb); Bm[:, :b_shape[1]] = _rand(b_shape, 12, 1.0)
    C0 = torch.zeros(M, N + pc); C0[:, :N] = _rand((M, N), 13, 1.0); C0[:, N:] = 7.0
    opA = A[:, :a_shape[1]].double().t() if ta else A[:, :a_shape[1]].double()
    opB = Bm[:, :b_shape[1]].double().t() if tb else Bm[:, :b_shape[1]].double()
    want = alpha * (opA @ opB) + beta * C0[:, :N].double()
    Cd = C0.to(cuda_device)
    ops.sgemm(A.to(cuda_device), Bm.to(cuda_device), Cd, trans_a=ta, trans_b=tb, M=M, N=N, K=K, lda=A.shape[1],
              ldb=Bm.shape[1], ldc=C0.shape[1], alpha=alpha, beta=beta)
    got = Cd.cpu()
    assert (got[:, :N].double() - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())
    assert (got[:, N:] == 7.0).all()


# ---- persistent LSTM ------------------------------------------------------------------------
def _lstm_case(B, T, I, H, seed, scale=1.0):
    x = _rand((B, T, I), seed, 1.0)
    w_ih = _rand((4 * H, I), seed + 1, scale / math.sqrt(H))
    w_hh = _rand((4 * H, H), seed + 2, scale / math.sqrt(H))
    dh = _rand((B, T, H), seed + 3, 1.0)
    return x, w_ih, w_hh, dh


# flavours of the recurrence kernels: matvec on the tensor cores (split fp16, H = 256 / 512) or FP32 FMA; the FMA
# kernels exchange through the global-memory ring or, for H <= 256, through a thread-block cluster (DSMEM)
LSTM_FLAVOURS = ["mma", "ffma-l2", "ffma-cluster"]


def _set_lstm_flavour(monkeypatch, flavour, H):
    if flavour == "mma" and H not in (256, 512):
        pytest.skip("tensor-core recurrence exists for H = 256, 512")
    if flavour == "ffma-cluster" and H == 512:
        pytest.skip("H=512 has no cluster flavour (32 CTAs per batch group)")
    monkeypatch.setenv("OPN_LSTM_MATH", "mma" if flavour == "mma" else "ffma")
    monkeypatch.setenv("OPN_LSTM_EXCHANGE", "cluster" if flavour == "ffma-cluster" else "l2")


@pytest.mark.parametrize("flavour", LSTM_FLAVOURS)
@pytest.mark.parametrize("H", [32, 64, 128, 256, 512])
@pytest.mark.parametrize("B,T", [(1, 1), (2, 8), (8, 5), (11, 17), (32, 12), (70, 9)])
def test_lstm_layer_forward_backward(cuda_device, monkeypatch, flavour, H, B, T):
    _set_lstm_flavour(monkeypatch, flavour, H)
    I = 6 if H != 64 else 75
    x, w_ih, w_hh, dh = _lstm_case(B, T, I, H, seed=100 * H + B + T)
    xr, wir, whr = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh)]
    ref = oracle.lstm_layer(xr, wir, whr)
    ref.backward(dh.double())

    xg, wig, whg = [t.to(cuda_device).requires_grad_(True) for t in (x, w_ih, w_hh)]
    out = ops.lstm_layer(xg, wig, whg)
    out.backward(dh.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 2e-5
    for name, got, want in (("dx", xg.grad, xr.grad), ("dw_ih", wig.grad, wir.grad), ("dw_hh", whg.grad, whr.grad)):
        scale = max(1.0, want.abs().max().item())
        err = (got.cpu().double() - want).abs().max().item()
        assert err <= 5e-5 * scale, (name, err, scale)


@pytest.mark.parametrize("flavour", LSTM_FLAVOURS)
@pytest.mark.parametrize("H", [256, 512])
def test_lstm_saturating_weights(cuda_device, monkeypatch, flavour, H):
    _set_lstm_flavour(monkeypatch, flavour, H)
    B, T, I = 9, 40, 90
    x, w_ih, w_hh, dh = _lstm_case(B, T, I, H, seed=7, scale=8.0)
    xr, wir, whr = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh)]
    ref = oracle.lstm_layer(xr, wir, whr)
    ref.backward(dh.double())
    xg, wig, whg = [t.to(cuda_device).requires_grad_(True) for t in (x, w_ih, w_hh)]
    out = ops.lstm_layer(xg, wig, whg)
    out.backward(dh.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-4
    for got, want in ((xg.grad, xr.grad), (wig.grad, wir.grad), (whg.grad, whr.grad)):
        assert (got.cpu().double() - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("flavour", LSTM_FLAVOURS)
@pytest.mark.parametrize("H,T", [(512, 300), (256, 300), (128, 120)])
def test_lstm_long_sequence_all_flavours(cuda_device, monkeypatch, flavour, H, T):
    """T = 300 at the OPNet LSTM shapes: 299 exchanges per launch, every ring / inbox slot reused ~150 times."""
    _set_lstm_flavour(monkeypatch, flavour, H)
    B, I = 32, 90
    x, w_ih, w_hh, dh = _lstm_case(B, T, I, H, seed=31)
    dh = dh * 0.01
    xr, wir, whr = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh)]
    ref = oracle.lstm_layer(xr, wir, whr)
    ref.backward(dh.double())
    xg, wig, whg = [t.to(cuda_device).requires_grad_(True) for t in (x, w_ih, w_hh)]
    out = ops.lstm_layer(xg, wig, whg)
    out.backward(dh.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 2e-5
    print(f"\n[{flavour} H={H} T={T}] max|h - ref| = {(out.detach().cpu().double() - ref.detach()).abs().max().item():.3e}")
    for name, got, want in (("dx", xg.grad, xr.grad), ("dw_ih", wig.grad, wir.grad), ("dw_hh", whg.grad, whr.grad)):
        scale = max(1.0, want.abs().max().item())
        err = (got.cpu().double() - want).abs().max().item()
        print(f"   {name}: err {err:.3e} (scale {scale:.3e})")
        assert err <= 1e-4 * scale, (name, err, scale)


@pytest.mark.parametrize("H,B,T,I", [(48, 5, 17, 6), (100, 3, 40, 90), (200, 9, 25, 6), (384, 4, 30, 12), (20, 2, 8, 75)])
def test_lstm_layer_any_hidden_size(cuda_device, H, B, T, I):
    """Hidden sizes the kernels are not instantiated for run zero-padded in the next larger one (ops.lstm_layer): same
    function and gradients as the unpadded layer (a kernel generic in H)."""
    x, w_ih, w_hh, dh = _lstm_case(B, T, I, H, seed=900 + H)
    xr, wir, whr = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh)]
    ref = oracle.lstm_layer(xr, wir, whr)
    ref.backward(dh.double())
    xg, wig, whg = [t.to(cuda_device).requires_grad_(True) for t in (x, w_ih, w_hh)]
    out = ops.lstm_layer(xg, wig, whg)
    assert out.shape == (B, T, H)
    out.backward(dh.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 2e-5
    for name, got, want in (("dx", xg.grad, xr.grad), ("dw_ih", wig.grad, wir.grad), ("dw_hh", whg.grad, whr.grad)):
        assert got.shape == want.shape
        err = (got.cpu().double() - want).abs().max().item()
        assert err <= 5e-5 * max(1.0, want.abs().max().item()), (name, err)


def test_baseline_lstm_with_an_unlisted_hidden_size(cuda_device):
    """ModelsFactory with videos_hidden_dim = 100 (not a kernel size): forward and parameter gradients against the oracle."""
    cfg = {"videos_hidden_dim": 100}
    from objectpermanence_b200.models_factory import ModelsFactory
    from objectpermanence_b200.synthetic import make_batch
    boxes_np, labels_np, _ = make_batch(3, 20, 5, seed=77)
    boxes, labels = torch.from_numpy(boxes_np), torch.from_numpy(labels_np)
    params = oracle.init_params("baseline_lstm", cfg, seed=3)
    y_ref, _, _, g_ref = oracle.loss_and_grads("baseline_lstm", params, boxes, labels, cfg, dtype=torch.float64)
    model = ModelsFactory.get_model("baseline_lstm", cfg).to(cuda_device)
    model.load_state_dict({k: v.float() for k, v in params.items()})
    y = model(boxes.to(cuda_device))
    assert (y.detach().cpu().double() - y_ref).abs().max().item() <= 1e-4
    loss3, dy = ops.loss_and_grad(y, labels.to(cuda_device), None, False)
    y.backward(dy)
    for k, want in g_ref.items():
        got = dict(model.named_parameters())[k].grad
        assert (got.cpu().double() - want).abs().max().item() <= 2e-4 * max(1e-3, want.abs().max().item()), k


# ---- batch-wide tcgen05 recurrence (opn_lstm_tc.cu): groups of 128 videos, weights in shared memory, TMEM accumulators ----
@pytest.mark.parametrize("H", [256, 512])
@pytest.mark.parametrize("B,T", [(1, 1), (3, 2), (40, 9), (128, 6), (130, 17), (256, 5)])
def test_lstm_tcgen05_forward_backward(cuda_device, monkeypatch, H, B, T):
    """OPN_LSTM_TC=1 forces the batch-wide kernels at every batch size: ragged groups (B % 128 != 0), two groups, T = 1."""
    monkeypatch.setenv("OPN_LSTM_TC", "1")
    x, w_ih, w_hh, dh = _lstm_case(B, T, 6, H, seed=7 * H + B + T)
    xr, wir, whr = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh)]
    ref = oracle.lstm_layer(xr, wir, whr)
    ref.backward(dh.double())
    xg, wig, whg = [t.to(cuda_device).requires_grad_(True) for t in (x, w_ih, w_hh)]
    out = ops.lstm_layer(xg, wig, whg)
    out.backward(dh.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 2e-5
    for name, got, want in (("dx", xg.grad, xr.grad), ("dw_ih", wig.grad, wir.grad), ("dw_hh", whg.grad, whr.grad)):
        scale = max(1.0, want.abs().max().item())
        err = (got.cpu().double() - want).abs().max().item()
        assert err <= 5e-5 * scale, (name, err, scale)


@pytest.mark.parametrize("H", [256, 512])
def test_lstm_tcgen05_long_sequence_and_saturating_weights(cuda_device, monkeypatch, H):
    """T = 300 (299 exchanges, every ring slot reused ~150 times) at default-init weights, and x8 weights at T = 40."""
    monkeypatch.setenv("OPN_LSTM_TC", "1")
    for B, T, scale, tol_h, tol_g in ((140, 300, 1.0, 2e-5, 1e-4), (9, 40, 8.0, 1e-4, 1e-3)):
        x, w_ih, w_hh, dh = _lstm_case(B, T, 90, H, seed=31, scale=scale)
        dh = dh * 0.01
        xr, wir, whr = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh)]
        ref = oracle.lstm_layer(xr, wir, whr)
        ref.backward(dh.double())
        xg, wig, whg = [t.to(cuda_device).requires_grad_(True) for t in (x, w_ih, w_hh)]
        out = ops.lstm_layer(xg, wig, whg)
        out.backward(dh.to(cuda_device))
        assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= tol_h
        for name, got, want in (("dx", xg.grad, xr.grad), ("dw_ih", wig.grad, wir.grad), ("dw_hh", whg.grad, whr.grad)):
            err = (got.cpu().double() - want).abs().max().item()
            assert err <= tol_g * max(1.0, want.abs().max().item()), (name, err)


def test_lstm_inference_mode_skips_stash(cuda_device):
    x, w_ih, w_hh, _ = _lstm_case(4, 9, 6, 128, seed=21)
    ref = oracle.lstm_layer(x.double(), w_ih.double(), w_hh.double())
    with torch.no_grad():
        out = ops.lstm_layer(x.to(cuda_device), w_ih.to(cuda_device), w_hh.to(cuda_device))
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-5


def test_no_grad_skips_the_stash_even_when_the_weights_require_grad(cuda_device, monkeypatch):
    """needs_input_grad mirrors requires_grad whatever the grad mode: real modules (parameters require grad) under
    torch.no_grad() must still launch without the gates / cells stash (NULL pointers), LSTM layer and fused OPNet."""
    seen = {}
    lib = _lib.load()

    class Spy:
        def __getattr__(self, name):
            fn = getattr(lib, name)
            if name not in ("opn_lstm_fwd", "opn_opnet_fwd"):
                return fn

            def call(*args):
                seen[name] = args
                return fn(*args)
            return call

    monkeypatch.setattr(ops._lib, "load", lambda: Spy())
    x, w_ih, w_hh, _ = _lstm_case(4, 9, 6, 128, seed=21)
    wi, wh = w_ih.to(cuda_device).requires_grad_(True), w_hh.to(cuda_device).requires_grad_(True)
    with torch.no_grad():
        out = ops.lstm_layer(x.to(cuda_device), wi, wh)
    assert seen["opn_lstm_fwd"][6] is None and seen["opn_lstm_fwd"][7] is None     # gates, cells
    assert not out.requires_grad
    out = ops.lstm_layer(x.to(cuda_device), wi, wh)
    assert seen["opn_lstm_fwd"][6] is not None and out.requires_grad
    H1, H2 = 256, 512
    w = [(_rand(s_, 70 + i, 0.05)).to(cuda_device).requires_grad_(True) for i, s_ in
         enumerate([(4 * H1, 90), (4 * H1, H1), (15, H1), (4 * H2, 6), (4 * H2, H2)])]
    boxes = _rand((2, 5, 15, 6), 9).abs().to(cuda_device)
    with torch.no_grad():
        ops.opnet_trunk(boxes, *w)
    a = seen["opn_opnet_fwd"]
    assert a[11] is None and a[12] is None and a[17] is None and a[18] is None        # gates1, cells1, gates2, cells2
    ops.opnet_trunk(boxes, *w)
    assert seen["opn_opnet_fwd"][11] is not None


def test_lstm_unsupported_hidden_size(cuda_device):
    """Hidden sizes up to 512 run (zero-padded into the next kernel size: test_lstm_layer_any_hidden_size); larger ones, and
    sizes the C ABI is called with directly, are refused with a message, never computed on another path."""
    x, w_ih, w_hh, _ = _lstm_case(2, 3, 6, 640, seed=5)
    with pytest.raises(_lib.OpnError, match="unsupported"):
        ops.lstm_layer(x.to(cuda_device), w_ih.to(cuda_device), w_hh.to(cuda_device))
    x, w_ih, w_hh, _ = _lstm_case(2, 3, 6, 48, seed=5)
    with pytest.raises(_lib.OpnError, match="unsupported"):
        ops.LstmLayerFn.apply(x.to(cuda_device), w_ih.to(cuda_device), w_hh.to(cuda_device), False)


@pytest.mark.parametrize("B,T", [(1, 1), (3, 2), (8, 40), (11, 37), (25, 64), (32, 300)])
def test_opnet_forward_producer_consumer_split_equals_the_single_kernel(cuda_device, monkeypatch, B, T):
    """The split forward (LSTM2 loop on 128 CTAs + LSTM1 / who-to-track producer on the idle SMs, opn_opnet_l1head.cu) against
    the single fused kernel (OPN_OPNET_SPLIT=0): every output of opn_opnet_fwd, ragged groups, T = 1 and 2 (ring slots and the
    producer's back-pressure from its head CTA), and no time-out on the status page."""
    lib = _lib.load()
    H1, H2 = 256, 512
    g = torch.Generator().manual_seed(17 * B + T)
    r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(cuda_device)
    boxes = torch.rand(B, T, 15, 6, generator=g).to(cuda_device)
    xproj1 = r(B, T, 4 * H1) * 0.5
    w_hh1, w_pred, w_ih2, w_hh2 = r(4 * H1, H1) / H1 ** 0.5, r(15, H1) / H1 ** 0.5, r(4 * H2, 6) / H2 ** 0.5, r(4 * H2, H2) / H2 ** 0.5
    shapes = [(B, T, H1), (B, T, 4 * H1), (B, T, H1), (B, 15, T), (B, T, 15), (B, T, 6), (B, T, H2), (B, T, 4 * H2), (B, T, H2)]
    s = torch.cuda.current_stream().cuda_stream
    results = {}
    for split in ("0", "1"):
        monkeypatch.setenv("OPN_OPNET_SPLIT", split)
        outs = [torch.full(sh, float("nan"), device=cuda_device) for sh in shapes]
        ws = torch.zeros(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=cuda_device)
        for _ in range(2):      # twice: the second call reuses the zeroed-per-call workspace
            rc = lib.opn_opnet_fwd(B, T, H1, H2, boxes.data_ptr(), xproj1.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(),
                                   w_ih2.data_ptr(), w_hh2.data_ptr(), *[o.data_ptr() for o in outs], ws.data_ptr(), ws.numel(), s)
            _lib.check(rc, "opn_opnet_fwd")
        ops.check_status(cuda_device, "opn_opnet_fwd")
        results[split] = outs
    for name, a, b in zip(["hs1", "gates1", "cells1", "logits", "probs", "fb", "hs2", "gates2", "cells2"], results["1"], results["0"]):
        assert not torch.isnan(a).any(), name
        assert (a - b).abs().max().item() <= 2e-6, name


@pytest.mark.parametrize("B,T", [(1, 1), (3, 2), (8, 40), (11, 37), (25, 64), (32, 300)])
def test_opnet_backward_two_kernel_split_equals_the_single_kernel(cuda_device, monkeypatch, B, T):
    """The split backward (LSTM2 loop on 128 CTAs + head backward / LSTM1 reverse recurrence on the idle SMs,
    opn_opnet_l1bwd.cu) against the single fused kernel (OPN_OPNET_SPLIT=0): d_gates1, d_gates2, d_logits; ragged groups, T = 1
    and 2, twice per workspace, no time-out on the status page."""
    lib = _lib.load()
    H1, H2 = 256, 512
    g = torch.Generator().manual_seed(29 * B + T)
    R = lambda *s: torch.rand(*s, generator=g).to(cuda_device)
    boxes = R(B, T, 15, 6)
    probs = torch.softmax(torch.randn(B, T, 15, generator=g), -1).to(cuda_device)
    w_hh1, w_pred = (R(4 * H1, H1) * 2 - 1) / H1 ** 0.5, (R(15, H1) * 2 - 1) / H1 ** 0.5
    w_ih2, w_hh2 = (R(4 * H2, 6) * 2 - 1) / H2 ** 0.5, (R(4 * H2, H2) * 2 - 1) / H2 ** 0.5
    g1, g2 = R(B, T, 4 * H1), R(B, T, 4 * H2)
    g1[..., 2 * H1:3 * H1] = g1[..., 2 * H1:3 * H1] * 2 - 1      # the g gate is a tanh
    g2[..., 2 * H2:3 * H2] = g2[..., 2 * H2:3 * H2] * 2 - 1
    c1, c2 = torch.randn(B, T, H1, generator=g).to(cuda_device) * 0.5, torch.randn(B, T, H2, generator=g).to(cuda_device) * 0.5
    dh2 = torch.randn(B, T, H2, generator=g).to(cuda_device) * 0.01
    s = torch.cuda.current_stream().cuda_stream
    results = {}
    for split in ("0", "1"):
        monkeypatch.setenv("OPN_OPNET_SPLIT", split)
        outs = [torch.full(sh, float("nan"), device=cuda_device) for sh in ((B, T, 4 * H1), (B, T, 4 * H2), (B, T, 15))]
        ws = torch.zeros(lib.opn_opnet_bwd_workspace_bytes(B, T), dtype=torch.uint8, device=cuda_device)
        for _ in range(2):
            rc = lib.opn_opnet_bwd(B, T, H1, H2, boxes.data_ptr(), probs.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(), w_ih2.data_ptr(),
                                   w_hh2.data_ptr(), g1.data_ptr(), c1.data_ptr(), g2.data_ptr(), c2.data_ptr(), dh2.data_ptr(),
                                   *[o.data_ptr() for o in outs], ws.data_ptr(), ws.numel(), s)
            _lib.check(rc, "opn_opnet_bwd")
        ops.check_status(cuda_device, "opn_opnet_bwd")
        results[split] = outs
    for name, a, b in zip(["d_gates1", "d_gates2", "d_logits"], results["1"], results["0"]):
        assert not torch.isnan(a).any(), name
        assert (a - b).abs().max().item() <= 2e-6 * max(1e-6, b.abs().max().item()) + 1e-12, name


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("B,T,launches", [(8, 2000, 12), (32, 300, 40)])
def test_opnet_exchange_is_repeatable_over_many_launches(cuda_device, monkeypatch, split, B, T, launches):
    """Stress of the inter-CTA hand-over (flagged words through L2 read by polls and by TMA bulk copies, the producer / consumer
    rings): a torn or stale word would change a hidden state and every frame after it, so many launches of the fused forward and
    backward on the same inputs -- 2000 dependent frames each at config 4's shape -- must agree BIT FOR BIT with the first one, in
    the single-kernel and in the split form, with no time-out on the status page (formal-memory-model item)."""
    lib = _lib.load()
    monkeypatch.setenv("OPN_OPNET_SPLIT", split)
    H1, H2 = 256, 512
    g = torch.Generator().manual_seed(5 * B + T)
    r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(cuda_device)
    boxes = torch.rand(B, T, 15, 6, generator=g).to(cuda_device)
    xproj1 = r(B, T, 4 * H1) * 0.5
    w_hh1, w_pred, w_ih2, w_hh2 = r(4 * H1, H1) / H1 ** 0.5, r(15, H1) / H1 ** 0.5, r(4 * H2, 6) / H2 ** 0.5, r(4 * H2, H2) / H2 ** 0.5
    dh2 = r(B, T, H2) * 0.01
    fshapes = [(B, T, H1), (B, T, 4 * H1), (B, T, H1), (B, 15, T), (B, T, 15), (B, T, 6), (B, T, H2), (B, T, 4 * H2), (B, T, H2)]
    bshapes = [(B, T, 4 * H1), (B, T, 4 * H2), (B, T, 15)]
    s = torch.cuda.current_stream().cuda_stream
    first = None
    for it in range(launches):
        f = [torch.full(sh, float("nan"), device=cuda_device) for sh in fshapes]
        ws = torch.empty(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=cuda_device)
        _lib.check(lib.opn_opnet_fwd(B, T, H1, H2, boxes.data_ptr(), xproj1.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(),
                                     w_ih2.data_ptr(), w_hh2.data_ptr(), *[o.data_ptr() for o in f], ws.data_ptr(), ws.numel(), s),
                   "opn_opnet_fwd")
        hs1, gates1, cells1, _, probs, _, _, gates2, cells2 = f
        b = [torch.full(sh, float("nan"), device=cuda_device) for sh in bshapes]
        wb = torch.empty(lib.opn_opnet_bwd_workspace_bytes(B, T), dtype=torch.uint8, device=cuda_device)
        _lib.check(lib.opn_opnet_bwd(B, T, H1, H2, boxes.data_ptr(), probs.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(),
                                     w_ih2.data_ptr(), w_hh2.data_ptr(), gates1.data_ptr(), cells1.data_ptr(), gates2.data_ptr(),
                                     cells2.data_ptr(), dh2.data_ptr(), *[o.data_ptr() for o in b], wb.data_ptr(), wb.numel(), s),
                   "opn_opnet_bwd")
        ops.check_status(cuda_device, "opn_opnet_fwd / opn_opnet_bwd")
        outs = f + b
        if first is None:
            first = outs
            for o in outs:
                assert torch.isfinite(o).all()
        else:
            for k, (a, ref) in enumerate(zip(outs, first)):
                assert torch.equal(a, ref), f"launch {it}: output {k} differs from the first launch"


# ---- fused OPNet forward ----------------------------------------------------------------------
# "fused_inline" after "fused" (two-stream weight gradients, the default) is the order in which [11-37] failed once in
# round 1; root cause and fix: DESIGN.md section 9 (leftover shared memory read by the fused backward's first sweep)
_BWD_MODES = ["fused", "fused_inline", "separate"]


@pytest.mark.parametrize("bwd", _BWD_MODES)
@pytest.mark.parametrize("B,T", [(1, 1), (3, 2), (11, 37), (32, 64), (70, 9), (8, 300)])
def test_opnet_fused_forward_matches_separate_kernels(cuda_device, monkeypatch, B, T, bwd):
    """LSTM1 + who-to-track + LSTM2 as one persistent kernel (and the mirror-image fused backward: both reverse
    recurrences + the who-to-track backward) against the chain of separate kernels and the fp64 oracle: outputs and the
    gradients of all five weight matrices through every path."""
    monkeypatch.setenv("OPN_OPNET_FUSED_BWD", "0" if bwd == "separate" else "1")
    monkeypatch.setenv("OPN_OPNET_WGRAD_OVERLAP", "0" if bwd == "fused_inline" else "1")
    H1, H2 = 256, 512
    gen = torch.Generator().manual_seed(1000 + B * T)   # the padding mask too: round 1 drew it from the global generator
    boxes = torch.rand(B, T, 15, 6, generator=gen) * (torch.rand(B, T, 15, 1, generator=gen) > 0.3)
    w = {"ih1": _rand((4 * H1, 90), 1, 1 / math.sqrt(H1)), "hh1": _rand((4 * H1, H1), 2, 1 / math.sqrt(H1)),
         "pred": _rand((15, H1), 3, 1 / math.sqrt(H1)), "ih2": _rand((4 * H2, 6), 4, 1 / math.sqrt(H2)),
         "hh2": _rand((4 * H2, H2), 5, 1 / math.sqrt(H2))}
    dh2 = _rand((B, T, H2), 6, 0.01)

    def run(fused):
        ws = {k: v.to(cuda_device).requires_grad_(True) for k, v in w.items()}
        bx = boxes.to(cuda_device)
        if fused:
            h2, logits = ops.opnet_trunk(bx, ws["ih1"], ws["hh1"], ws["pred"], ws["ih2"], ws["hh2"])
        else:
            h1 = ops.lstm_layer(bx.reshape(B, T, -1), ws["ih1"], ws["hh1"])
            fb, logits = ops.who_to_track(bx, h1, ws["pred"])
            h2 = ops.lstm_layer(fb, ws["ih2"], ws["hh2"])
        h2.backward(dh2.to(cuda_device))
        return h2.detach().cpu(), logits.detach().cpu(), {k: v.grad.cpu() for k, v in ws.items()}

    h2_f, lg_f, g_f = run(True)
    h2_s, lg_s, g_s = run(False)
    wr = {k: v.double().requires_grad_(True) for k, v in w.items()}
    h1_r = oracle.lstm_layer(boxes.double().reshape(B, T, -1), wr["ih1"], wr["hh1"])
    fb_r, lg_r = oracle.who_to_track(boxes.double(), h1_r, wr["pred"])
    h2_r = oracle.lstm_layer(fb_r, wr["ih2"], wr["hh2"])
    h2_r.backward(dh2.double())
    assert (h2_f.double() - h2_r.detach()).abs().max().item() <= 2e-5
    assert (lg_f.double() - lg_r.detach().permute(0, 2, 1)).abs().max().item() <= 2e-5
    assert (h2_f - h2_s).abs().max().item() <= 1e-5 and (lg_f - lg_s).abs().max().item() <= 1e-5
    for k in w:
        want = wr[k].grad
        scale = max(1.0, want.abs().max().item())
        assert (g_f[k].double() - want).abs().max().item() <= 1e-4 * scale, k
        assert (g_s[k].double() - want).abs().max().item() <= 1e-4 * scale, k


def test_opnet_fused_backward_falls_back_when_logits_carry_gradient(cuda_device, monkeypatch):
    """The fused backward assumes no gradient on the who-to-track logits (the reference's training loss); when one
    arrives the separate kernels run.  Both must agree with the oracle."""
    monkeypatch.setenv("OPN_OPNET_FUSED_BWD", "1")
    B, T, H1, H2 = 5, 21, 256, 512
    boxes = torch.rand(B, T, 15, 6, generator=torch.Generator().manual_seed(2))
    w = {"ih1": _rand((4 * H1, 90), 1, 1 / math.sqrt(H1)), "hh1": _rand((4 * H1, H1), 2, 1 / math.sqrt(H1)),
         "pred": _rand((15, H1), 3, 1 / math.sqrt(H1)), "ih2": _rand((4 * H2, 6), 4, 1 / math.sqrt(H2)),
         "hh2": _rand((4 * H2, H2), 5, 1 / math.sqrt(H2))}
    dh2, dlg = _rand((B, T, H2), 6, 0.01), _rand((B, 15, T), 7, 0.01)
    ws = {k: v.to(cuda_device).requires_grad_(True) for k, v in w.items()}
    h2, logits = ops.opnet_trunk(boxes.to(cuda_device), ws["ih1"], ws["hh1"], ws["pred"], ws["ih2"], ws["hh2"])
    torch.autograd.backward([h2, logits], [dh2.to(cuda_device), dlg.to(cuda_device)])
    wr = {k: v.double().requires_grad_(True) for k, v in w.items()}
    h1_r = oracle.lstm_layer(boxes.double().reshape(B, T, -1), wr["ih1"], wr["hh1"])
    fb_r, lg_r = oracle.who_to_track(boxes.double(), h1_r, wr["pred"])
    h2_r = oracle.lstm_layer(fb_r, wr["ih2"], wr["hh2"])
    torch.autograd.backward([h2_r, lg_r.permute(0, 2, 1)], [dh2.double(), dlg.double()])
    for k in w:
        want = wr[k].grad
        assert (ws[k].grad.cpu().double() - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item()), k


def test_opnet_fused_forward_rejects_other_configs(cuda_device):
    lib = _lib.load()
    assert lib.opn_opnet_fwd(2, 2, 128, 512, *([None] * 16), 0, None) != 0
    assert b"shipped OPNet config" in lib.opn_last_error()
    assert lib.opn_opnet_bwd(2, 2, 256, 256, *([None] * 15), 0, None) != 0
    assert b"shipped OPNet config" in lib.opn_last_error()


# ---- who-to-track ---------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,H1", [(1, 1, 32), (3, 11, 256), (5, 300, 64)])
def test_who_to_track(cuda_device, B, T, H1):
    from objectpermanence_b200.synthetic import make_batch
    boxes = torch.from_numpy(make_batch(B, T, 6, seed=B + T)[0])
    hs1, wp = _rand((B, T, H1), 31), _rand((15, H1), 32, 2.0 / math.sqrt(H1))
    dfb, dlog = _rand((B, T, 6), 33), _rand((B, 15, T), 34)
    hr, wr = hs1.double().requires_grad_(True), wp.double().requires_grad_(True)
    fb_ref, logits_ref = oracle.who_to_track(boxes.double(), hr, wr)
    logits_ref = logits_ref.permute(0, 2, 1).contiguous()
    (fb_ref * dfb.double()).sum().add((logits_ref * dlog.double()).sum()).backward()

    hg, wg = hs1.to(cuda_device).requires_grad_(True), wp.to(cuda_device).requires_grad_(True)
    fb, logits = ops.who_to_track(boxes.to(cuda_device), hg, wg)
    assert logits.shape == (B, 15, T)
    assert (fb.detach().cpu().double() - fb_ref.detach()).abs().max().item() <= 1e-5
    assert (logits.detach().cpu().double() - logits_ref.detach()).abs().max().item() <= 1e-5
    ((fb * dfb.to(cuda_device)).sum() + (logits * dlog.to(cuda_device)).sum()).backward()
    assert (hg.grad.cpu().double() - hr.grad).abs().max().item() <= 1e-5
    assert (wg.grad.cpu().double() - wr.grad).abs().max().item() <= 2e-4


# ---- encoder helpers ------------------------------------------------------------------------
def test_add_layer_norm(cuda_device):
    rows, D = 77, 256
    x, r, w, b, dy = _rand((rows, D), 41), _rand((rows, D), 42), 1 + _rand((D,), 43, 0.1), _rand((D,), 44, 0.1), _rand(
        (rows, D), 45)
    xr, rr, wr, br = [t.double().requires_grad_(True) for t in (x, r, w, b)]
    ref = oracle.layer_norm(xr + rr, wr, br)
    ref.backward(dy.double())
    xg, rg, wg, bg = [t.to(cuda_device).requires_grad_(True) for t in (x, r, w, b)]
    out = ops.add_layer_norm(xg, rg, wg, bg)
    out.backward(dy.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-5
    for got, want in ((xg.grad, xr.grad), (rg.grad, rr.grad), (wg.grad, wr.grad), (bg.grad, br.grad)):
        assert (got.cpu().double() - want).abs().max().item() <= 1e-4


@pytest.mark.parametrize("S,D,nhead", [(48, 32, 2), (300, 256, 2), (1000, 64, 4)])
def test_self_attention(cuda_device, S, D, nhead):
    qkv, dctx = _rand((S, 3 * D), 51), _rand((S, D), 52)
    qr = qkv.double().requires_grad_(True)
    d = D // nhead
    q, k, v = [z.reshape(S, nhead, d).permute(1, 0, 2) for z in (qr[:, :D], qr[:, D:2 * D], qr[:, 2 * D:])]
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1) @ v).permute(1, 0, 2).reshape(S, D)
    ref.backward(dctx.double())
    qg = qkv.to(cuda_device).requires_grad_(True)
    out = ops.self_attention(qg, nhead)
    out.backward(dctx.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-5
    assert (qg.grad.cpu().double() - qr.grad).abs().max().item() <= 1e-4


@pytest.mark.parametrize("S,nhead", [(1, 1), (12, 2), (64, 1), (130, 2), (333, 1), (600, 2), (1000, 1)])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_fused_attention_head_dim_128(cuda_device, S, nhead, p_drop):
    """The flash-style tcgen05 kernels (head dimension 128: the shipped transformer_lstm config) against plain fp64 math:
    ragged S (tail tiles of both the 128-row and the 64-row tilings), one and two heads, with and without attention-weight
    dropout (mask from the oracle's restatement of the generator); output and the gradient of the packed q|k|v."""
    from oracle import dropout_mask
    D, d = 128 * nhead, 128
    seed, offset = 4242, 77
    qkv, dctx = _rand((S, 3 * D), 81 + S), _rand((S, D), 82 + S)
    qr = qkv.double().requires_grad_(True)
    q, k, v = [z.reshape(S, nhead, d).permute(1, 0, 2) for z in (qr[:, :D], qr[:, D:2 * D], qr[:, 2 * D:])]
    probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
    if p_drop > 0:
        blocks = (S * S + 3) // 4
        keep = torch.stack([torch.from_numpy(dropout_mask.keep_mask(S * S, p_drop, seed, offset + h * blocks)).reshape(S, S)
                            for h in range(nhead)])
        probs = probs * keep.double() / (1.0 - p_drop)
    ref = (probs @ v).permute(1, 0, 2).reshape(S, D)
    ref.backward(dctx.double())
    qg = qkv.to(cuda_device).requires_grad_(True)
    out = ops.self_attention(qg, nhead, p_drop, seed, offset)
    assert out.grad_fn.__class__.__name__.startswith("FusedSelfAttentionFn")
    out.backward(dctx.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-5
    assert (qg.grad.cpu().double() - qr.grad).abs().max().item() <= 1e-4 * max(1.0, qr.grad.abs().max().item())


@pytest.mark.parametrize("S,nhead,sms", [(600, 2, 3), (1000, 1, 7), (333, 1, 2), (130, 2, 3)])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_fused_attention_split_rows(cuda_device, monkeypatch, S, nhead, sms, p_drop):
    """The work split of the fused kernels: with more (resident tile, head) rows than SMs the rows of the last, partial
    round are cut into segments whose partial results are merged (config 3: 150 rows on 148 SMs).  OPN_ATTN_SMS pretends
    a small SM count so that whole rows AND segments occur at sizes fp64 math handles; the result must not depend on it."""
    from oracle import dropout_mask
    D, d = 128 * nhead, 128
    seed, offset = 99, 5
    qkv, dctx = _rand((S, 3 * D), 181 + S), _rand((S, D), 182 + S)
    qr = qkv.double().requires_grad_(True)
    q, k, v = [z.reshape(S, nhead, d).permute(1, 0, 2) for z in (qr[:, :D], qr[:, D:2 * D], qr[:, 2 * D:])]
    probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
    if p_drop > 0:
        blocks = (S * S + 3) // 4
        keep = torch.stack([torch.from_numpy(dropout_mask.keep_mask(S * S, p_drop, seed, offset + h * blocks)).reshape(S, S)
                            for h in range(nhead)])
        probs = probs * keep.double() / (1.0 - p_drop)
    ref = (probs @ v).permute(1, 0, 2).reshape(S, D)
    ref.backward(dctx.double())
    results = []
    for env in (str(sms), "148"):
        monkeypatch.setenv("OPN_ATTN_SMS", env)
        qg = qkv.to(cuda_device).requires_grad_(True)
        out = ops.self_attention(qg, nhead, p_drop, seed, offset)
        out.backward(dctx.to(cuda_device))
        assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-5
        assert (qg.grad.cpu().double() - qr.grad).abs().max().item() <= 1e-4 * max(1.0, qr.grad.abs().max().item())
        results.append((out.detach().cpu(), qg.grad.cpu()))
    # the two schedules agree far inside the tolerance (different summation order only)
    assert (results[0][0] - results[1][0]).abs().max().item() <= 2e-6
    assert (results[0][1] - results[1][1]).abs().max().item() <= 2e-5 * max(1.0, qr.grad.abs().max().item())


@pytest.mark.parametrize("B,T,M,N,shift", [(1, 64, 128, 64, 0), (3, 50, 128, 6, 0), (2, 100, 256, 90, 0), (5, 37, 256, 256, 1),
                                            (4, 300, 1024, 256, 1), (7, 33, 2048, 512, 1), (2, 70, 512, 300, 1), (1, 1, 128, 15, 1)])
@pytest.mark.parametrize("mode", ["fp32", "16bit"])
def test_wgrad_tc_kernel(cuda_device, B, T, M, N, shift, mode):
    """The tcgen05 weight-gradient kernel (csrc/opn_wgrad_tc.cu) against fp64: ragged row counts (tail k-block), N below /
    across / above the 64- and 256-column blocks, unaligned rows of b (N = 6, 15, 90), the one-row shift with the first
    frame of every video excluded (dW_hh), several jobs in one launch."""
    rows = B * T
    a, b = _rand((rows, M), 300 + rows), _rand((rows, N), 301 + rows)
    a2 = _rand((rows, M), 302 + rows)
    bs = b.double()
    if shift:
        bs = torch.roll(bs, 1, dims=0)
        bs[::T] = 0.0      # rows r with r % T == 0 meet nothing
    want, want2 = a.double().t() @ bs, a2.double().t() @ b.double()
    ad, bd, a2d = a.to(cuda_device), b.to(cuda_device), a2.to(cuda_device)
    out = torch.full((M, N), float("nan"), device=cuda_device)
    out2 = torch.full((M, N + 3), float("nan"), device=cuda_device)      # a strided destination
    out3 = torch.full((N, M), float("nan"), device=cuda_device)          # the transposed form of job 1
    ops.set_precision(mode)
    try:
        ops.wgrad_jobs_run([(ad, bd, out, T, shift), (a2d, bd, out2[:, :N], T, 0), (ad, bd, out3, T, shift, True)])
    finally:
        ops.set_precision("fp32")
    assert torch.equal(out3.t(), out)
    torch.cuda.synchronize()
    tol = (3e-5 if mode == "fp32" else 1e-2) * max(1.0, want.abs().max().item())
    assert (out.cpu().double() - want).abs().max().item() <= tol
    assert (out2[:, :N].cpu().double() - want2).abs().max().item() <= tol
    assert torch.isnan(out2[:, N:]).all()      # nothing written outside the N columns


# ---- dropout (train mode of the encoder layer) ------------------------------------------------
@pytest.mark.parametrize("n,p,seed,offset", [(1, 0.1, 1, 0), (4, 0.1, 7, 3), (1027, 0.1, 1234, 0), (65536, 0.5, 2 ** 40 + 5, 2 ** 33),
                                              (100003, 0.0, 9, 11), (300 * 300 + 1, 0.9, 3, 12345)])
def test_dropout_mask_is_the_specified_philox_stream(cuda_device, n, p, seed, offset):
    from oracle import dropout_mask
    x = _rand((n,), 70) + 2.0           # no zeros: a zero in the output is a dropped element
    want = dropout_mask.dropout(x.numpy(), p, seed, offset)
    xd = x.to(cuda_device)
    out = torch.full_like(xd, float("nan"))
    ops.dropout_(xd, out, p, seed, offset)
    assert torch.equal(out.cpu(), torch.from_numpy(want))
    # unaligned views take the scalar path and see the same stream; in place is allowed
    if n > 8:
        buf = torch.zeros(n + 1, device=cuda_device)
        buf[1:] = xd
        ops.dropout_(buf[1:], buf[1:], p, seed, offset)
        assert torch.equal(buf[1:].cpu(), torch.from_numpy(want))


def test_dropout_keep_rate_and_backward(cuda_device):
    torch.manual_seed(5)
    x = (_rand((512, 2048), 71) + 2.0).to(cuda_device).requires_grad_(True)
    y = ops.dropout(x, 0.1, training=True)
    kept = (y != 0)
    assert abs(kept.float().mean().item() - 0.9) < 2e-3
    assert (y[kept] - x.detach()[kept] / 0.9).abs().max().item() < 1e-5
    w = _rand((512, 2048), 72).to(cuda_device)
    (y * w).sum().backward()
    assert (x.grad - torch.where(kept, w / 0.9, torch.zeros_like(w))).abs().max().item() < 1e-6
    # consecutive sites never share mask words; eval mode / p = 0 are the identity (same tensor, no launch)
    y2 = ops.dropout(x, 0.1, training=True)
    assert not torch.equal(y2 != 0, kept)
    assert ops.dropout(x, 0.1, training=False) is x and ops.dropout(x, 0.0, training=True) is x
    # the stream restarts with torch.manual_seed
    torch.manual_seed(5)
    assert torch.equal(ops.dropout(x, 0.1, training=True), y)


@pytest.mark.parametrize("S,D,nhead", [(48, 32, 2), (301, 64, 2)])
def test_self_attention_with_weight_dropout(cuda_device, S, D, nhead):
    """nn.MultiheadAttention drops attention weights after the softmax; checked against plain fp64 math with the
    mask of the same (seed, offset) taken from the oracle's restatement of the generator."""
    from oracle import dropout_mask
    p_drop, seed, offset = 0.1, 99, 1000
    qkv, dctx = _rand((S, 3 * D), 53), _rand((S, D), 54)
    qr = qkv.double().requires_grad_(True)
    d = D // nhead
    q, k, v = [z.reshape(S, nhead, d).permute(1, 0, 2) for z in (qr[:, :D], qr[:, D:2 * D], qr[:, 2 * D:])]
    blocks = (S * S + 3) // 4
    keep = torch.stack([torch.from_numpy(dropout_mask.keep_mask(S * S, p_drop, seed, offset + h * blocks)).reshape(S, S)
                        for h in range(nhead)])
    probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1) * keep.double() / (1.0 - p_drop)
    ref = (probs @ v).permute(1, 0, 2).reshape(S, D)
    ref.backward(dctx.double())
    qg = qkv.to(cuda_device).requires_grad_(True)
    out = ops.self_attention(qg, nhead, p_drop, seed, offset)
    out.backward(dctx.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-5
    assert (qg.grad.cpu().double() - qr.grad).abs().max().item() <= 1e-4


def test_linear_bias_relu(cuda_device):
    x, w, b, dy = _rand((4, 9, 20), 61), _rand((33, 20), 62), _rand((33,), 63), _rand((4, 9, 33), 64)
    xr, wr, br = [t.double().requires_grad_(True) for t in (x, w, b)]
    ref = torch.relu(xr @ wr.t() + br)
    ref.backward(dy.double())
    xg, wg, bg = [t.to(cuda_device).requires_grad_(True) for t in (x, w, b)]
    out = ops.linear(xg, wg, bg, relu=True)
    out.backward(dy.to(cuda_device))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-5
    for got, want in ((xg.grad, xr.grad), (wg.grad, wr.grad), (bg.grad, br.grad)):
        assert (got.cpu().double() - want).abs().max().item() <= 1e-4


# ---- training loss --------------------------------------------------------------------------
@pytest.mark.parametrize("no_labels", [False, True])
def test_training_loss(cuda_device, no_labels):
    B, T = 5, 37
    y, labels = _rand((B, T, 4), 71, 0.5) + 0.5, _rand((B, T, 4), 72, 0.5) + 0.5
    y[0, 3] = labels[0, 3]           # exact zeros: sign(0) = 0
    y[1, 6] = y[1, 5]                # zero-length consistency step: sub-gradient 0
    mask = (_rand((B, T, 1), 73) > 0).expand(B, T, 4).contiguous()
    yr = y.double().requires_grad_(True)
    ref = oracle.training_loss(yr, labels.double(), mask.double(), no_labels)
    ref.backward()
    yg = y.to(cuda_device).requires_grad_(True)
    out = ops.training_loss(yg, labels.to(cuda_device), mask.to(cuda_device), no_labels)
    out[0].backward()
    assert abs(out[0].item() - ref.item()) <= 1e-6
    assert (yg.grad.cpu().double() - yr.grad).abs().max().item() <= 1e-7
