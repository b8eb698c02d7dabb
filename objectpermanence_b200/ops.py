"""Host-side operators: torch.autograd.Functions whose forward and backward enqueue
hand-written sm_100a kernels through the C ABI (include/opnet_b200.h).

PyTorch is plumbing here: it owns device memory (torch.empty), the current stream and the
autograd graph.  No arithmetic on the hot path is done by a PyTorch kernel, and there is no
CPU path -- CPU tensors raise.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib

_DEBUG_SYNC = os.environ.get("OPN_DEBUG_SYNC", "0") not in ("", "0")


def set_debug_sync(flag: bool) -> None:
    """When on, every persistent-LSTM launch is followed by a device sync + status check."""
    global _DEBUG_SYNC
    _DEBUG_SYNC = bool(flag)


PRECISIONS = {"fp32": 0, "bf16": 1, "16bit": 1}


def set_precision(mode: str) -> None:
    """Arithmetic mode of the tensor-core kernels (C ABI: opn_set_precision).  "fp32" (default): every fp32 operand is a
    hi + lo pair of 16-bit numbers, three products per slice, predicted boxes within 1e-4 of the reference.  "bf16" (alias
    "16bit"): the north star's 1e-2 mode -- one 16-bit pass with fp32 accumulation, fp32 cell state and stash; a third of
    the tensor work.  Applies to the kernels launched afterwards by this thread."""
    if mode not in PRECISIONS:
        raise ValueError(f"unknown precision {mode!r} (known: {sorted(PRECISIONS)})")
    _lib.check(_lib.load().opn_set_precision(PRECISIONS[mode]), "opn_set_precision")


def get_precision() -> str:
    return "bf16" if _lib.load().opn_get_precision() == 1 else "fp32"


class LaunchTimer:
    """Optional CUDA-event brackets around the two persistent OPNet launches, on the stream they are launched on.
    bench.py installs one for its timed region so that the roofline figure of the dominant kernel comes from the
    launches of the measured steps themselves; without one nothing is recorded."""

    def __init__(self):
        self.pairs = {}

    def bracket(self, name: str):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.pairs.setdefault(name, []).append((e0, e1))
        return e0, e1

    def mean_ms(self, name: str) -> Optional[float]:
        """Average duration of the recorded launches (call after a device synchronise)."""
        pairs = self.pairs.get(name)
        return sum(a.elapsed_time(b) for a, b in pairs) / len(pairs) if pairs else None


_launch_timer: Optional[LaunchTimer] = None


def set_launch_timer(timer: Optional[LaunchTimer]) -> None:
    global _launch_timer
    _launch_timer = timer


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "objectpermanence_b200 runs on CUDA (sm_100a) only: there is no CPU path. Move the module and its "
                "inputs to a B200 (`.to('cuda')`).")
        if t.dtype != torch.float32:
            raise RuntimeError(f"objectpermanence_b200 expects float32 tensors, got {t.dtype}")
    for t in tensors:
        if t is not None:
            status_page(t.device)   # registered before the first persistent launch on this device
            break


# ---- gradient landing slots ---------------------------------------------------------------------------------------------
# A data-parallel reducer / flat optimiser registers, per parameter, the slice of its flat gradient buffer; the backward
# functions below then write weight gradients straight into those slices (autograd adopts the tensor as p.grad when p.grad
# is None), so no concatenation pass runs before the all-reduce.
_grad_slots = {}


def register_grad_slots(params, flat: torch.Tensor) -> None:
    """`flat`: 1-D fp32 buffer holding the gradients of `params` back to back in that order."""
    off = 0
    for p in params:
        _grad_slots[p.data_ptr()] = (flat, off, tuple(p.shape))
        off += p.numel()


def unregister_grad_slots(params) -> None:
    for p in params:
        _grad_slots.pop(p.data_ptr(), None)


# Called by OPNet's backward with the gradients that are final before the rest of the backward pass has finished (the weights of
# LSTM2): a data-parallel reducer starts their all-reduce then, in the shadow of the remaining kernels.
_early_grads_hook = None


def set_early_grads_hook(fn) -> None:
    global _early_grads_hook
    _early_grads_hook = fn


def _grad_like(weight: torch.Tensor) -> torch.Tensor:
    """Uninitialised tensor for the gradient of `weight`: its registered flat-buffer slice, else a fresh allocation."""
    slot = _grad_slots.get(weight.data_ptr())
    if slot is not None and slot[2] == tuple(weight.shape) and slot[0].device == weight.device:
        flat, off, shape = slot
        return flat[off:off + weight.numel()].view(shape)
    return torch.empty_like(weight)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter: 0.3 us instead of 20 us through
    torch.cuda.current_stream(), five times per step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------
# raw kernels wrappers (no autograd)
# ------------------------------------------------------------------------------------------
def sgemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, trans_a: bool, trans_b: bool, M: int, N: int,
          K: int, lda: int, ldb: int, ldc: int, alpha: float = 1.0, beta: float = 0.0,
          bias: Optional[torch.Tensor] = None, relu: bool = False, a_off: int = 0, b_off: int = 0,
          c_off: int = 0) -> None:
    """out = alpha * op(a) op(b) + beta * out + bias.  Offsets are in elements."""
    if K == 0:
        if beta == 0.0:
            out.zero_()
        return
    lib = _lib.load()
    ws_bytes = lib.opn_sgemm_workspace_bytes(M, N, K)   # > 0: the tcgen05 split-bf16 path will be taken
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device) if ws_bytes > 0 else None
    rc = lib.opn_sgemm(int(trans_a), int(trans_b), M, N, K, alpha, a.data_ptr() + 4 * a_off, lda,
                       b.data_ptr() + 4 * b_off, ldb, beta, out.data_ptr() + 4 * c_off, ldc, _ptr(bias),
                       int(relu), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "opn_sgemm")


def _lstm_workspace(B: int, T: int, H: int, device) -> torch.Tensor:
    n = _lib.load().opn_lstm_workspace_bytes(B, T, H)
    return torch.empty(n, dtype=torch.uint8, device=device)


_status_pages = {}


def status_page(device) -> torch.Tensor:
    """The sticky status page of `device` (int32[1024], registered with the library on first use): the persistent
    kernels report inter-CTA time-outs there (word 0 = code, 1..3 = step / CTA / thread).  Callers that synchronise
    anyway (TrainingStep, inference, evaluation) copy its first 4 words along and hand them to `raise_if_failed`."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    page = _status_pages.get(idx)
    if page is None:
        page = torch.zeros(1024, dtype=torch.int32, device=torch.device("cuda", idx))
        with torch.cuda.device(idx):
            _lib.check(_lib.load().opn_set_status_page(page.data_ptr()), "opn_set_status_page")
        _status_pages[idx] = page
    return page


def raise_if_failed(words, device, what: str = "persistent kernel") -> None:
    """`words`: the first 4 status words on the HOST (after the copy has completed).  Clears the page and raises
    OpnError(OPN_ERR_TIMEOUT) when a wait expired; outputs of the affected launches are garbage."""
    code = int(words[0])
    if code != 0:
        detail = (int(words[1]), int(words[2]), int(words[3]))
        status_page(device).zero_()
        raise _lib.OpnError(f"libopnet_b200 {what} failed (code {_lib.OPN_ERR_TIMEOUT}): persistent LSTM kernel timed out "
                            f"(status {code}, step {detail[0]}, cta {detail[1]}, thread {detail[2]}); the results of "
                            "this step are invalid")


def check_status(device, what: str = "persistent kernel") -> None:
    """Blocking form: synchronise the current stream and check the status page of `device`."""
    page = status_page(device)
    words = page[:4].cpu()      # stream-ordered copy + host synchronisation
    raise_if_failed(words, device, what)


def _lstm_check(device, what: str) -> None:
    if _DEBUG_SYNC:
        check_status(device, what)


def dropout_(x: torch.Tensor, out: torch.Tensor, p: float, seed: int, offset: int) -> None:
    """out = x * keep / (1 - p) with the counter-based mask of opn_dropout (out may be x)."""
    rc = _lib.load().opn_dropout(x.numel(), x.data_ptr(), out.data_ptr(), p, seed, offset, _stream())
    _lib.check(rc, "opn_dropout")


class DropoutStream:
    """Hands out (seed, offset) pairs for dropout sites.  Every site draws a fresh 62-bit Philox key from torch's
    default CPU generator (a host-side draw, no device work), so masks follow torch.manual_seed() the way the
    reference's nn.Dropout masks do, and no two sites share mask words."""

    def take(self, n_elements: int) -> Tuple[int, int]:
        return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item()), 0


dropout_stream = DropoutStream()


# ------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------
class DropoutFn(torch.autograd.Function):
    """nn.Dropout of the encoder layer in train mode (dropout / dropout1 / dropout2 of
    nn.TransformerEncoderLayer, baselines/learned_models.py:166).  The mask is regenerated in backward."""

    @staticmethod
    def forward(ctx, x, p: float, seed: int, offset: int):
        _require_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        dropout_(x, y, p, seed, offset)
        ctx.args = (p, seed, offset)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        dropout_(dy, dx, *ctx.args)
        return dx, None, None, None


class LinearFn(torch.autograd.Function):
    """y = x W^T (+ bias) (ReLU).  x [..., K] contiguous, W [N, K].  nn.Linear call sites of
    baselines/learned_models.py (:30,:33,:67,:69,:70,:102,:130,:133,:167,:172)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool):
        _require_cuda(x, weight, bias)
        x = x.contiguous()
        weight = weight.contiguous()
        K = x.shape[-1]
        N = weight.shape[0]
        M = x.numel() // K
        y = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32)
        sgemm(x, weight, y, trans_a=False, trans_b=True, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bias, relu=relu)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.has_bias = bias is not None
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        K = x.shape[-1]
        N = weight.shape[0]
        M = x.numel() // K
        dy = dy.contiguous()
        if ctx.relu:
            dy = dy.clone()
            _lib.check(_lib.load().opn_relu_bwd(dy.numel(), y.data_ptr(), dy.data_ptr(), _stream()), "opn_relu_bwd")
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            sgemm(dy, weight, dx, trans_a=False, trans_b=False, M=M, N=K, K=N, lda=N, ldb=K, ldc=K)
        if ctx.needs_input_grad[1]:
            dw = _grad_like(weight)
            if wgrad_tc_wanted(M, N) and K >= 64:
                # dW = dy^T x contracts both operands over their rows, like the LSTM weight gradients: same kernel
                wgrad_jobs_run([(dy.reshape(M, N), x.reshape(M, K), dw, M, 0)])
            else:
                sgemm(dy, x, dw, trans_a=True, trans_b=False, M=N, N=K, K=M, lda=N, ldb=K, ldc=K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(N, device=x.device, dtype=torch.float32)
            _lib.check(_lib.load().opn_colsum(M, N, dy.data_ptr(), N, db.data_ptr(), 0, _stream()), "opn_colsum")
        return dx, dw, db, None


class SlotLinearReluFn(torch.autograd.Function):
    """relu(boxes[:, :, slot, :] W^T) without materialising the slice: the snitch-slot rows of
    boxes [B,T,15,F] are read in place with a row stride of 15*F.  transformer_lstm keeps only
    slot 0 of the encoder output (baselines/learned_models.py:178,185); the other 14 slots never
    reach the output or any gradient, so they are not computed."""

    @staticmethod
    def forward(ctx, boxes, weight, slot: int):
        _require_cuda(boxes, weight)
        boxes = boxes.contiguous()
        weight = weight.contiguous()
        B, T, NO, F = boxes.shape
        D = weight.shape[0]
        y = torch.empty(B, T, D, device=boxes.device, dtype=torch.float32)
        sgemm(boxes, weight, y, trans_a=False, trans_b=True, M=B * T, N=D, K=F, lda=NO * F, ldb=F, ldc=D, relu=True,
              a_off=slot * F)
        ctx.save_for_backward(boxes, weight, y)
        ctx.slot = slot
        return y

    @staticmethod
    def backward(ctx, dy):
        boxes, weight, y = ctx.saved_tensors
        B, T, NO, F = boxes.shape
        D = weight.shape[0]
        dy = dy.contiguous().clone()
        _lib.check(_lib.load().opn_relu_bwd(dy.numel(), y.data_ptr(), dy.data_ptr(), _stream()), "opn_relu_bwd")
        dw = torch.empty_like(weight)
        sgemm(dy, boxes, dw, trans_a=True, trans_b=False, M=D, N=F, K=B * T, lda=D, ldb=NO * F, ldc=F,
              b_off=ctx.slot * F)
        return None, dw, None


class LstmLayerFn(torch.autograd.Function):
    """One bias-free LSTM layer (batch-first, zero initial state): the persistent recurrence
    kernels plus the time-parallel input projection / gradient contractions."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, stash: bool = True):
        _require_cuda(x, w_ih, w_hh)
        x = x.contiguous()
        w_ih = w_ih.contiguous()
        w_hh = w_hh.contiguous()
        B, T, I = x.shape
        H = w_hh.shape[1]
        lib = _lib.load()
        dev = x.device
        xproj = torch.empty(B, T, 4 * H, device=dev, dtype=torch.float32)
        sgemm(x, w_ih, xproj, trans_a=False, trans_b=True, M=B * T, N=4 * H, K=I, lda=I, ldb=I, ldc=4 * H)
        hs = torch.empty(B, T, H, device=dev, dtype=torch.float32)
        # `stash` comes from the front-end (lstm_layer): needs_input_grad mirrors requires_grad whatever the grad mode,
        # and grad mode is always off inside Function.forward, so torch.no_grad() has to be seen by the caller
        need_grad = stash and any(ctx.needs_input_grad)
        gates = cells = None
        if need_grad:
            gates = torch.empty(B, T, 4 * H, device=dev, dtype=torch.float32)
            cells = torch.empty(B, T, H, device=dev, dtype=torch.float32)
        ws = _lstm_workspace(B, T, H, dev)
        rc = lib.opn_lstm_fwd(B, T, H, xproj.data_ptr(), w_hh.data_ptr(), hs.data_ptr(), _ptr(gates), _ptr(cells),
                              ws.data_ptr(), ws.numel(), _stream())
        _lib.check(rc, "opn_lstm_fwd")
        _lstm_check(dev, "opn_lstm_fwd")
        if need_grad:
            ctx.save_for_backward(x, w_ih, w_hh, hs, gates, cells)
        return hs

    @staticmethod
    def backward(ctx, dhs):
        x, w_ih, w_hh, hs, gates, cells = ctx.saved_tensors
        return _lstm_backward(x, w_ih, w_hh, hs, gates, cells, dhs, ctx.needs_input_grad) + (None,)


def _lstm_recurrence_backward(w_hh, gates, cells, dhs):
    """Reverse recurrence of one LSTM layer -> dgates [B,T,4H] = dLoss / d(pre-activation gates)."""
    B, T, H4 = gates.shape
    H = H4 // 4
    lib = _lib.load()
    dhs = dhs.contiguous()
    dgates = torch.empty(B, T, 4 * H, device=gates.device, dtype=torch.float32)
    ws = _lstm_workspace(B, T, H, gates.device)
    rc = lib.opn_lstm_bwd(B, T, H, w_hh.data_ptr(), gates.data_ptr(), cells.data_ptr(), dhs.data_ptr(),
                          dgates.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "opn_lstm_bwd")
    _lstm_check(gates.device, "opn_lstm_bwd")
    return dgates


def wgrad_jobs_run(jobs) -> None:
    """One launch of opn_wgrad over `jobs` = [(a [rows, M], b [rows, N], out [M, N], T, shift[, trans_out]), ...]: out = sum_r
    a[r]^T b[r - shift] (shift 1: rows with r % T == 0 excluded; trans_out: `out` is [N, M] and receives the transpose).
    The tcgen05 weight-gradient kernel of csrc/opn_wgrad_tc.cu."""
    lib = _lib.load()
    arr = (_lib.WgradJob * len(jobs))()
    for j, job in enumerate(jobs):
        a, b, out, T, shift = job[:5]
        trans_out = bool(job[5]) if len(job) > 5 else False      # out given as [N, M]: written transposed
        rows, M = a.shape
        N = b.shape[1]
        assert b.shape[0] == rows and out.shape == ((N, M) if trans_out else (M, N))
        assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
        arr[j] = _lib.WgradJob(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.stride(0), b.stride(0), out.stride(0), rows, T, M, N,
                               int(shift), int(trans_out))
    n = lib.opn_wgrad_workspace_bytes(len(jobs), arr)
    if n <= 0:
        _lib.check(-1, "opn_wgrad_workspace_bytes")
    ws = torch.empty(n, dtype=torch.uint8, device=jobs[0][0].device)
    rc = lib.opn_wgrad(len(jobs), arr, ws.data_ptr(), n, _stream())
    _lib.check(rc, "opn_wgrad")
    _lstm_check(jobs[0][0].device, "opn_wgrad")


def wgrad_tc_wanted(rows: int, M: int) -> bool:
    """The tcgen05 weight-gradient kernel takes the LSTM weight gradients from 1024 rows on (below that the exact-fp32
    FFMA contraction of opn_sgemm is as fast); OPN_WGRAD=sgemm keeps the general path (A/B runs, tests)."""
    return rows >= 1024 and M % 128 == 0 and os.environ.get("OPN_WGRAD", "tc") != "sgemm"


def _lstm_wgrad_jobs(dgates, x, hs, w_ih, w_hh, need_dw_ih, need_dw_hh):
    """-> (jobs for wgrad_jobs_run, dW_ih, dW_hh) of one LSTM layer."""
    B, T, I = x.shape
    H = w_hh.shape[1]
    jobs, dw_ih, dw_hh = [], None, None
    dg = dgates.reshape(B * T, 4 * H)
    if need_dw_ih:
        dw_ih = _grad_like(w_ih)
        jobs.append((dg, x.reshape(B * T, I), dw_ih, T, 0))
    if need_dw_hh:
        dw_hh = _grad_like(w_hh)
        jobs.append((dg, hs.reshape(B * T, H), dw_hh, T, 1))
    return jobs, dw_ih, dw_hh


def _lstm_weight_grads(dgates, x, hs, w_ih, w_hh, need_dw_ih, need_dw_hh):
    """The two time-parallel weight-gradient contractions of one LSTM layer -> (dW_ih, dW_hh)."""
    B, T, I = x.shape
    H = w_hh.shape[1]
    if wgrad_tc_wanted(B * T, 4 * H) and (need_dw_ih or need_dw_hh):
        jobs, dw_ih, dw_hh = _lstm_wgrad_jobs(dgates.contiguous(), x.contiguous(), hs.contiguous(), w_ih, w_hh, need_dw_ih, need_dw_hh)
        wgrad_jobs_run(jobs)
        return dw_ih, dw_hh
    dw_ih = dw_hh = None
    if need_dw_ih:
        dw_ih = _grad_like(w_ih)
        sgemm(dgates, x, dw_ih, trans_a=True, trans_b=False, M=4 * H, N=I, K=B * T, lda=4 * H, ldb=I, ldc=I)
    if need_dw_hh:
        dw_hh = _grad_like(w_hh)
        if T > 1:
            # dW_hh = sum_{b, t>=1} dgates[b,t]^T hs[b,t-1].  Contract the flat row pairs (r+1, r) over all
            # B*T-1 rows, then take out the B-1 pairs that straddle two videos (rows b*T and b*T-1).
            sgemm(dgates, hs, dw_hh, trans_a=True, trans_b=False, M=4 * H, N=H, K=B * T - 1, lda=4 * H, ldb=H,
                  ldc=H, a_off=4 * H)
            if B > 1:
                sgemm(dgates, hs, dw_hh, trans_a=True, trans_b=False, M=4 * H, N=H, K=B - 1, lda=T * 4 * H,
                      ldb=T * H, ldc=H, alpha=-1.0, beta=1.0, a_off=T * 4 * H, b_off=(T - 1) * H)
        else:
            dw_hh.zero_()
    return dw_ih, dw_hh


def _lstm_backward(x, w_ih, w_hh, hs, gates, cells, dhs, needs):
    """Reverse recurrence + the three time-parallel contractions of one LSTM layer -> (dx, dW_ih, dW_hh)."""
    B, T, I = x.shape
    H = w_hh.shape[1]
    dgates = _lstm_recurrence_backward(w_hh, gates, cells, dhs)
    dx = None
    if needs[0]:
        dx = torch.empty_like(x)
        sgemm(dgates, w_ih, dx, trans_a=False, trans_b=False, M=B * T, N=I, K=4 * H, lda=4 * H, ldb=I, ldc=I)
    dw_ih, dw_hh = _lstm_weight_grads(dgates, x, hs, w_ih, w_hh, needs[1], needs[2])
    return dx, dw_ih, dw_hh


_side_streams = {}


def _side_stream(device) -> "torch.cuda.Stream":
    """One auxiliary stream per device for work that runs beside a recurrence kernel."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


class WhoToTrackFn(torch.autograd.Function):
    """OPNet's soft object selection (baselines/learned_models.py:40-43,50)."""

    @staticmethod
    def forward(ctx, boxes, hs1, w_pred):
        _require_cuda(boxes, hs1, w_pred)
        boxes = boxes.contiguous()
        hs1 = hs1.contiguous()
        w_pred = w_pred.contiguous()
        B, T, NO, F = boxes.shape
        if NO != 15 or F != 6 or w_pred.shape[0] != 15:
            raise RuntimeError("who-to-track expects boxes [B,T,15,6] and object_to_track_pred_dim == 15 "
                               "(the reference einsum requires it, baselines/learned_models.py:43)")
        H1 = hs1.shape[-1]
        dev = boxes.device
        logits = torch.empty(B, 15, T, device=dev, dtype=torch.float32)
        probs = torch.empty(B, T, 15, device=dev, dtype=torch.float32)
        fb = torch.empty(B, T, 6, device=dev, dtype=torch.float32)
        rc = _lib.load().opn_wtt_fwd(B, T, H1, boxes.data_ptr(), hs1.data_ptr(), w_pred.data_ptr(), logits.data_ptr(),
                                     probs.data_ptr(), fb.data_ptr(), _stream())
        _lib.check(rc, "opn_wtt_fwd")
        ctx.save_for_backward(boxes, hs1, w_pred, probs)
        return fb, logits

    @staticmethod
    def backward(ctx, dfb, dlogits):
        boxes, hs1, w_pred, probs = ctx.saved_tensors
        dhs1, dw = _wtt_backward(boxes, hs1, w_pred, probs, dfb, dlogits, ctx.needs_input_grad[2])
        return None, dhs1, dw


def _wtt_backward_kernel(boxes, hs1, w_pred, probs, dfb, dlogits):
    """who-to-track backward -> (d hs1, d logits [B,T,15] in row layout)."""
    B, T, NO, F = boxes.shape
    H1 = hs1.shape[-1]
    dev = boxes.device
    if dfb is None:
        dfb = torch.zeros(B, T, 6, device=dev, dtype=torch.float32)
    dfb = dfb.contiguous()
    dl_up = dlogits.contiguous() if dlogits is not None else None
    dl = torch.empty(B, T, 15, device=dev, dtype=torch.float32)
    dhs1 = torch.empty_like(hs1)
    rc = _lib.load().opn_wtt_bwd(B, T, H1, boxes.data_ptr(), probs.data_ptr(), w_pred.data_ptr(), dfb.data_ptr(),
                                 _ptr(dl_up), dl.data_ptr(), dhs1.data_ptr(), _stream())
    _lib.check(rc, "opn_wtt_bwd")
    return dhs1, dl


def _wtt_weight_grad(hs1, dl):
    """dW_pred^T [H1,15] = hs1^T dl: contracted in the transposed orientation so that the long dimension (H1) maps
    to the 128-row tile and the 15 objects to the 16-wide one (as M=15 it ran 7x slower)."""
    B, T, H1 = hs1.shape
    dw_t = torch.empty(H1, 15, device=hs1.device, dtype=torch.float32)
    sgemm(hs1, dl, dw_t, trans_a=True, trans_b=False, M=H1, N=15, K=B * T, lda=H1, ldb=15, ldc=15)
    return dw_t.t()


def _wtt_backward(boxes, hs1, w_pred, probs, dfb, dlogits, need_dw):
    """who-to-track backward -> (d hs1, dW_pred)."""
    dhs1, dl = _wtt_backward_kernel(boxes, hs1, w_pred, probs, dfb, dlogits)
    return dhs1, (_wtt_weight_grad(hs1, dl) if need_dw else None)


class OPNetTrunkFn(torch.autograd.Function):
    """OPNet up to the bbox head (baselines/learned_models.py:36-46) with the forward as ONE persistent kernel:
    LSTM1, the who-to-track head and LSTM2 advance frame by frame together (csrc/opn_opnet_fused.cu), each hiding
    the other's exchange latency.  Shipped config only (H1 = 256, H2 = 512).  The stash it leaves is that of the
    separate kernels, so the backward pass is theirs: reverse LSTM2, who-to-track backward, reverse LSTM1 and the
    time-parallel weight-gradient contractions."""

    @staticmethod
    def forward(ctx, boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2, stash: bool = True):
        _require_cuda(boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2)
        boxes = boxes.contiguous()
        w_ih1, w_hh1, w_pred, w_ih2, w_hh2 = [w.contiguous() for w in (w_ih1, w_hh1, w_pred, w_ih2, w_hh2)]
        B, T, NO, F = boxes.shape
        H1, H2 = w_hh1.shape[1], w_hh2.shape[1]
        lib = _lib.load()
        dev = boxes.device
        f32 = dict(device=dev, dtype=torch.float32)
        xproj1 = torch.empty(B, T, 4 * H1, **f32)
        sgemm(boxes, w_ih1, xproj1, trans_a=False, trans_b=True, M=B * T, N=4 * H1, K=NO * F, lda=NO * F, ldb=NO * F,
              ldc=4 * H1)
        need_grad = stash and any(ctx.needs_input_grad)   # see LstmLayerFn.forward
        ctx.set_materialize_grads(False)   # backward sees None (not zeros) when an output carries no gradient
        hs1 = torch.empty(B, T, H1, **f32)
        hs2 = torch.empty(B, T, H2, **f32)
        logits = torch.empty(B, 15, T, **f32)
        probs = torch.empty(B, T, 15, **f32)
        fb = torch.empty(B, T, 6, **f32)
        gates1 = cells1 = gates2 = cells2 = None
        if need_grad:
            gates1 = torch.empty(B, T, 4 * H1, **f32)
            cells1 = torch.empty(B, T, H1, **f32)
            gates2 = torch.empty(B, T, 4 * H2, **f32)
            cells2 = torch.empty(B, T, H2, **f32)
        ws = torch.empty(lib.opn_opnet_fwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
        timed = _launch_timer.bracket("opnet_fwd_fused") if _launch_timer is not None else None
        if timed:
            timed[0].record()
        rc = lib.opn_opnet_fwd(B, T, H1, H2, boxes.data_ptr(), xproj1.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(),
                               w_ih2.data_ptr(), w_hh2.data_ptr(), hs1.data_ptr(), _ptr(gates1), _ptr(cells1),
                               logits.data_ptr(), probs.data_ptr(), fb.data_ptr(), hs2.data_ptr(), _ptr(gates2),
                               _ptr(cells2), ws.data_ptr(), ws.numel(), _stream())
        if timed:
            timed[1].record()
        _lib.check(rc, "opn_opnet_fwd")
        _lstm_check(dev, "opn_opnet_fwd")
        if need_grad:
            ctx.save_for_backward(boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2, hs1, gates1, cells1, probs, fb, hs2,
                                  gates2, cells2)
        return hs2, logits

    @staticmethod
    def backward(ctx, dhs2, dlogits):
        (boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2, hs1, gates1, cells1, probs, fb, hs2, gates2,
         cells2) = ctx.saved_tensors
        B, T = boxes.shape[:2]
        need = ctx.needs_input_grad
        if dhs2 is None:
            dhs2 = torch.zeros_like(hs2)
        H2 = w_hh2.shape[1]
        if dlogits is None and os.environ.get("OPN_OPNET_FUSED_BWD", "1") not in ("0", ""):
            # both reverse recurrences and the who-to-track backward in one persistent kernel (opn_opnet_fused_bwd.cu)
            lib = _lib.load()
            H1 = w_hh1.shape[1]
            dev = boxes.device
            dgates2 = torch.empty(B, T, 4 * H2, device=dev, dtype=torch.float32)
            dgates1 = torch.empty(B, T, 4 * H1, device=dev, dtype=torch.float32)
            dl = torch.empty(B, T, 15, device=dev, dtype=torch.float32)
            ws = torch.empty(lib.opn_opnet_bwd_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
            dhs2 = dhs2.contiguous()
            timed = _launch_timer.bracket("opnet_bwd_fused") if _launch_timer is not None else None
            if timed:
                timed[0].record()
            x1 = boxes.reshape(B, T, -1)
            early = wgrad_tc_wanted(B * T, 4 * H1) and all(need[1:6]) and os.environ.get("OPN_WGRAD_EARLY", "1") not in ("0", "")
            entry = lib.opn_opnet_bwd_begin if early else lib.opn_opnet_bwd
            rc = entry(B, T, H1, H2, boxes.data_ptr(), probs.data_ptr(), w_hh1.data_ptr(), w_pred.data_ptr(),
                       w_ih2.data_ptr(), w_hh2.data_ptr(), gates1.data_ptr(), cells1.data_ptr(),
                       gates2.data_ptr(), cells2.data_ptr(), dhs2.data_ptr(), dgates1.data_ptr(),
                       dgates2.data_ptr(), dl.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
            _lib.check(rc, "opn_opnet_bwd")
            if early:
                # The LSTM2 loop (d_gates2) finishes ~0.15 ms before the head / LSTM1 kernel beside it: its two weight gradients
                # run on the 128 SMs it has freed while that kernel completes, then the join, then the other three products
                j2, dw_ih2, dw_hh2 = _lstm_wgrad_jobs(dgates2, fb, hs2, w_ih2, w_hh2, True, True)
                wgrad_jobs_run(j2)
                if _early_grads_hook is not None:
                    _early_grads_hook((dw_ih2, dw_hh2))
                _lib.check(lib.opn_opnet_bwd_join(_stream()), "opn_opnet_bwd_join")
                if timed:
                    timed[1].record()
                _lstm_check(dev, "opn_opnet_bwd")
                j1, dw_ih1, dw_hh1 = _lstm_wgrad_jobs(dgates1, x1, hs1, w_ih1, w_hh1, True, True)
                dw_pred = _grad_like(w_pred)
                wgrad_jobs_run(j1 + [(hs1.reshape(B * T, H1), dl.reshape(B * T, 15), dw_pred, T, 0, True)])
                return None, dw_ih1, dw_hh1, dw_pred, dw_ih2, dw_hh2, None
            if timed:
                timed[1].record()
            _lstm_check(dev, "opn_opnet_bwd")
            if wgrad_tc_wanted(B * T, 4 * H1) and all(need[1:6]):
                # all five weight gradients as ONE launch of the tcgen05 weight-gradient kernel (+ its reduction);
                # dW_pred in the transposed orientation (H1 = the 128-wide dimension)
                j2, dw_ih2, dw_hh2 = _lstm_wgrad_jobs(dgates2, fb, hs2, w_ih2, w_hh2, True, True)
                j1, dw_ih1, dw_hh1 = _lstm_wgrad_jobs(dgates1, x1, hs1, w_ih1, w_hh1, True, True)
                dw_pred = _grad_like(w_pred)
                wgrad_jobs_run(j2 + j1 + [(hs1.reshape(B * T, H1), dl.reshape(B * T, 15), dw_pred, T, 0, True)])
                return None, dw_ih1, dw_hh1, dw_pred, dw_ih2, dw_hh2, None
            if os.environ.get("OPN_OPNET_WGRAD_OVERLAP", "1") not in ("0", "") and all(need[1:6]):
                # The five weight-gradient contractions are independent of each other and none fills the GPU (pre-pass,
                # 16-64 output tiles, split-K): LSTM2's stay on the main stream, LSTM1's and dW_pred run beside them
                # (2.61 -> 2.58 ms per step, tools/step_ab.py).  OPN_OPNET_WGRAD_OVERLAP=0: all in line.
                main, side = torch.cuda.current_stream(), _side_stream(dev)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    dw_ih1, dw_hh1 = _lstm_weight_grads(dgates1, x1, hs1, w_ih1, w_hh1, True, True)
                    dw_pred = _wtt_weight_grad(hs1, dl)
                dw_ih2, dw_hh2 = _lstm_weight_grads(dgates2, fb, hs2, w_ih2, w_hh2, True, True)
                for t_ in (dgates1, boxes, hs1, dl):
                    t_.record_stream(side)
                done = torch.cuda.Event()
                done.record(side)
                main.wait_event(done)
                for t_ in (dw_ih1, dw_hh1, dw_pred):
                    t_.record_stream(main)
                return None, dw_ih1, dw_hh1, dw_pred, dw_ih2, dw_hh2, None
            dw_ih2, dw_hh2 = _lstm_weight_grads(dgates2, fb, hs2, w_ih2, w_hh2, need[4], need[5])
            dw_pred = _wtt_weight_grad(hs1, dl) if need[3] else None
            dw_ih1, dw_hh1 = _lstm_weight_grads(dgates1, x1, hs1, w_ih1, w_hh1, need[1], need[2])
            return None, dw_ih1, dw_hh1, dw_pred, dw_ih2, dw_hh2, None
        dgates2 = _lstm_recurrence_backward(w_hh2, gates2, cells2, dhs2)
        dfb = torch.empty_like(fb)
        sgemm(dgates2, w_ih2, dfb, trans_a=False, trans_b=False, M=B * T, N=6, K=4 * H2, lda=4 * H2, ldb=6, ldc=6)
        # The weight gradients of LSTM2 depend only on dgates2: they run on a side stream beside the who-to-track
        # backward and the LSTM1 reverse recurrence, which occupies 64 of the 148 SMs (OPN_OPNET_OVERLAP=0: in line).
        overlap = os.environ.get("OPN_OPNET_OVERLAP", "1") not in ("0", "") and (need[3] or need[4] or need[5])
        if overlap:
            main, side = torch.cuda.current_stream(), _side_stream(boxes.device)
            ready = torch.cuda.Event()
            ready.record(main)     # dgates2 and d frames_boxes are complete
        dhs1, dl = _wtt_backward_kernel(boxes, hs1, w_pred, probs, dfb, dlogits)
        if overlap:
            ready_dl = torch.cuda.Event()
            ready_dl.record(main)
        x1 = boxes.reshape(B, T, -1)
        dgates1 = _lstm_recurrence_backward(w_hh1, gates1, cells1, dhs1)     # enqueued first: takes its 64 SMs
        if overlap:
            with torch.cuda.stream(side):
                side.wait_event(ready)
                dw_ih2, dw_hh2 = _lstm_weight_grads(dgates2, fb, hs2, w_ih2, w_hh2, need[4], need[5])
                side.wait_event(ready_dl)
                dw_pred = _wtt_weight_grad(hs1, dl) if need[3] else None
            for t_ in (dgates2, fb, hs2, hs1, dl):
                t_.record_stream(side)
            done = torch.cuda.Event()
            done.record(side)
            main.wait_event(done)
            for t_ in (dw_ih2, dw_hh2, dw_pred):
                if t_ is not None:
                    t_.record_stream(main)
        else:
            dw_ih2, dw_hh2 = _lstm_weight_grads(dgates2, fb, hs2, w_ih2, w_hh2, need[4], need[5])
            dw_pred = _wtt_weight_grad(hs1, dl) if need[3] else None
        dw_ih1, dw_hh1 = _lstm_weight_grads(dgates1, x1, hs1, w_ih1, w_hh1, need[1], need[2])
        return None, dw_ih1, dw_hh1, dw_pred, dw_ih2, dw_hh2, None


class AddLayerNormFn(torch.autograd.Function):
    """LayerNorm(x + res) (post-norm residual blocks of nn.TransformerEncoderLayer)."""

    @staticmethod
    def forward(ctx, x, res, weight, bias, eps: float):
        _require_cuda(x, res, weight, bias)
        x = x.contiguous()
        res = res.contiguous()
        D = x.shape[-1]
        rows = x.numel() // D
        y = torch.empty_like(x)
        xhat = torch.empty_like(x)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        rc = _lib.load().opn_layernorm_fwd(rows, D, x.data_ptr(), res.data_ptr(), weight.data_ptr(), bias.data_ptr(),
                                           eps, y.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), _stream())
        _lib.check(rc, "opn_layernorm_fwd")
        ctx.save_for_backward(xhat, rstd, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        xhat, rstd, weight = ctx.saved_tensors
        D = xhat.shape[-1]
        rows = xhat.numel() // D
        dy = dy.contiguous()
        dx = torch.empty_like(xhat)
        dw = torch.zeros_like(weight)
        db = torch.zeros_like(weight)
        rc = _lib.load().opn_layernorm_bwd(rows, D, xhat.data_ptr(), rstd.data_ptr(), weight.data_ptr(), dy.data_ptr(),
                                           dx.data_ptr(), dw.data_ptr(), db.data_ptr(), _stream())
        _lib.check(rc, "opn_layernorm_bwd")
        return dx, dx, dw, db, None


class FusedSelfAttentionFn(torch.autograd.Function):
    """softmax(Q K^T / sqrt(d)) V per head over ONE sequence of S rows, head dimension 128, as fused flash-style tcgen05
    kernels (csrc/opn_attention_tc.cu): no [S,S] tensor, the backward pass recomputes the scores from the row statistics
    kept in the workspace.  Same arguments and dropout stream as SelfAttentionFn (the materialised form, kept for other
    head dimensions and for A/B: OPN_ATTENTION=materialized)."""

    @staticmethod
    def forward(ctx, qkv, nhead: int, p_drop: float = 0.0, seed: int = 0, offset: int = 0):
        _require_cuda(qkv)
        qkv = qkv.contiguous()
        S, D3 = qkv.shape
        D = D3 // 3
        lib = _lib.load()
        out = torch.empty(S, D, device=qkv.device, dtype=torch.float32)
        ws = torch.empty(lib.opn_attention_workspace_bytes(S, D, nhead), dtype=torch.uint8, device=qkv.device)
        rc = lib.opn_attention_fwd(S, D, nhead, qkv.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), float(p_drop),
                                   int(seed), int(offset), _stream())
        _lib.check(rc, "opn_attention_fwd")
        _lstm_check(qkv.device, "opn_attention_fwd")
        ctx.save_for_backward(out, ws)
        ctx.args = (S, D, nhead, float(p_drop), int(seed), int(offset))
        return out

    @staticmethod
    def backward(ctx, dctx):
        out, ws = ctx.saved_tensors
        S, D, nhead, p_drop, seed, offset = ctx.args
        dctx = dctx.contiguous()
        dqkv = torch.empty(S, 3 * D, device=out.device, dtype=torch.float32)
        rc = _lib.load().opn_attention_bwd(S, D, nhead, out.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(), ws.data_ptr(),
                                           ws.numel(), p_drop, seed, offset, _stream())
        _lib.check(rc, "opn_attention_bwd")
        _lstm_check(out.device, "opn_attention_bwd")
        return dqkv, None, None, None, None


class SelfAttentionFn(torch.autograd.Function):
    """softmax(Q K^T / sqrt(d)) V per head over ONE sequence of S rows.

    qkv [S, 3D] packed q|k|v as produced by in_proj (nn.MultiheadAttention ordering), nhead heads
    of d = D/nhead.  The sequence axis of the reference is B*T (baselines/learned_models.py:183-184
    hands a (B*T, 15, D) tensor to a sequence-first encoder), so S = B*T here.  The score matrix is
    materialised per head in HBM ([S,S] fp32; 369 MB at S = 9600) and kept for the backward pass."""

    @staticmethod
    def forward(ctx, qkv, nhead: int, p_drop: float = 0.0, seed: int = 0, offset: int = 0):
        _require_cuda(qkv)
        qkv = qkv.contiguous()
        S, D3 = qkv.shape
        D = D3 // 3
        d = D // nhead
        scale = 1.0 / (d ** 0.5)
        dev = qkv.device
        lib = _lib.load()
        ctx_out = torch.empty(S, D, device=dev, dtype=torch.float32)
        probs = torch.empty(nhead, S, S, device=dev, dtype=torch.float32)
        # train mode: nn.MultiheadAttention drops attention weights after the softmax; the dropped copy feeds P V
        # (and dV in backward), the softmax backward needs the undropped one
        dropped = torch.empty_like(probs) if p_drop > 0.0 else None
        head_blocks = (S * S + 3) // 4
        for h in range(nhead):
            p = probs[h]
            # scores = q_h k_h^T
            sgemm(qkv, qkv, p, trans_a=False, trans_b=True, M=S, N=S, K=d, lda=D3, ldb=D3, ldc=S, a_off=h * d,
                  b_off=D + h * d)
            _lib.check(lib.opn_softmax_rows(S, S, p.data_ptr(), S, scale, _stream()), "opn_softmax_rows")
            if dropped is not None:
                dropout_(p, dropped[h], p_drop, seed, offset + h * head_blocks)
                p = dropped[h]
            # ctx_h = P v_h
            sgemm(p, qkv, ctx_out, trans_a=False, trans_b=False, M=S, N=d, K=S, lda=S, ldb=D3, ldc=D,
                  b_off=2 * D + h * d, c_off=h * d)
        ctx.save_for_backward(qkv, probs, dropped)
        ctx.nhead = nhead
        ctx.drop = (p_drop, seed, offset)
        return ctx_out

    @staticmethod
    def backward(ctx, dctx):
        qkv, probs, dropped = ctx.saved_tensors
        nhead = ctx.nhead
        p_drop, seed, offset = ctx.drop
        S, D3 = qkv.shape
        D = D3 // 3
        d = D // nhead
        scale = 1.0 / (d ** 0.5)
        dev = qkv.device
        lib = _lib.load()
        dctx = dctx.contiguous()
        dqkv = torch.empty_like(qkv)
        dp = torch.empty(S, S, device=dev, dtype=torch.float32)
        head_blocks = (S * S + 3) // 4
        for h in range(nhead):
            p = probs[h]
            # dV_h = P^T dctx_h  (the dropped P in train mode)
            sgemm(p if dropped is None else dropped[h], dctx, dqkv, trans_a=True, trans_b=False, M=S, N=d, K=S, lda=S,
                  ldb=D, ldc=D3, b_off=h * d, c_off=2 * D + h * d)
            # dP = dctx_h V_h^T
            sgemm(dctx, qkv, dp, trans_a=False, trans_b=True, M=S, N=S, K=d, lda=D, ldb=D3, ldc=S, a_off=h * d,
                  b_off=2 * D + h * d)
            if dropped is not None:
                dropout_(dp, dp, p_drop, seed, offset + h * head_blocks)
            # dS = scale * P * (dP - rowsum(P dP))
            _lib.check(lib.opn_softmax_rows_bwd(S, S, p.data_ptr(), dp.data_ptr(), S, scale, _stream()),
                       "opn_softmax_rows_bwd")
            # dQ_h = dS K_h ; dK_h = dS^T Q_h
            sgemm(dp, qkv, dqkv, trans_a=False, trans_b=False, M=S, N=d, K=S, lda=S, ldb=D3, ldc=D3, b_off=D + h * d,
                  c_off=h * d)
            sgemm(dp, qkv, dqkv, trans_a=True, trans_b=False, M=S, N=d, K=S, lda=S, ldb=D3, ldc=D3, b_off=h * d,
                  c_off=D + h * d)
        return dqkv, None, None, None, None


def loss_and_grad(y, labels, mask=None, no_labels: bool = False):
    """The loss of baselines/training_main.py:192-210 and d total / dy in one launch, outside autograd:
    (3-vector (total, prediction, consistency), dy [B,T,4]).  `y.backward(dy)` continues into the model."""
    _require_cuda(y, labels)
    y = y.detach().contiguous()
    labels = labels.contiguous()
    B, T, C = y.shape
    if C != 4:
        raise RuntimeError("training loss expects [B,T,4] predictions")
    m = None
    if no_labels:
        if mask is None:
            raise RuntimeError("*_no_labels models need the visibility mask")
        m = mask.to(torch.uint8).contiguous()
    out = torch.empty(3, device=y.device, dtype=torch.float32)
    dy = torch.empty_like(y)
    rc = _lib.load().opn_loss_fwd_bwd(B, T, y.data_ptr(), labels.data_ptr(), _ptr(m), int(no_labels),
                                      out.data_ptr(), dy.data_ptr(), _stream())
    _lib.check(rc, "opn_loss_fwd_bwd")
    return out, dy


def head_loss_available(h: torch.Tensor, weight: torch.Tensor, bias) -> bool:
    """The fused bbox head + loss pass (opn_head_loss) exists for bias-free heads [4, H] with H a multiple of 128 up to 512;
    OPN_HEAD_LOSS=0 keeps the separate launches."""
    H = weight.shape[1]
    return (bias is None and weight.shape[0] == 4 and H % 128 == 0 and 128 <= H <= 512 and h.dim() == 3 and h.shape[2] == H
            and os.environ.get("OPN_HEAD_LOSS", "1") not in ("0", ""))


def head_loss(h, weight, labels, mask=None, no_labels: bool = False):
    """Bbox head, training loss and the backward of both in one launch, outside autograd:
    -> (y [B,T,4], 3-vector (total, prediction, consistency), d total / d h [B,T,H], d total / d weight [4,H]).
    `h.backward(dh)` continues into the model; dW lands in the weight's registered gradient slot if it has one."""
    _require_cuda(h, weight, labels)
    h = h.detach().contiguous()
    w = weight.detach().contiguous()
    labels = labels.contiguous()
    B, T, H = h.shape
    m = None
    if no_labels:
        if mask is None:
            raise RuntimeError("*_no_labels models need the visibility mask")
        m = mask.to(torch.uint8).contiguous()
    lib = _lib.load()
    y = torch.empty(B, T, 4, device=h.device, dtype=torch.float32)
    out = torch.empty(3, device=h.device, dtype=torch.float32)
    dh = torch.empty_like(h)
    dw = _grad_like(weight)
    n = lib.opn_head_loss_workspace_bytes(B, T, H)
    ws = torch.empty(n, dtype=torch.uint8, device=h.device)
    rc = lib.opn_head_loss(B, T, H, h.data_ptr(), w.data_ptr(), labels.data_ptr(), _ptr(m), int(no_labels), y.data_ptr(),
                           out.data_ptr(), dh.data_ptr(), dw.data_ptr(), ws.data_ptr(), n, _stream())
    _lib.check(rc, "opn_head_loss")
    return y, out, dh, dw


class TrainingLossFn(torch.autograd.Function):
    """loss_and_grad as an autograd node: returns the 3-vector (total, prediction, consistency); only `total`
    carries gradient."""

    @staticmethod
    def forward(ctx, y, labels, mask, no_labels: bool):
        out, dy = loss_and_grad(y, labels, mask, no_labels)
        ctx.save_for_backward(dy)
        return out

    @staticmethod
    def backward(ctx, dout):
        (dy,) = ctx.saved_tensors
        # d total / dy was produced by the forward launch; scale by the incoming scalar
        return dy * dout[0], None, None, None


# ------------------------------------------------------------------------------------------
# functional front-ends
# ------------------------------------------------------------------------------------------
def linear(x, weight, bias=None, relu: bool = False):
    return LinearFn.apply(x, weight, bias, relu)


def _wants_stash(*tensors) -> bool:
    """True when a backward pass can follow: grad mode on and some input requires grad.  Evaluated by the front-ends
    because inside autograd.Function.forward grad mode is always off and needs_input_grad ignores torch.no_grad()."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


LSTM_HIDDEN_SIZES = (32, 64, 128, 256, 512)      # the sizes the persistent recurrence kernels are instantiated for


def lstm_layer(x, w_ih, w_hh):
    """One bias-free LSTM layer (gate order i, f, g, o; zero initial state) -> hs [B, T, H].
    Hidden sizes between the instantiated ones run in the next larger kernel with zero-padded weights: a padded unit has
    zero pre-activations, so its cell and output stay exactly 0 and it feeds nothing back -- the function and (through
    autograd's pad / slice) every gradient are those of the unpadded layer."""
    H = w_hh.shape[1]
    if H in LSTM_HIDDEN_SIZES or H > LSTM_HIDDEN_SIZES[-1]:       # larger sizes: the library reports them as unsupported
        return LstmLayerFn.apply(x, w_ih, w_hh, _wants_stash(x, w_ih, w_hh))
    Hp = next(h for h in LSTM_HIDDEN_SIZES if h >= H)
    pad = torch.nn.functional.pad
    w_ih_p = pad(w_ih.reshape(4, H, -1), (0, 0, 0, Hp - H)).reshape(4 * Hp, -1)
    w_hh_p = pad(w_hh.reshape(4, H, H), (0, Hp - H, 0, Hp - H)).reshape(4 * Hp, Hp)
    return LstmLayerFn.apply(x, w_ih_p, w_hh_p, _wants_stash(x, w_ih, w_hh))[..., :H]


def who_to_track(boxes, hs1, w_pred):
    return WhoToTrackFn.apply(boxes, hs1, w_pred)


def opnet_fused_available(h1: int, h2: int, pred_dim: int, batch: int = 0) -> bool:
    """The fused OPNet forward exists for the shipped config; OPN_OPNET_FUSED=0 selects the separate kernels.  At per-GPU
    batches where the batch-wide tcgen05 recurrence takes LSTM2 (opn_lstm_batchwide: groups of 128 videos instead of
    sequential waves of 32) the separate kernels are the faster path and are used."""
    if (h1, h2, pred_dim) != (256, 512, 15) or os.environ.get("OPN_OPNET_FUSED", "1") in ("0", ""):
        return False
    return not (batch > 0 and _lib.load().opn_lstm_batchwide(batch, h2))


def opnet_trunk(boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2):
    """(hs2 [B,T,H2], who-to-track logits [B,15,T]) of OPNet through the fused forward kernel."""
    return OPNetTrunkFn.apply(boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2,
                              _wants_stash(boxes, w_ih1, w_hh1, w_pred, w_ih2, w_hh2))


def add_layer_norm(x, res, weight, bias, eps: float = 1e-5):
    return AddLayerNormFn.apply(x, res, weight, bias, eps)


def self_attention(qkv, nhead: int, p_drop: float = 0.0, seed: int = 0, offset: int = 0):
    """p_drop > 0: attention-weight dropout with the mask of (seed, offset); one head consumes (S*S+3)//4 blocks."""
    if qkv.shape[-1] // 3 == 128 * nhead and os.environ.get("OPN_ATTENTION", "fused") != "materialized":
        return FusedSelfAttentionFn.apply(qkv, nhead, p_drop, seed, offset)
    return SelfAttentionFn.apply(qkv, nhead, p_drop, seed, offset)


def dropout(x, p: float, training: bool = True):
    """F.dropout with the library's counter-based mask; identity in eval mode or at p = 0."""
    if not training or p <= 0.0:
        return x
    seed, offset = dropout_stream.take(x.numel())
    return DropoutFn.apply(x, p, seed, offset)


def slot_linear_relu(boxes, weight, slot: int = 0):
    return SlotLinearReluFn.apply(boxes, weight, slot)


def training_loss(y, labels, mask=None, no_labels: bool = False):
    """(total, prediction, consistency) as a 3-vector; differentiate `[0]`."""
    return TrainingLossFn.apply(y, labels, mask, no_labels)
