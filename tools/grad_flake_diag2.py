"""Diagnostic for one shape: fresh fused forward + backward in several modes, every intermediate and weight gradient
compared with an fp64 autograd restatement evaluated on the GPU (explicit per-step LSTM with retained pre-activations)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
ops.set_debug_sync(True)
H1, H2 = 256, 512
B, T = int(os.environ.get("BB", "11")), int(os.environ.get("TT", "37"))


def rnd(shape, seed, scale):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * scale).float()


cap = {}
orig_wg, orig_pg = ops._lstm_weight_grads, ops._wtt_weight_grad


def wg(dgates, x, hs, w_ih, w_hh, a, b):
    out = orig_wg(dgates, x, hs, w_ih, w_hh, a, b)
    cap["L2" if w_hh.shape[1] == H2 else "L1"] = (dgates, out[0], out[1], x)
    return out


def pg(hs1, dl):
    out = orig_pg(hs1, dl)
    cap["P"] = (dl, out)
    return out


ops._lstm_weight_grads, ops._wtt_weight_grad = wg, pg


def lstm_ref(x, w_ih, w_hh):
    Bn, Tn, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(Bn, H); c = x.new_zeros(Bn, H)
    hs, pre = [], []
    for t in range(Tn):
        a = x[:, t] @ w_ih.t() + h @ w_hh.t()
        a.retain_grad(); pre.append(a)
        i, f, g, o = a.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        hs.append(h)
    return torch.stack(hs, 1), pre


w = {"ih1": rnd((4 * H1, 90), 1, 1 / math.sqrt(H1)), "hh1": rnd((4 * H1, H1), 2, 1 / math.sqrt(H1)),
     "pred": rnd((15, H1), 3, 1 / math.sqrt(H1)), "ih2": rnd((4 * H2, 6), 4, 1 / math.sqrt(H2)),
     "hh2": rnd((4 * H2, H2), 5, 1 / math.sqrt(H2))}
dh2 = rnd((B, T, H2), 6, 0.01).to(dev)
for rep, (fused_bwd, overlap) in enumerate((("1", "1"), ("1", "0"), ("0", "1"), ("1", "0"), ("1", "1"), ("1", "0"))):
    boxes = (torch.rand(B, T, 15, 6) * (torch.rand(B, T, 15, 1) > 0.3)).to(dev)     # fresh data each time, as the test
    os.environ["OPN_OPNET_FUSED_BWD"], os.environ["OPN_OPNET_WGRAD_OVERLAP"] = fused_bwd, overlap
    ws = {k: v.to(dev).requires_grad_(True) for k, v in w.items()}
    cap.clear()
    h2, logits = ops.opnet_trunk(boxes, ws["ih1"], ws["hh1"], ws["pred"], ws["ih2"], ws["hh2"])
    h2.backward(dh2)
    torch.cuda.synchronize()
    wr = {k: v.double().to(dev).requires_grad_(True) for k, v in w.items()}
    bx = boxes.double()
    h1_r, pre1 = lstm_ref(bx.reshape(B, T, -1), wr["ih1"], wr["hh1"])
    lg_r = h1_r @ wr["pred"].t()
    lg_r.retain_grad()
    fb_r = torch.einsum("bfot,bfo->bft", bx, torch.softmax(lg_r, -1))
    h2_r, pre2 = lstm_ref(fb_r, wr["ih2"], wr["hh2"])
    h2_r.backward(dh2.double())
    dg2_r = torch.stack([a.grad for a in pre2], 1); dg1_r = torch.stack([a.grad for a in pre1], 1)

    def err(got, want):
        d = (got.double() - want).abs()
        i = int(d.argmax())
        return f"{d.max().item():.2e}/{want.abs().max().item():.1e}@{i}"

    dg2, dwih2, dwhh2, fb = cap["L2"]; dg1, dwih1, dwhh1, _ = cap["L1"]; dl, dwp = cap["P"]
    ih2_from_cap = dg2.double().reshape(B * T, -1).t() @ fb.double().reshape(B * T, -1)
    print(f"[{rep}] fused_bwd={fused_bwd} overlap={overlap}: h2 {err(h2, h2_r)} dgates2 {err(dg2, dg2_r)} dlogits {err(dl, lg_r.grad)} "
          f"dgates1 {err(dg1, dg1_r)} | dW ih2 {err(ws['ih2'].grad, wr['ih2'].grad)} hh2 {err(ws['hh2'].grad, wr['hh2'].grad)} "
          f"pred {err(ws['pred'].grad, wr['pred'].grad)} ih1 {err(ws['ih1'].grad, wr['ih1'].grad)} hh1 {err(ws['hh1'].grad, wr['hh1'].grad)}"
          f" | fb {err(fb, fb_r)} ih2 vs contraction of the captured dgates2, fb: {err(dwih2, ih2_from_cap)}",
          flush=True)
    bad = (dg2.double() - dg2_r).abs().reshape(B, T, -1).amax(-1)
    if bad.max().item() > 1e-6:
        idx = torch.nonzero(bad > 1e-6)
        print(f"    dgates2 off at {idx.shape[0]} (b,t) frames, first {idx[:8].tolist()}, last {idx[-4:].tolist()}", flush=True)
