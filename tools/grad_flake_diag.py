"""Diagnostic: repeat the fused OPNet backward + weight-gradient contractions on one forward graph and report every
iteration whose intermediates (dgates2, dgates1, d logits: deterministic kernels, compared bit-wise with iteration 0)
or weight gradients (compared with an fp64 contraction of the captured intermediates) are off."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
ops.set_debug_sync(True)
H1, H2 = 256, 512
ITERS = int(os.environ.get("ITERS", "40"))


def rnd(shape, seed, scale):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * scale).float()


cap = {}
orig_wg, orig_pg = ops._lstm_weight_grads, ops._wtt_weight_grad


def wg(dgates, x, hs, w_ih, w_hh, a, b):
    out = orig_wg(dgates, x, hs, w_ih, w_hh, a, b)
    cap["L2" if w_hh.shape[1] == H2 else "L1"] = (dgates, x, hs, out[0], out[1])
    return out


def pg(hs1, dl):
    out = orig_pg(hs1, dl)
    cap["P"] = (hs1, dl, out)
    return out


ops._lstm_weight_grads, ops._wtt_weight_grad = wg, pg

for B, T in ((11, 37), (4, 16), (32, 64)):
    boxes = (torch.rand(B, T, 15, 6, generator=torch.Generator().manual_seed(5 + B)) * (torch.rand(B, T, 15, 1) > 0.3)).to(dev)
    w = {"ih1": rnd((4 * H1, 90), 1, 1 / math.sqrt(H1)), "hh1": rnd((4 * H1, H1), 2, 1 / math.sqrt(H1)),
         "pred": rnd((15, H1), 3, 1 / math.sqrt(H1)), "ih2": rnd((4 * H2, 6), 4, 1 / math.sqrt(H2)),
         "hh2": rnd((4 * H2, H2), 5, 1 / math.sqrt(H2))}
    ws = {k: v.to(dev).requires_grad_(True) for k, v in w.items()}
    dh2 = rnd((B, T, H2), 6, 0.01).to(dev)
    h2, logits = ops.opnet_trunk(boxes, ws["ih1"], ws["hh1"], ws["pred"], ws["ih2"], ws["hh2"])
    bad = 0
    for fused_bwd, mode in (("1", "0"), ("1", "1"), ("0", "1")):
        if (fused_bwd, mode) != ("1", "1"):
            first = None          # separate kernels are not bit-equal to the fused backward
        os.environ["OPN_OPNET_FUSED_BWD"] = fused_bwd
        os.environ["OPN_OPNET_WGRAD_OVERLAP"] = mode
        for it in range(ITERS):
            for v in ws.values():
                v.grad = None
            cap.clear()
            h2.backward(dh2, retain_graph=True)
            torch.cuda.synchronize()
            dg2, fb, hs2, dwih2, dwhh2 = cap["L2"]
            dg1, x1, hs1, dwih1, dwhh1 = cap["L1"]
            _, dl, dwp = cap["P"]
            inter = {"dgates2": dg2, "dgates1": dg1, "dlogits": dl}
            if first is None:
                first = {k: v.clone() for k, v in inter.items()}
            msgs = []
            for k, v in inter.items():
                if not torch.equal(v, first[k]):
                    d = (v - first[k]).abs()
                    idx = torch.nonzero(d.reshape(B, T, -1).amax(-1))
                    msgs.append(f"{k} differs from iteration 0: max {d.max().item():.3e} at (b,t) {idx[:6].tolist()} ({idx.shape[0]} frames)")
            refs = {"ih2": (dg2.double().reshape(B * T, -1).t() @ fb.double().reshape(B * T, -1), dwih2),
                    "hh2": (torch.einsum("btg,bth->gh", dg2[:, 1:].double(), hs2[:, :-1].double()), dwhh2),
                    "ih1": (dg1.double().reshape(B * T, -1).t() @ x1.double().reshape(B * T, -1), dwih1),
                    "hh1": (torch.einsum("btg,bth->gh", dg1[:, 1:].double(), hs1[:, :-1].double()), dwhh1),
                    "pred": (dl.double().reshape(B * T, -1).t() @ hs1.double().reshape(B * T, -1), dwp)}
            for k, (want, got) in refs.items():
                err = (got.double() - want).abs().max().item()
                tol = 2e-4 * want.abs().max().item() + 1e-8
                if err > tol or not torch.equal(got, ws[k].grad):
                    r, c = divmod(int((got.double() - want).abs().argmax()), want.shape[1])
                    msgs.append(f"dW_{k}: err {err:.3e} (max |want| {want.abs().max().item():.3e}) at [{r},{c}] got {got[r, c].item():.6e} want {want[r, c].item():.6e}; "
                                f"returned==param.grad {torch.equal(got, ws[k].grad)}")
            if msgs:
                bad += 1
                print(f"[B={B} T={T} fused_bwd={fused_bwd} overlap={mode} it={it}] " + " | ".join(msgs), flush=True)
    print(f"B={B} T={T}: {bad} bad iterations of {3 * ITERS}", flush=True)
