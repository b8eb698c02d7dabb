"""GPU: regression cases added at the very end of round 1.  They run LAST (file name) so that the long validated order
of the other GPU tests is unchanged (seen green on a B200 in the round's last call, gpurun_out c20)."""
import math

import pytest
import torch

from objectpermanence_b200 import ops
from oracle import opnet_oracle as oracle

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * scale).float()


@pytest.mark.parametrize("T", [38, 39])
def test_opnet_fused_backward_back_to_back_launches(cuda_device, monkeypatch, T):
    """The fused backward takes its first LSTM2 sweep (iteration 1) from shared-memory landing slots that no sweep was
    issued into yet; step 0 carries parity 1, and for T = 2, 3 (mod 4) the LAST sweep of a launch also carries parity 1,
    so what a launch leaves in shared memory would pass as ready words in the next launch on the same SM.  The kernel
    zeroes the slots at its start; this runs the backward three times on one graph (nothing with a large shared-memory
    footprint in between at this size) and compares every launch with the fp64 oracle."""
    monkeypatch.setenv("OPN_OPNET_FUSED_BWD", "1")
    monkeypatch.setenv("OPN_OPNET_WGRAD_OVERLAP", "0")
    B, H1, H2 = 5, 256, 512
    boxes = torch.rand(B, T, 15, 6, generator=torch.Generator().manual_seed(40 + T))
    w = {"ih1": _rand((4 * H1, 90), 1, 1 / math.sqrt(H1)), "hh1": _rand((4 * H1, H1), 2, 1 / math.sqrt(H1)),
         "pred": _rand((15, H1), 3, 1 / math.sqrt(H1)), "ih2": _rand((4 * H2, 6), 4, 1 / math.sqrt(H2)),
         "hh2": _rand((4 * H2, H2), 5, 1 / math.sqrt(H2))}
    dh2 = _rand((B, T, H2), 6, 0.01)
    wr = {k: v.double().requires_grad_(True) for k, v in w.items()}
    h1_r = oracle.lstm_layer(boxes.double().reshape(B, T, -1), wr["ih1"], wr["hh1"])
    fb_r, _ = oracle.who_to_track(boxes.double(), h1_r, wr["pred"])
    oracle.lstm_layer(fb_r, wr["ih2"], wr["hh2"]).backward(dh2.double())
    ws = {k: v.to(cuda_device).requires_grad_(True) for k, v in w.items()}
    h2, _ = ops.opnet_trunk(boxes.to(cuda_device), ws["ih1"], ws["hh1"], ws["pred"], ws["ih2"], ws["hh2"])
    dh2_d = dh2.to(cuda_device)
    for launch in range(3):
        for v in ws.values():
            v.grad = None
        h2.backward(dh2_d, retain_graph=True)
        for k in w:
            want = wr[k].grad
            err = (ws[k].grad.cpu().double() - want).abs().max().item()
            assert err <= 2e-4 * max(1e-3, want.abs().max().item()), (launch, k, err)


def test_transformer_lstm_train_mode_matches_the_oracle_with_pinned_masks(cuda_device, monkeypatch):
    """Train mode end to end: the (seed, offset) of every dropout site of the run is recorded, the oracle applies the
    restated Philox masks (oracle/dropout_mask.py) at the same sites, and outputs and every gradient must agree."""
    from objectpermanence_b200.models_factory import ModelsFactory
    from objectpermanence_b200.synthetic import make_batch
    from oracle import dropout_mask

    class Recorder(ops.DropoutStream):
        def __init__(self):
            self.log = []

        def take(self, n):
            key = super().take(n)
            self.log.append(key)
            return key

    rec = Recorder()
    monkeypatch.setattr(ops, "dropout_stream", rec)
    cfg = {"boxes_features_dim": 32, "num_attention_heads": 2, "num_attention_layers": 2, "num_lstm_layers": 2,
           "lstm_hidden_dim": 32}
    B, T, p_drop = 2, 12, 0.1
    boxes_np, labels_np, _ = make_batch(B, T, 5, seed=91)
    boxes, labels = torch.from_numpy(boxes_np), torch.from_numpy(labels_np)
    params = oracle.init_params("transformer_lstm", cfg, seed=4)
    model = ModelsFactory.get_model("transformer_lstm", cfg)
    model.load_state_dict(params)
    model = model.to(cuda_device).train()
    torch.manual_seed(17)
    y = model(boxes.to(cuda_device))
    ops.training_loss(y, labels.to(cuda_device))[0].backward()
    assert len(rec.log) == 4 * cfg["num_attention_layers"]
    keys = iter(rec.log)

    def drop(site, t):
        seed, offset = next(keys)
        if site.endswith(".attn"):          # [1, nhead, S, S]: head h consumes its own (S*S+3)//4 Philox blocks
            S = t.shape[-1]
            blocks = (S * S + 3) // 4
            keep = torch.stack([torch.from_numpy(dropout_mask.keep_mask(S * S, p_drop, seed, offset + h * blocks)).reshape(S, S)
                                for h in range(t.shape[1])])[None]
        else:
            keep = torch.from_numpy(dropout_mask.keep_mask(t.numel(), p_drop, seed, offset)).reshape(t.shape)
        return t * keep.to(t.dtype) / (1.0 - p_drop)

    y_ref, _, _, g_ref = oracle.loss_and_grads("transformer_lstm", params, boxes, labels, cfg, dtype=torch.float64, drop=drop)
    assert (y.detach().cpu().double() - y_ref).abs().max().item() <= 1e-4
    for k, v in model.named_parameters():
        err = (v.grad.cpu().double() - g_ref[k]).abs().max().item()
        assert err <= 2e-3 * max(1e-3, g_ref[k].abs().max().item()), (k, err)
