"""CPU oracle for the OPNet temporal-reasoning hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``objectpermanence_b200``) never imports
anything under ``oracle/`` and has no CPU fallback.

What it restates
----------------
The five learned models of the reference, ``baselines/learned_models.py``:

* ``OPNet``            learned_models.py:18-52
* ``OPNetLstmMlp``     learned_models.py:55-89
* ``BaselineLstm``     learned_models.py:92-118
* ``NonLinearLstm``    learned_models.py:121-151
* ``TransformerLstm``  learned_models.py:154-197

The reference's arithmetic lives in a third-party dependency that is not vendored:
PyTorch (reference pin ``pytorch=1.4.0``, environment.yml:97).  The published
algorithms restated here, explicitly and without calling ``nn.LSTM`` /
``nn.TransformerEncoder``:

* LSTM (bias-free, unidirectional, zero initial state), PyTorch gate order
  i, f, g, o in the ``4H`` row blocks of ``weight_ih`` / ``weight_hh``::

      a   = W_ih x_t + W_hh h_{t-1}
      i,f,o = sigmoid(a_i), sigmoid(a_f), sigmoid(a_o);  g = tanh(a_g)
      c_t = f * c_{t-1} + i * g ;  h_t = o * tanh(c_t)

* Transformer encoder layer, PyTorch defaults (post-norm, ReLU, dim_feedforward=2048,
  layer_norm_eps=1e-5, *sequence-first* layout: the reference hands it a
  ``(B*T, 15, D)`` tensor, so the attended axis is ``B*T`` and the 15 object slots are
  the independent "batch" axis -- learned_models.py:166,183-185).  Dropout is the
  identity unless the caller pins the masks of the layer's four nn.Dropout sites
  (``encoder_layer(..., drop=...)``): parity with the reference is defined in ``eval()``
  mode, parity of the library's own train mode against pinned masks.

Parity pinning
--------------
The reference ships no tests, golden vectors or checkpoints for this path, so the
oracle is pinned against *outputs of the reference itself*: ``oracle/make_golden.py``
imports the unmodified reference modules from ``/root/reference`` (possible only in the
build container), runs them on seeded inputs, and commits the inputs, weights,
outputs and parameter gradients as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
checks this file against those fixtures on every CPU test run.

Gradients come from ``torch.autograd`` over the explicit restatement (any dtype; use
float64 for a truth value).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

Params = Dict[str, torch.Tensor]

MAX_OBJECTS = 15  # learned_models.py:13
BB_OUT_DIM = 4    # learned_models.py:15
FFN_DIM = 2048    # nn.TransformerEncoderLayer default dim_feedforward
LN_EPS = 1e-5     # nn.TransformerEncoderLayer default layer_norm_eps

# model name -> family, mirroring the dispatch of models_factory.py:42-74 (including the
# reference's "opent_no_labels" spelling at models_factory.py:64).
_FAMILY = {
    "baseline_lstm": "baseline_lstm", "baseline_lstm_no_labels": "baseline_lstm",
    "non_linear_lstm": "non_linear_lstm", "non_linear_lstm_no_labels": "non_linear_lstm",
    "transformer_lstm": "transformer_lstm", "transformer_lstm_no_labels": "transformer_lstm",
    "opnet": "opnet", "opent_no_labels": "opnet",
    "opnet_no_labels": "opnet",  # offered by the reference CLI (supported_models.py:12,30); its factory only
                                  # matches the misspelling above, the B200 factory accepts both
    "opnet_lstm_mlp": "opnet_lstm_mlp", "opnet_lstm_mlp_no_labels": "opnet_lstm_mlp",
}


def family_of(model_name: str) -> str:
    if model_name not in _FAMILY:
        raise AttributeError("Model name is incorrect")  # models_factory.py:74
    return _FAMILY[model_name]


def in_features_of(model_name: str) -> int:
    """6 tracks for the OPNet family, 5 for the rest (learned_models.py:14,21,58)."""
    return 6 if family_of(model_name) in ("opnet", "opnet_lstm_mlp") else 5


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def lstm_layer(x: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor) -> torch.Tensor:
    """One bias-free LSTM layer, batch-first, zero initial state.

    x [B,T,I], w_ih [4H,I], w_hh [4H,H] -> h [B,T,H].  Follows the nn.LSTM call sites
    learned_models.py:29,32,39,46 (bias=False, batch_first=True, num_layers=1).
    """
    B, T, _ = x.shape
    H = w_hh.shape[1]
    xp = x @ w_ih.t()  # [B,T,4H]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs: List[torch.Tensor] = []
    for t in range(T):
        a = xp[:, t, :] + h @ w_hh.t()
        i = torch.sigmoid(a[:, 0 * H:1 * H])
        f = torch.sigmoid(a[:, 1 * H:2 * H])
        g = torch.tanh(a[:, 2 * H:3 * H])
        o = torch.sigmoid(a[:, 3 * H:4 * H])
        c = f * c + i * g
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def lstm_layer_fast(x: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor) -> torch.Tensor:
    """Same function through PyTorch's own fused CPU LSTM (what the reference's
    ``nn.LSTM`` dispatches to on CPU).  Used only for the timed ``cpu_baseline`` so the
    baseline is not handicapped by a Python time loop."""
    B = x.shape[0]
    H = w_hh.shape[1]
    h0 = x.new_zeros(1, B, H)
    c0 = x.new_zeros(1, B, H)
    out, _, _ = torch._VF.lstm(x, (h0, c0), [w_ih, w_hh], False, 1, 0.0, False, False, True)
    return out


def lstm_stack(x: torch.Tensor, p: Params, prefix: str, num_layers: int, fast: bool = False) -> torch.Tensor:
    layer = lstm_layer_fast if fast else lstm_layer
    for k in range(num_layers):
        x = layer(x, p[f"{prefix}.weight_ih_l{k}"], p[f"{prefix}.weight_hh_l{k}"])
    return x


def who_to_track(boxes: torch.Tensor, h1: torch.Tensor, w_pred: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """learned_models.py:40-43: logits = h1 Wp^T, p = softmax over the 15 objects,
    frames_boxes[b,t,:] = sum_o p[b,t,o] * boxes[b,t,o,:]."""
    logits = h1 @ w_pred.t()
    probs = torch.softmax(logits, dim=-1)
    frames_boxes = (boxes * probs.unsqueeze(-1)).sum(dim=2)
    return frames_boxes, logits


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def encoder_layer(x: torch.Tensor, p: Params, prefix: str, nhead: int, drop=None, q_chunk: Optional[int] = None) -> torch.Tensor:
    """One post-norm encoder layer over x [S, N, D] (sequence-first, N independent
    columns).  Restates nn.TransformerEncoderLayer as built at
    learned_models.py:166 (d_model=D, nhead, defaults otherwise).

    ``drop`` = None is eval mode.  Train mode: ``drop(site, tensor) -> tensor`` is called at the layer's four
    nn.Dropout sites in execution order -- "<prefix>.attn" (attention weights [N,nhead,S,S], after the softmax),
    "<prefix>.dropout1" (the attention block's output [S,N,D]), "<prefix>.dropout" (after the ReLU, [S,N,2048]),
    "<prefix>.dropout2" ([S,N,D]) -- and applies whatever mask the caller pins (tests: the counter-based mask of
    oracle/dropout_mask.py; PyTorch's own random stream cannot be pinned).

    ``q_chunk``: evaluate the attention in blocks of that many query rows (the same arithmetic per row; the [S,S] score
    matrix of BASELINE config 3, S = 9600, is 1.5 GB per head in fp64 -- in blocks it fits any host)."""
    if drop is None:
        drop = lambda site, t: t
    S, N, D = x.shape
    dh = D // nhead
    qkv = x @ p[f"{prefix}.self_attn.in_proj_weight"].t() + p[f"{prefix}.self_attn.in_proj_bias"]
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]

    def heads(z):  # [S,N,D] -> [N,nhead,S,dh]
        return z.reshape(S, N, nhead, dh).permute(1, 2, 0, 3)

    q, k, v = heads(q), heads(k), heads(v)
    if q_chunk is None:
        scores = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
        attn = drop(f"{prefix}.attn", torch.softmax(scores, dim=-1)) @ v   # [N,nhead,S,dh]
    else:
        blocks = []
        for s0 in range(0, S, q_chunk):
            sc = (q[:, :, s0:s0 + q_chunk] @ k.transpose(-1, -2)) / math.sqrt(dh)
            blocks.append(drop(f"{prefix}.attn", torch.softmax(sc, dim=-1)) @ v)
        attn = torch.cat(blocks, dim=2)
    attn = attn.permute(2, 0, 1, 3).reshape(S, N, D)
    attn = attn @ p[f"{prefix}.self_attn.out_proj.weight"].t() + p[f"{prefix}.self_attn.out_proj.bias"]
    x = layer_norm(x + drop(f"{prefix}.dropout1", attn), p[f"{prefix}.norm1.weight"], p[f"{prefix}.norm1.bias"])
    ff = drop(f"{prefix}.dropout", torch.relu(x @ p[f"{prefix}.linear1.weight"].t() + p[f"{prefix}.linear1.bias"]))
    ff = drop(f"{prefix}.dropout2", ff @ p[f"{prefix}.linear2.weight"].t() + p[f"{prefix}.linear2.bias"])
    return layer_norm(x + ff, p[f"{prefix}.norm2.weight"], p[f"{prefix}.norm2.bias"])


# --------------------------------------------------------------------------------------
# the five models
# --------------------------------------------------------------------------------------
def opnet_forward(p: Params, boxes: torch.Tensor, fast: bool = False):
    """learned_models.py:35-52.  boxes [B,T,15,6] -> (y [B,T,4], logits [B,15,T])."""
    B, T = boxes.shape[:2]
    scene = boxes.reshape(B, T, -1)
    h1 = lstm_stack(scene, p, "object_to_track_LSTM", 1, fast)
    fb, logits = who_to_track(boxes, h1, p["object_to_track_prediction.weight"])
    h2 = lstm_stack(fb, p, "video_LSTM", 1, fast)
    y = h2 @ p["prediction_layer.weight"].t()
    return y, logits.permute(0, 2, 1).contiguous()


def opnet_lstm_mlp_forward(p: Params, boxes: torch.Tensor, fast: bool = False):
    """learned_models.py:72-89."""
    B, T = boxes.shape[:2]
    scene = boxes.reshape(B, T, -1)
    h1 = lstm_stack(scene, p, "object_to_track_LSTM", 1, fast)
    fb, logits = who_to_track(boxes, h1, p["object_to_track_prediction.weight"])
    hidden = torch.relu(fb @ p["hidden_layer.weight"].t())
    y = hidden @ p["prediction_layer.weight"].t()
    return y, logits.permute(0, 2, 1).contiguous()


def baseline_lstm_forward(p: Params, x: torch.Tensor, fast: bool = False):
    """learned_models.py:104-118."""
    B, T = x.shape[:2]
    h = lstm_stack(x.reshape(B, T, -1), p, "video_LSTM", 1, fast)
    return h @ p["predictions_layer.weight"].t()


def non_linear_lstm_forward(p: Params, x: torch.Tensor, fast: bool = False):
    """learned_models.py:135-151."""
    B, T = x.shape[:2]
    feats = torch.relu(x @ p["boxes_linear.weight"].t())
    h = lstm_stack(feats.reshape(B, T, -1), p, "video_LSTM", 2, fast)
    return h @ p["predictions_layer.weight"].t()


def transformer_lstm_forward(p: Params, x: torch.Tensor, config: Dict[str, int], fast: bool = False,
                             all_slots: bool = False, drop=None, q_chunk: Optional[int] = None):
    """learned_models.py:174-197; eval mode unless ``drop`` pins the dropout masks (see encoder_layer).

    ``all_slots=True`` evaluates the encoder on the full (B*T, 15, D) tensor exactly as
    the reference does; the default evaluates slot 0 only, which is the same function:
    the 15 slots never interact (they are the encoder's batch axis) and only slot 0 is
    read (learned_models.py:185)."""
    B, T = x.shape[:2]
    nhead = config["num_attention_heads"]
    feats = torch.relu(x @ p["boxes_linear.weight"].t())           # [B,T,15,D]
    seq = feats.reshape(B * T, MAX_OBJECTS, -1)
    if not all_slots:
        seq = seq[:, :1, :]
    for i in range(config["num_attention_layers"]):
        seq = encoder_layer(seq, p, f"attention_encoder.layers.{i}", nhead, drop, q_chunk)
    snitch = seq[:, 0, :].reshape(B, T, -1)
    h = lstm_stack(snitch, p, "video_LSTM", config["num_lstm_layers"], fast)
    return h @ p["predictions_layer.weight"].t()


def forward(model_name: str, p: Params, boxes: torch.Tensor, config: Optional[Dict[str, int]] = None,
            fast: bool = False, **kw):
    fam = family_of(model_name)
    if fam == "opnet":
        return opnet_forward(p, boxes, fast)
    if fam == "opnet_lstm_mlp":
        return opnet_lstm_mlp_forward(p, boxes, fast)
    if fam == "baseline_lstm":
        return baseline_lstm_forward(p, boxes, fast)
    if fam == "non_linear_lstm":
        return non_linear_lstm_forward(p, boxes, fast)
    return transformer_lstm_forward(p, boxes, config, fast, **kw)


# --------------------------------------------------------------------------------------
# parameter shapes and default init
# --------------------------------------------------------------------------------------
def param_shapes(model_name: str, config: Dict[str, int]) -> Dict[str, Tuple[int, ...]]:
    fam = family_of(model_name)
    s: Dict[str, Tuple[int, ...]] = {}
    if fam in ("opnet", "opnet_lstm_mlp"):
        H1, P, H2 = config["object_to_track_hidden_dim"], config["object_to_track_pred_dim"], config["videos_hidden_dim"]
        s["object_to_track_LSTM.weight_ih_l0"] = (4 * H1, 6 * MAX_OBJECTS)
        s["object_to_track_LSTM.weight_hh_l0"] = (4 * H1, H1)
        s["object_to_track_prediction.weight"] = (P, H1)
        if fam == "opnet":
            s["video_LSTM.weight_ih_l0"] = (4 * H2, 6)
            s["video_LSTM.weight_hh_l0"] = (4 * H2, H2)
        else:
            s["hidden_layer.weight"] = (H2, 6)
        s["prediction_layer.weight"] = (BB_OUT_DIM, H2)
    elif fam == "baseline_lstm":
        H = config["videos_hidden_dim"]
        s["video_LSTM.weight_ih_l0"] = (4 * H, 5 * MAX_OBJECTS)
        s["video_LSTM.weight_hh_l0"] = (4 * H, H)
        s["predictions_layer.weight"] = (BB_OUT_DIM, H)
    elif fam == "non_linear_lstm":
        D, H = config["boxes_features_dim"], config["videos_hidden_dim"]
        s["boxes_linear.weight"] = (D, 5)
        s["video_LSTM.weight_ih_l0"] = (4 * H, MAX_OBJECTS * D)
        s["video_LSTM.weight_hh_l0"] = (4 * H, H)
        s["video_LSTM.weight_ih_l1"] = (4 * H, H)
        s["video_LSTM.weight_hh_l1"] = (4 * H, H)
        s["predictions_layer.weight"] = (BB_OUT_DIM, H)
    else:
        D, H = config["boxes_features_dim"], config["lstm_hidden_dim"]
        s["boxes_linear.weight"] = (D, 5)
        for i in range(config["num_attention_layers"]):
            q = f"attention_encoder.layers.{i}"
            s[f"{q}.self_attn.in_proj_weight"] = (3 * D, D)
            s[f"{q}.self_attn.in_proj_bias"] = (3 * D,)
            s[f"{q}.self_attn.out_proj.weight"] = (D, D)
            s[f"{q}.self_attn.out_proj.bias"] = (D,)
            s[f"{q}.linear1.weight"] = (FFN_DIM, D)
            s[f"{q}.linear1.bias"] = (FFN_DIM,)
            s[f"{q}.linear2.weight"] = (D, FFN_DIM)
            s[f"{q}.linear2.bias"] = (D,)
            s[f"{q}.norm1.weight"] = (D,)
            s[f"{q}.norm1.bias"] = (D,)
            s[f"{q}.norm2.weight"] = (D,)
            s[f"{q}.norm2.bias"] = (D,)
        for k in range(config["num_lstm_layers"]):
            s[f"video_LSTM.weight_ih_l{k}"] = (4 * H, D if k == 0 else H)
            s[f"video_LSTM.weight_hh_l{k}"] = (4 * H, H)
        s["predictions_layer.weight"] = (BB_OUT_DIM, H)
    return s


def init_params(model_name: str, config: Dict[str, int], seed: int = 0, scale: float = 1.0,
                dtype=torch.float32) -> Params:
    """Random parameters with the reference's shapes and init *ranges* (U(-1/sqrt(H),
    1/sqrt(H)) for LSTM weights, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for Linear; LayerNorm
    weight 1 / bias 0 perturbed slightly so their gradients are exercised).  Not the
    reference's RNG stream -- parity tests load the same tensors into both sides."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for name, shape in param_shapes(model_name, config).items():
        if ".norm" in name:
            base = 1.0 if name.endswith("weight") else 0.0
            p[name] = (base + 0.1 * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)).to(dtype)
            continue
        if "LSTM" in name:
            bound = 1.0 / math.sqrt(shape[0] // 4)
        elif len(shape) == 2:
            bound = 1.0 / math.sqrt(shape[1])
        else:
            bound = 0.05
        w = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound * scale
        p[name] = w.to(dtype)
    return p


# --------------------------------------------------------------------------------------
# loss (training_main.py:192-210) and metric (tracking_utils.py:138-159)
# --------------------------------------------------------------------------------------
def training_loss(y: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None,
                  no_labels: bool = False) -> torch.Tensor:
    """L1 mean; for *_no_labels models the masked L1 plus 0.5 * consistency
    (mean L2 norm of consecutive-frame differences)."""
    pred = (y - labels).abs()
    if no_labels:
        pred = (pred * mask).mean()
        cons = torch.linalg.vector_norm(y[:, 1:, :] - y[:, :-1, :], ord=2, dim=-1).mean()
        return pred + 0.5 * cons
    return pred.mean()


def loss_and_grads(model_name: str, p: Params, boxes: torch.Tensor, labels: torch.Tensor,
                   config: Optional[Dict[str, int]] = None, dtype=torch.float64, fast: bool = False,
                   mask: Optional[torch.Tensor] = None, **kw):
    """Forward + L1 loss + autograd backward in ``dtype``.  Returns (y, logits|None, loss, grads).
    Keyword arguments go to the model's forward (transformer_lstm: all_slots, drop)."""
    q = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items()}
    out = forward(model_name, q, boxes.to(dtype), config, fast, **kw)
    y, logits = out if isinstance(out, tuple) else (out, None)
    loss = training_loss(y, labels.to(dtype), None if mask is None else mask.to(dtype),
                         no_labels=model_name.endswith("no_labels"))
    loss.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    return y.detach(), None if logits is None else logits.detach(), loss.detach(), grads


FRAME_SHAPE = np.array([320, 240, 320, 240])  # training_main.py:45


def video_iou(pred_px: np.ndarray, gt_px: np.ndarray) -> np.ndarray:
    """Per-frame IoU with the reference's +1-pixel convention (tracking_utils.py:138-159).
    pred_px, gt_px: int arrays [T,4] in xyxy pixels."""
    x11, y11, x12, y12 = [pred_px[:, i].astype(np.int64) for i in range(4)]
    x21, y21, x22, y22 = [gt_px[:, i].astype(np.int64) for i in range(4)]
    xa, ya = np.maximum(x11, x21), np.maximum(y11, y21)
    xb, yb = np.minimum(x12, x22), np.minimum(y12, y22)
    inter = np.maximum(xb - xa + 1, 0) * np.maximum(yb - ya + 1, 0)
    a1 = (x12 - x11 + 1) * (y12 - y11 + 1)
    a2 = (x22 - x21 + 1) * (y22 - y21 + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (a1 + a2 - inter)


def mean_iou(y: np.ndarray, labels: np.ndarray) -> float:
    """training_main.py:97-112: scale to pixels, truncate to int32, per-video mean IoU,
    mean over videos.  y, labels: float [N,T,4] normalised."""
    pred = (y.reshape(-1, 4) * FRAME_SHAPE).reshape(y.shape).astype(np.int32)
    gt = (labels.reshape(-1, 4) * FRAME_SHAPE).reshape(labels.shape).astype(np.int32)
    per_video = [float(np.mean(video_iou(pred[n], gt[n]))) for n in range(pred.shape[0])]
    return float(np.mean(per_video))
