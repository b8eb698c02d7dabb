"""The five learned temporal-reasoning models behind the reference's nn.Module surface.

Class names, constructor signature ``Model(config: Dict[str, int])``, parameter names and
shapes, and return conventions are those of the reference's ``baselines/learned_models.py``
(state dicts are interchangeable); the arithmetic is not: every forward and
backward runs hand-written sm_100a kernels through ``objectpermanence_b200.ops``.

Parameters are held as plain ``nn.Parameter``s under holder modules named like the reference's
``nn.LSTM`` / ``nn.Linear`` attributes so that ``state_dict()`` keys match exactly.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn

from . import ops

FFN_DIM = 2048   # nn.TransformerEncoderLayer default dim_feedforward (learned_models.py:166)
LN_EPS = 1e-5
DROPOUT_P = 0.1  # nn.TransformerEncoderLayer default dropout (learned_models.py:166)


class LstmWeights(nn.Module):
    """Parameter holder with nn.LSTM's key names (weight_ih_l{k}, weight_hh_l{k}; bias=False) and
    default init U(-1/sqrt(H), 1/sqrt(H)).  Calling it runs the stacked persistent LSTM."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        bound = 1.0 / math.sqrt(hidden_size)
        for k in range(num_layers):
            in_k = input_size if k == 0 else hidden_size
            self.register_parameter(f"weight_ih_l{k}", nn.Parameter(torch.empty(4 * hidden_size, in_k).uniform_(-bound, bound)))
            self.register_parameter(f"weight_hh_l{k}", nn.Parameter(torch.empty(4 * hidden_size, hidden_size).uniform_(-bound, bound)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for k in range(self.num_layers):
            x = ops.lstm_layer(x, getattr(self, f"weight_ih_l{k}"), getattr(self, f"weight_hh_l{k}"))
        return x


class LinearWeights(nn.Module):
    """Parameter holder with nn.Linear's key names and default (Kaiming-uniform, a=sqrt(5)) init."""

    def __init__(self, in_features: int, out_features: int, bias: bool = False):
        super().__init__()
        bound = 1.0 / math.sqrt(in_features)
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-bound, bound))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def forward(self, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
        return ops.linear(x, self.weight, self.bias, relu)


class AbstractCaterModel(nn.Module):
    """learned_models.py:8-15."""

    def __init__(self, config: Dict[str, int]):
        super().__init__()
        self.config: Dict[str, int] = config
        self.max_objects_in_frame = 15
        self.bb_in_dim = 5
        self.bb_out_dim = 4

    def _check(self, boxes: torch.Tensor) -> None:
        if boxes.dim() != 4 or boxes.shape[2] != self.max_objects_in_frame or boxes.shape[3] != self.bb_in_dim:
            raise RuntimeError(f"{type(self).__name__} expects boxes [B, T, {self.max_objects_in_frame}, "
                               f"{self.bb_in_dim}], got {tuple(boxes.shape)}")


class _WhoToTrackMixin:
    def _track(self, boxes: torch.Tensor):
        B, T = boxes.shape[:2]
        h1 = self.object_to_track_LSTM(boxes.reshape(B, T, -1))
        return ops.who_to_track(boxes, h1, self.object_to_track_prediction.weight)


class OPNet(AbstractCaterModel, _WhoToTrackMixin):
    """learned_models.py:18-52: LSTM1 -> who-to-track -> LSTM2 -> bbox head.
    Returns (y [B,T,4], who-to-track logits [B,15,T])."""

    def __init__(self, config: Dict[str, int]):
        super().__init__(config)
        self.bb_in_dim = 6
        h1, h2 = config["object_to_track_hidden_dim"], config["videos_hidden_dim"]
        self.object_to_track_LSTM = LstmWeights(self.bb_in_dim * 15, h1)
        self.object_to_track_prediction = LinearWeights(h1, config["object_to_track_pred_dim"])
        self.video_LSTM = LstmWeights(self.bb_in_dim, h2)
        self.prediction_layer = LinearWeights(h2, self.bb_out_dim)

    head_name = "prediction_layer"

    def trunk(self, boxes: torch.Tensor):
        """Everything in front of the bbox head -> (hidden states [B,T,H2], who-to-track logits [B,15,T])."""
        self._check(boxes)
        l1, l2, wp = self.object_to_track_LSTM, self.video_LSTM, self.object_to_track_prediction.weight
        if ops.opnet_fused_available(l1.hidden_size, l2.hidden_size, wp.shape[0], boxes.shape[0]):
            # one persistent kernel for LSTM1 + who-to-track + LSTM2 (shipped config)
            return ops.opnet_trunk(boxes, l1.weight_ih_l0, l1.weight_hh_l0, wp, l2.weight_ih_l0, l2.weight_hh_l0)
        frames_boxes, logits = self._track(boxes)
        return self.video_LSTM(frames_boxes), logits

    def forward(self, boxes: torch.Tensor):
        h2, logits = self.trunk(boxes)
        return self.prediction_layer(h2), logits


class OPNetLstmMlp(AbstractCaterModel, _WhoToTrackMixin):
    """learned_models.py:55-89: the second LSTM replaced by a one-layer MLP."""

    def __init__(self, config: Dict[str, int]):
        super().__init__(config)
        self.bb_in_dim = 6
        h1, h2 = config["object_to_track_hidden_dim"], config["videos_hidden_dim"]
        self.object_to_track_LSTM = LstmWeights(self.bb_in_dim * 15, h1)
        self.object_to_track_prediction = LinearWeights(h1, config["object_to_track_pred_dim"])
        self.hidden_layer = LinearWeights(self.bb_in_dim, h2)
        self.prediction_layer = LinearWeights(h2, self.bb_out_dim)

    head_name = "prediction_layer"

    def trunk(self, boxes: torch.Tensor):
        self._check(boxes)
        frames_boxes, logits = self._track(boxes)
        return self.hidden_layer(frames_boxes, relu=True), logits

    def forward(self, boxes: torch.Tensor):
        h, logits = self.trunk(boxes)
        return self.prediction_layer(h), logits


class BaselineLstm(AbstractCaterModel):
    """learned_models.py:92-118 (note the attribute is `predictions_layer`, with an s)."""

    def __init__(self, config: Dict[str, int]):
        super().__init__(config)
        h = config["videos_hidden_dim"]
        self.video_LSTM = LstmWeights(self.max_objects_in_frame * self.bb_in_dim, h)
        self.predictions_layer = LinearWeights(h, self.bb_out_dim)

    head_name = "predictions_layer"

    def trunk(self, x: torch.Tensor):
        self._check(x)
        B, T = x.shape[:2]
        return self.video_LSTM(x.reshape(B, T, -1)), None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.predictions_layer(self.trunk(x)[0])


class NonLinearLstm(AbstractCaterModel):
    """learned_models.py:121-151: per-object relu(Linear 5->D), 2-layer LSTM, bbox head."""

    def __init__(self, config: Dict[str, int]):
        super().__init__(config)
        d, h = config["boxes_features_dim"], config["videos_hidden_dim"]
        self.boxes_linear = LinearWeights(self.bb_in_dim, d)
        self.video_LSTM = LstmWeights(self.max_objects_in_frame * d, h, num_layers=2)
        self.predictions_layer = LinearWeights(h, self.bb_out_dim)

    head_name = "predictions_layer"

    def trunk(self, x: torch.Tensor):
        self._check(x)
        B, T = x.shape[:2]
        feats = self.boxes_linear(x, relu=True)
        return self.video_LSTM(feats.reshape(B, T, -1)), None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.predictions_layer(self.trunk(x)[0])


class _SelfAttnWeights(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        nn.init.xavier_uniform_(self.in_proj_weight)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = LinearWeights(d, d, bias=True)
        with torch.no_grad():
            self.out_proj.bias.zero_()


class _NormWeights(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class EncoderLayer(nn.Module):
    """Post-norm encoder layer with nn.TransformerEncoderLayer's parameter names, evaluated on one
    sequence of S rows.  In train mode the layer's four dropout sites (attention weights, dropout1 after the
    attention block, dropout inside the feed-forward block, dropout2 after it; p = 0.1, the PyTorch default the
    reference relies on) are applied with the library's counter-based mask (ops.dropout): the same distribution
    as the reference, not the same random stream, so parity with the reference is defined in eval() mode or at
    `dropout_p = 0`."""

    def __init__(self, d: int, nhead: int, dropout_p: float = DROPOUT_P):
        super().__init__()
        self.nhead = nhead
        self.dropout_p = dropout_p
        self.self_attn = _SelfAttnWeights(d)
        self.linear1 = LinearWeights(d, FFN_DIM, bias=True)
        self.linear2 = LinearWeights(FFN_DIM, d, bias=True)
        self.norm1 = _NormWeights(d)
        self.norm2 = _NormWeights(d)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # x [S, D]
        p = self.dropout_p if self.training else 0.0
        qkv = ops.linear(x, self.self_attn.in_proj_weight, self.self_attn.in_proj_bias)
        if p > 0.0:
            S = x.shape[0]
            seed, offset = ops.dropout_stream.take(self.nhead * (4 * ((S * S + 3) // 4)))
            attended = ops.self_attention(qkv, self.nhead, p, seed, offset)
        else:
            attended = ops.self_attention(qkv, self.nhead)
        attn = ops.dropout(self.self_attn.out_proj(attended), p, self.training)
        x = ops.add_layer_norm(x, attn, self.norm1.weight, self.norm1.bias, LN_EPS)
        ff = self.linear2(ops.dropout(self.linear1(x, relu=True), p, self.training))
        ff = ops.dropout(ff, p, self.training)
        return ops.add_layer_norm(x, ff, self.norm2.weight, self.norm2.bias, LN_EPS)


class _Encoder(nn.Module):
    def __init__(self, d: int, nhead: int, num_layers: int):
        super().__init__()
        self.layers = nn.ModuleList([EncoderLayer(d, nhead) for _ in range(num_layers)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for layer in self.layers:
            x = layer(x)
        return x


class TransformerLstm(AbstractCaterModel):
    """learned_models.py:154-197.  The reference feeds a (B*T, 15, D) tensor to a sequence-first
    encoder and keeps slot 0, so the function actually computed is self-attention over the B*T
    snitch-slot rows of the mini-batch; the other 14 slots are dead compute and are skipped."""

    def __init__(self, config: Dict[str, int]):
        super().__init__(config)
        d = config["boxes_features_dim"]
        self.boxes_linear = LinearWeights(self.bb_in_dim, d)
        self.attention_encoder = _Encoder(d, config["num_attention_heads"], config["num_attention_layers"])
        self.video_LSTM = LstmWeights(d, config["lstm_hidden_dim"], num_layers=config["num_lstm_layers"])
        self.predictions_layer = LinearWeights(config["lstm_hidden_dim"], self.bb_out_dim)

    head_name = "predictions_layer"

    def trunk(self, x: torch.Tensor):
        self._check(x)
        B, T = x.shape[:2]
        snitch = ops.slot_linear_relu(x, self.boxes_linear.weight, 0)      # [B,T,D]
        attended = self.attention_encoder(snitch.reshape(B * T, -1))       # [B*T, D]
        return self.video_LSTM(attended.reshape(B, T, -1)), None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.predictions_layer(self.trunk(x)[0])
