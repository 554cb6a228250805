"""Drop-in ``nn.Module``s for the reference GNN layers, running on sm_100a kernels.

Same constructor signatures, attribute names, ``state_dict`` keys
(``layers.{i}.linear.{weight,bias}``, ``layers.{i}.lynorm.{weight,bias}``) and
initialisation order as /root/reference/src/components/graphs/models.py:15-170,
so checkpoints written by the reference (``model_train.py:411-419``,
``utils/training.py:47-49``) load unchanged and the train / predict drivers
(``model_train.py:156-163,320``, ``model_predict.py:113-122,145``) can swap the
import and nothing else.  ``forward`` accepts a ``PageGraphBatch`` or any
DGL-like graph on a CUDA device.  There is no CPU path.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layers as L
from ._lib import GteError
from .graph import as_page_graph_batch


def _is_relu(act) -> bool:
    return act is F.relu or act is torch.relu or isinstance(act, nn.ReLU)


class GcnSAGELayer(nn.Module):
    """models.py:15-78."""

    def __init__(self, in_feats, out_feats, activation, dropout, bias=True, use_pp=False, use_lynorm=True):
        super().__init__()
        # The input feature size gets doubled: [h | ah * norm] (models.py:25-27)
        self.linear = nn.Linear(2 * in_feats, out_feats, bias=bias)
        self.activation = activation
        self.use_pp = use_pp
        if dropout:
            self.dropout = nn.Dropout(p=dropout)
        else:
            self.dropout = 0.0
        if use_lynorm:
            self.lynorm = nn.LayerNorm(out_feats, elementwise_affine=True)
        else:
            self.lynorm = lambda x: x
        self.reset_parameters()

    def reset_parameters(self):  # models.py:40-44
        stdv = 1.0 / math.sqrt(self.linear.weight.size(1))
        self.linear.weight.data.uniform_(-stdv, stdv)
        if self.linear.bias is not None:
            self.linear.bias.data.uniform_(-stdv, stdv)

    # reference helper methods, kept for API parity --------------------------
    def concat(self, h, ah, norm):  # models.py:69-72
        ah = ah * norm
        return torch.cat((h, ah), dim=1)

    def get_norm(self, g):  # models.py:74-78
        pg = as_page_graph_batch(g)
        return pg.norm().unsqueeze(1).to(self.linear.weight.device)

    def _dropout_active(self) -> bool:
        return bool(self.dropout) and self.training and self.dropout.p > 0

    def forward(self, g, h):
        pg = as_page_graph_batch(g)  # local copy of ndata/edata: the caller's graph is never mutated (models.py:47)
        has_ln = isinstance(self.lynorm, nn.LayerNorm)
        act = self.activation
        fused_relu = _is_relu(act)
        gamma = self.lynorm.weight if has_ln else None
        beta = self.lynorm.bias if has_ln else None
        eps = self.lynorm.eps if has_ln else 1e-5
        w_edge = None
        if not self.use_pp:
            if "feat" not in pg.edata:
                raise KeyError("feat")  # same failure as the reference when edge weights are absent (models.py:53)
            w_edge = pg.edata["feat"]

        drop = None
        if self._dropout_active():
            # nn.Dropout on [h | ah * norm] (models.py:60-61) inside the layer's kernels: Philox keyed from torch's CUDA
            # generator, the concat is never materialised, the mask is recomputed in backward
            drop = L.torch_generator_dropout_spec(self.dropout.p, h.device, h.shape[0], h.shape[1] if self.use_pp else 2 * h.shape[1])
        out = L.SageLayerFunction.apply(h, self.linear.weight, self.linear.bias, gamma, beta, w_edge, pg,
                                        has_ln, fused_relu, eps, L.GCN, self.use_pp, drop)
        if act and not fused_relu:
            out = act(out)
        return out


class GcnSAGE(nn.Module):
    """models.py:80-116."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout, use_pp=False):
        super().__init__()
        self.layers = nn.ModuleList()
        self.dropout = nn.Dropout(dropout)
        # input layer
        self.layers.append(GcnSAGELayer(in_feats, n_hidden, activation=activation, dropout=dropout, use_pp=use_pp,
                                        use_lynorm=True))
        # hidden layers
        for _ in range(n_layers - 2):
            self.layers.append(GcnSAGELayer(n_hidden, n_hidden, activation=activation, dropout=dropout, use_pp=False,
                                            use_lynorm=True))
        # output layer
        self.layers.append(GcnSAGELayer(n_hidden, n_classes, activation=None, dropout=False, use_pp=False,
                                        use_lynorm=False))

    def forward(self, g):
        pg = as_page_graph_batch(g)
        h = pg.ndata["feat"]
        if self.training and self.dropout.p > 0:  # models.py:113, native kernel (mask recomputed in backward)
            h = L.DropoutFunction.apply(h, L.torch_generator_dropout_spec(self.dropout.p, h.device, h.shape[0], h.shape[1]))
        for layer in self.layers:
            h = layer(pg, h)
        return h


class WeightedMeanSAGELayer(nn.Module):
    """models.py:118-152: mean_{u->v}(h[u] * w_e) then Linear(cat[h, h_N])."""

    def __init__(self, in_feat, out_feat):
        super().__init__()
        self.linear = nn.Linear(in_feat * 2, out_feat)

    def forward(self, g, h, w):
        pg = as_page_graph_batch(g)
        return L.SageLayerFunction.apply(h, self.linear.weight, self.linear.bias, None, None, w, pg, False, False,
                                         1e-5, L.MEAN, False)


class MeanSAGE(nn.Module):
    """models.py:154-170."""

    def __init__(self, in_feats, h_feats, num_classes, n_layers):
        super().__init__()
        self.n_layers = n_layers
        self.layers = nn.ModuleList()
        self.layers.append(WeightedMeanSAGELayer(in_feats, h_feats))
        for _ in range(n_layers - 1):
            self.layers.append(WeightedMeanSAGELayer(h_feats, h_feats))
        self.layers.append(WeightedMeanSAGELayer(h_feats, num_classes))

    def forward(self, g, h, w):
        pg = as_page_graph_batch(g)
        for l, layer in enumerate(self.layers):
            h = layer(pg, h, w)
            if l != len(self.layers) - 1:
                h = L.ReluL2NormFunction.apply(h, 1e-12)  # F.relu + F.normalize (models.py:168-169)
        return h


class CrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss(weight=class_weights)`` on the CUDA kernels (model_train.py:171,327)."""

    def __init__(self, weight=None):
        super().__init__()
        self.register_buffer("weight", weight)

    def forward(self, logits, labels):
        return L.cross_entropy(logits, labels, self.weight)
