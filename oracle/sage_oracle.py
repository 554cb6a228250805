"""fp32 CPU restatement of the reference GNN layers -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/components/graphs/models.py line by line, with the
DGL calls replaced by plain tensor arithmetic (``index_select`` + ``index_add``);
gradients come from torch autograd exactly as in the reference.  See
``oracle/__init__.py`` for what is and is not pinned.

Graph arguments are duck-typed: anything exposing ``edges() -> (src, dst)``,
``num_nodes()``, ``edata['feat']`` and (for ``GcnSAGE.forward``) ``ndata['feat']``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------
# DGL primitives as tensor ops
# --------------------------------------------------------------------------
def in_degrees(dst: torch.Tensor, n: int) -> torch.Tensor:
    """models.py:75 ``g.in_degrees()``: multiplicity-counting, id dtype (int32)."""
    return torch.bincount(dst.long(), minlength=n).to(torch.int32)


def u_mul_e_sum(src, dst, w, h, n):
    """models.py:53-54 ``update_all(fn.u_mul_e('h','feat','m'), fn.sum('m','h'))``.

    out[v,:] = sum over edges e=(u->v) of h[u,:] * w[e]; rows without incoming
    edges stay exactly 0.  ``w`` is [E] and broadcasts over features.
    """
    m = h.index_select(0, src.long()) * w.reshape(-1, 1)
    out = torch.zeros((n, h.shape[1]), dtype=h.dtype, device=h.device)
    return out.index_add(0, dst.long(), m)


def u_mul_e_mean(src, dst, w, h, n):
    """models.py:149 ``fn.mean``: sum / clamp(in_degree, 1)."""
    s = u_mul_e_sum(src, dst, w, h, n)
    deg = in_degrees(dst, n).to(h.dtype).clamp(min=1).unsqueeze(1)
    return s / deg


def _graph_parts(g):
    src, dst = g.edges()
    return src, dst, int(g.num_nodes())


# --------------------------------------------------------------------------
# GcnSAGELayer / GcnSAGE   (models.py:15-116)
# --------------------------------------------------------------------------
class OracleGcnSAGELayer(nn.Module):
    def __init__(self, in_feats, out_feats, activation, dropout, bias=True, use_pp=False, use_lynorm=True):
        super().__init__()
        self.linear = nn.Linear(2 * in_feats, out_feats, bias=bias)  # models.py:27
        self.activation = activation
        self.use_pp = use_pp
        self.dropout = nn.Dropout(p=dropout) if dropout else 0.0  # models.py:30-33
        if use_lynorm:
            self.lynorm = nn.LayerNorm(out_feats, elementwise_affine=True)  # models.py:34-35
        else:
            self.lynorm = lambda x: x
        self.reset_parameters()

    def reset_parameters(self):  # models.py:40-44
        stdv = 1.0 / math.sqrt(self.linear.weight.size(1))
        self.linear.weight.data.uniform_(-stdv, stdv)
        if self.linear.bias is not None:
            self.linear.bias.data.uniform_(-stdv, stdv)

    def get_norm(self, g):  # models.py:74-78
        _, dst, n = _graph_parts(g)
        norm = 1.0 / in_degrees(dst, n).float().unsqueeze(1)
        norm[torch.isinf(norm)] = 0
        return norm.to(self.linear.weight.device)

    def concat(self, h, ah, norm):  # models.py:69-72 -- self block first
        return torch.cat((h, ah * norm), dim=1)

    def forward(self, g, h, relu_mask=None):  # models.py:46-67
        """``relu_mask`` (test-only): evaluate the activation with a GIVEN on/off pattern
        (h * mask) instead of ``activation(h)``.  ReLU's derivative is discontinuous at 0, so
        two fp32 implementations can legitimately disagree on elements whose pre-activation is
        within rounding noise of 0; the parity tests compare gradients under the same pattern
        and separately check that the patterns differ only at such elements."""
        if not self.use_pp:
            src, dst, n = _graph_parts(g)
            norm = self.get_norm(g)
            ah = u_mul_e_sum(src, dst, g.edata["feat"], h, n)
            h = self.concat(h, ah, norm)
        if self.dropout:
            h = self.dropout(h)
        h = self.linear(h)
        h = self.lynorm(h)
        self.last_preact = h.detach()
        if self.activation:
            h = self.activation(h) if relu_mask is None else h * relu_mask.to(h.dtype)
        return h


class OracleGcnSAGE(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout, use_pp=False):
        super().__init__()
        self.layers = nn.ModuleList()
        self.dropout = nn.Dropout(dropout)
        self.layers.append(  # models.py:93-94
            OracleGcnSAGELayer(in_feats, n_hidden, activation=activation, dropout=dropout, use_pp=use_pp, use_lynorm=True)
        )
        for _ in range(n_layers - 2):  # models.py:96-99
            self.layers.append(
                OracleGcnSAGELayer(n_hidden, n_hidden, activation=activation, dropout=dropout, use_pp=False, use_lynorm=True)
            )
        self.layers.append(  # models.py:101-102
            OracleGcnSAGELayer(n_hidden, n_classes, activation=None, dropout=False, use_pp=False, use_lynorm=False)
        )

    def forward(self, g, relu_masks=None):  # models.py:105-116
        h = g.ndata["feat"]
        h = self.dropout(h)
        for i, layer in enumerate(self.layers):
            h = layer(g, h) if relu_masks is None else layer(g, h, relu_masks[i])
        return h


# --------------------------------------------------------------------------
# WeightedMeanSAGELayer / MeanSAGE   (models.py:118-170)
# --------------------------------------------------------------------------
class OracleWeightedMeanSAGELayer(nn.Module):
    def __init__(self, in_feat, out_feat):
        super().__init__()
        self.linear = nn.Linear(in_feat * 2, out_feat)  # models.py:131

    def forward(self, g, h, w):  # models.py:133-152
        src, dst, n = _graph_parts(g)
        h_N = u_mul_e_mean(src, dst, w, h, n)
        return self.linear(torch.cat([h, h_N], dim=1))


class OracleMeanSAGE(nn.Module):
    def __init__(self, in_feats, h_feats, num_classes, n_layers):
        super().__init__()
        self.n_layers = n_layers
        self.layers = nn.ModuleList()
        self.layers.append(OracleWeightedMeanSAGELayer(in_feats, h_feats))
        for _ in range(n_layers - 1):
            self.layers.append(OracleWeightedMeanSAGELayer(h_feats, h_feats))
        self.layers.append(OracleWeightedMeanSAGELayer(h_feats, num_classes))

    def forward(self, g, h, w):  # models.py:164-170
        for l, layer in enumerate(self.layers):
            h = layer(g, h, w)
            if l != len(self.layers) - 1:
                h = F.relu(h)
                h = F.normalize(h)
        return h


# --------------------------------------------------------------------------
# Train-step envelope (model_train.py:168-171, 320-332)
# --------------------------------------------------------------------------
def make_optimizer(model, lr=0.01, weight_decay=5e-4):
    return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)


def train_step(model, g, labels, optimizer, class_weights: Optional[torch.Tensor] = None):
    """logits -> CrossEntropy(labels.long()) -> zero_grad/backward/step.  Returns
    (loss, logits, {name: grad}) with the gradients as they were before the step."""
    loss_fn = nn.CrossEntropyLoss(weight=class_weights)
    logits = model(g)
    loss = loss_fn(logits, labels.type(torch.long))
    optimizer.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    optimizer.step()
    return loss.detach(), logits.detach(), grads


# --------------------------------------------------------------------------
# fp64 dense-adjacency tie-breaker (small graphs only)
# --------------------------------------------------------------------------
def dense_adjacency_fp64(src, dst, w, n):
    """A[v,u] = sum of w over edges u->v (duplicates add), fp64."""
    A = torch.zeros((n, n), dtype=torch.float64)
    A.index_put_((dst.long(), src.long()), w.double(), accumulate=True)
    return A


def gcn_sage_forward_fp64(model: OracleGcnSAGE, src, dst, w, feat):
    """Whole-model forward in fp64 with a dense (A o W) product."""
    n = feat.shape[0]
    A = dense_adjacency_fp64(src, dst, w, n)
    deg = torch.bincount(dst.long(), minlength=n).double()
    norm = torch.where(deg > 0, 1.0 / deg.clamp(min=1), torch.zeros_like(deg)).unsqueeze(1)
    h = feat.double()
    for layer in model.layers:
        x = torch.cat([h, (A @ h) * norm], dim=1)
        z = x @ layer.linear.weight.double().t()
        if layer.linear.bias is not None:
            z = z + layer.linear.bias.double()
        if isinstance(layer.lynorm, nn.LayerNorm):
            z = F.layer_norm(z, (z.shape[1],), layer.lynorm.weight.double(), layer.lynorm.bias.double(), layer.lynorm.eps)
        if layer.activation:
            z = layer.activation(z)
        h = z
    return h
