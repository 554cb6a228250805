"""Shared builders for the parity tests (test infrastructure)."""
import numpy as np
import torch

from oracle import sage_oracle as so
from oracle.csx import OracleGraph, batch_coo


def oracle_graph_from_pages(pages):
    src, dst, w, noff, eoff = batch_coo(pages)
    feat = np.concatenate([p.feat for p in pages], 0)
    g = OracleGraph(src, dst, int(noff[-1]), w, feat)
    g.ndata["label"] = torch.from_numpy(np.concatenate([p.label for p in pages]))
    return g


def oracle_graph_from_golden(d):
    g = OracleGraph(d["src"], d["dst"], int(d["num_nodes"]), d["weight"], d["feat"])
    if "label" in d:
        g.ndata["label"] = torch.from_numpy(d["label"])
    return g


def cuda_graph_from_arrays(src, dst, n, weight=None, feat=None, label=None, device="cuda"):
    from gnn_tableextraction_b200 import PageGraphBatch

    t = lambda a, dt: torch.as_tensor(np.asarray(a)).to(dt).to(device)
    g = PageGraphBatch(t(src, torch.int32), t(dst, torch.int32), int(n))
    if weight is not None:
        g.edata["feat"] = t(weight, torch.float32)
    if feat is not None:
        g.ndata["feat"] = t(feat, torch.float32)
    if label is not None:
        g.ndata["label"] = t(label, torch.float32)
    return g


def copy_state(dst_model, src_state, device=None):
    sd = {k: (v.to(device) if device else v) for k, v in src_state.items()}
    dst_model.load_state_dict(sd)
    return dst_model


def cuda_forward_with_masks(cuda_model, g):
    """logits + the ReLU on/off pattern of every layer of the CUDA model (None where no activation)."""
    masks = [None] * len(cuda_model.layers)
    hooks = []
    for i, layer in enumerate(cuda_model.layers):
        if layer.activation is not None:
            hooks.append(layer.register_forward_hook(
                lambda m, inp, out, i=i: masks.__setitem__(i, (out.detach() > 0).cpu())))
    logits = cuda_model(g)
    for h in hooks:
        h.remove()
    return logits, masks


def check_relu_patterns(oracle_model, masks, tol=1e-5):
    """The CUDA and oracle activation patterns may differ only where the oracle's pre-activation is
    within `tol` (relative to the layer's largest pre-activation) of the ReLU kink.  Returns #flips."""
    flips = 0
    for layer, m in zip(oracle_model.layers, masks):
        if m is None:
            continue
        pre = layer.last_preact
        diff = (pre > 0) != m
        flips += int(diff.sum())
        if diff.any():
            assert pre[diff].abs().max().item() <= tol * pre.abs().max().item(), "activation pattern differs away from 0"
    return flips
