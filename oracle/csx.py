"""Integer oracle for the sparse-format construction -- TEST INFRASTRUCTURE ONLY.

The reference never builds CSC/CSR itself: DGL does it lazily on the first
``update_all`` (CSC, rows = destination) and on the first backward (CSR, rows =
source) of every ``dgl.batch``'ed graph (model_train.py:297,320,331).  DGL's
COO->CSR is a *stable* conversion that keeps an edge-id map, so the definition
used for "bit-exact" parity is: stable sort of the batched COO by the row key.

(SciPy's ``coo.tocsr()`` is not a valid oracle: it sums duplicate edges.)
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def csx_from_coo(key: np.ndarray, other: np.ndarray, n: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Compressed rows over ``key`` (dst for CSC, src for CSR).

    Returns (indptr int32 [n+1], indices int32 [E] = ``other`` in row order,
    eid int32 [E] = original edge position), ties inside a row in edge order.
    """
    key = np.asarray(key, dtype=np.int64)
    other = np.asarray(other, dtype=np.int32)
    order = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=n)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr.astype(np.int32), other[order].astype(np.int32), order.astype(np.int32)


def batch_coo(pages: Sequence) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """``dgl.batch`` semantics on COO (model_train.py:297): list-order
    concatenation with node ids shifted by the cumulative node count.
    Returns (src, dst, weight, node_offsets[P+1], edge_offsets[P+1])."""
    noff = np.zeros(len(pages) + 1, dtype=np.int64)
    eoff = np.zeros(len(pages) + 1, dtype=np.int64)
    srcs, dsts, ws = [], [], []
    for i, p in enumerate(pages):
        srcs.append(p.src.astype(np.int64) + noff[i])
        dsts.append(p.dst.astype(np.int64) + noff[i])
        ws.append(p.weight)
        noff[i + 1] = noff[i] + p.num_nodes
        eoff[i + 1] = eoff[i] + p.src.shape[0]
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    return cat(srcs, np.int32), cat(dsts, np.int32), cat(ws, np.float32), noff, eoff


class OracleGraph:
    """Plain COO graph holder satisfying the duck type of ``sage_oracle``."""

    def __init__(self, src, dst, n, weight=None, feat=None):
        import torch

        self._src = torch.as_tensor(np.asarray(src)).to(torch.int32)
        self._dst = torch.as_tensor(np.asarray(dst)).to(torch.int32)
        self._n = int(n)
        self.ndata, self.edata = {}, {}
        if weight is not None:
            self.edata["feat"] = torch.as_tensor(np.asarray(weight)).float()
        if feat is not None:
            self.ndata["feat"] = torch.as_tensor(np.asarray(feat)).float()

    def edges(self):
        return self._src, self._dst

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.numel())
