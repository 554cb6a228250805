"""Minimal CPU stand-in for the slice of DGL the reference layers touch.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  DGL is an un-vendored,
un-pinned third-party dependency of the reference (no version in
/root/reference/requirements.txt or setup.py; API usage implies 0.6-0.9) and is
not installable here.  This file restates, from DGL's published behaviour, the
handful of primitives that ``/root/reference/src/components/graphs/models.py``
calls, so that the reference file itself can be imported and executed by
``tests/golden/make_golden.py``.  Semantics encoded (each has a unit test in
``tests/test_oracle.py``):

  S1  ``fn.u_mul_e('h','w','m')`` with h:[N,F], w:[E] -> w is viewed as [E,1] and
      broadcast over the feature axis (models.py:53, :149).
  S2  ``fn.sum`` writes a zero-initialised [N,F] output; zero-in-degree rows are 0.
  S3  ``fn.mean`` = sum / clamp(in_degree, min=1).
  S4  ``in_degrees()`` counts edges with multiplicity, returned in the id dtype.
  S5  ``dgl.batch``: nodes / edges concatenated in list order, node ids shifted by
      the cumulative node count, ndata/edata concatenated on dim 0,
      ``batch_num_nodes/edges`` recorded.
  S6  messages are reduced per destination over the incoming edges; the float
      summation order inside a row is edge order (stable CSC) -- only relevant at
      the 1e-7 level.
  S7  autograd: d(sum)/dh is the same reduction over the reversed edges; edge
      data that does not require grad receives none.
  S8  ``local_var()`` / ``local_scope()`` isolate ndata/edata writes from the caller.
"""
from __future__ import annotations

import contextlib
import sys
import types
from typing import Dict, List, Sequence

import torch


class _Msg:
    def __init__(self, kind, lhs, rhs, out):
        self.kind, self.lhs, self.rhs, self.out = kind, lhs, rhs, out


class _Red:
    def __init__(self, kind, msg, out):
        self.kind, self.msg, self.out = kind, msg, out


def u_mul_e(lhs_field, rhs_field, out):
    return _Msg("u_mul_e", lhs_field, rhs_field, out)


def copy_u(u, out):
    return _Msg("copy_u", u, None, out)


def _sum(msg, out):
    return _Red("sum", msg, out)


def _mean(msg, out):
    return _Red("mean", msg, out)


class ShimGraph:
    """Directed multigraph in COO form with per-node / per-edge feature dicts."""

    def __init__(self, src, dst, num_nodes, idtype=torch.int32, batch_nodes=None, batch_edges=None):
        self._src = torch.as_tensor(src).to(idtype).contiguous()
        self._dst = torch.as_tensor(dst).to(idtype).contiguous()
        self._n = int(num_nodes)
        self.idtype = idtype
        self.ndata: Dict[str, torch.Tensor] = {}
        self.edata: Dict[str, torch.Tensor] = {}
        self._bn = list(batch_nodes) if batch_nodes is not None else [self._n]
        self._be = list(batch_edges) if batch_edges is not None else [int(self._src.numel())]

    # structure ---------------------------------------------------------
    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self._src.numel())

    number_of_edges = num_edges

    def edges(self):
        return self._src, self._dst

    def batch_num_nodes(self):
        return torch.tensor(self._bn, dtype=torch.int64)

    def batch_num_edges(self):
        return torch.tensor(self._be, dtype=torch.int64)

    def in_degrees(self):  # S4
        return torch.bincount(self._dst.long(), minlength=self._n).to(self.idtype)

    def to(self, device):
        g = self._clone()
        g._src, g._dst = self._src.to(device), self._dst.to(device)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device) for k, v in self.edata.items()}
        return g

    # scoping (S8) ------------------------------------------------------
    def _clone(self):
        g = ShimGraph.__new__(ShimGraph)
        g._src, g._dst, g._n, g.idtype = self._src, self._dst, self._n, self.idtype
        g._bn, g._be = self._bn, self._be
        g.ndata, g.edata = dict(self.ndata), dict(self.edata)
        return g

    def local_var(self):
        return self._clone()

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = dict(self.ndata), dict(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    # message passing (S1, S2, S3, S6, S7) ------------------------------
    def update_all(self, message_func, reduce_func):
        src, dst = self._src.long(), self._dst.long()
        h = self.ndata[message_func.lhs]
        m = h.index_select(0, src)
        if message_func.kind == "u_mul_e":
            w = self.edata[message_func.rhs]
            if w.dim() < m.dim():
                w = w.reshape(w.shape + (1,) * (m.dim() - w.dim()))
            m = m * w
        out = torch.zeros((self._n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        out = out.index_add(0, dst, m)
        if reduce_func.kind == "mean":
            deg = self.in_degrees().to(out.dtype).clamp(min=1)
            out = out / deg.reshape((-1,) + (1,) * (out.dim() - 1))
        self.ndata[reduce_func.out] = out


def graph(data, num_nodes=None, idtype=torch.int32):
    u, v = data
    u, v = torch.as_tensor(u), torch.as_tensor(v)
    if num_nodes is None:
        num_nodes = int(max(u.max().item(), v.max().item())) + 1 if u.numel() else 0
    return ShimGraph(u, v, num_nodes, idtype)


def batch(graphs: Sequence[ShimGraph]) -> ShimGraph:  # S5
    offs, srcs, dsts, bn, be = 0, [], [], [], []
    for g in graphs:
        s, d = g.edges()
        srcs.append(s.long() + offs)
        dsts.append(d.long() + offs)
        bn.append(g.num_nodes())
        be.append(g.num_edges())
        offs += g.num_nodes()
    idt = graphs[0].idtype
    out = ShimGraph(torch.cat(srcs), torch.cat(dsts), offs, idt, bn, be)
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata:
        out.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return out


def install_as_dgl() -> types.ModuleType:
    """Register this shim as ``dgl`` in ``sys.modules`` so that the reference's
    ``models.py`` (``import dgl.function as fn``; ``from dgl.nn.pytorch.conv import
    SAGEConv``) imports unmodified.  Used by tests/golden/make_golden.py only."""
    dgl = types.ModuleType("dgl")
    fn = types.ModuleType("dgl.function")
    fn.u_mul_e, fn.copy_u, fn.sum, fn.mean = u_mul_e, copy_u, _sum, _mean
    nn_ = types.ModuleType("dgl.nn")
    nnp = types.ModuleType("dgl.nn.pytorch")
    conv = types.ModuleType("dgl.nn.pytorch.conv")

    class SAGEConv(torch.nn.Module):  # imported by the reference, never instantiated
        pass

    conv.SAGEConv = SAGEConv
    nnp.conv, nn_.pytorch = conv, nnp
    dgl.function, dgl.nn = fn, nn_
    dgl.graph, dgl.batch, dgl.DGLGraph = graph, batch, ShimGraph
    for name, mod in [
        ("dgl", dgl),
        ("dgl.function", fn),
        ("dgl.nn", nn_),
        ("dgl.nn.pytorch", nnp),
        ("dgl.nn.pytorch.conv", conv),
    ]:
        sys.modules[name] = mod
    return dgl
