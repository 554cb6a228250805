"""``PageGraphBatch`` -- the graph argument of the B200 layers.

Stands in for the ``dgl.batch``'ed DGLGraph the reference hands to its layers
(/root/reference/src/models/model_train.py:297,320): a block-diagonal union of
page graphs with int32 ids (builder.py:425), ``ndata['feat']`` node features,
``edata['feat']`` edge weights (loader.py:332-344) and ``ndata['label']``.

Differences that matter for speed, not for results:
  * the CSC (rows = destination, used by the forward aggregation) and the CSR
    (rows = source, used by the backward) are built on the GPU once per batch by
    ``gte_csx_from_coo`` (or assembled from device-resident per-page formats by
    ``PagePool.batch``) instead of lazily inside every ``update_all``/backward;
  * the degree normaliser ``1/in_deg`` (models.py:74-78) is computed once per
    batch instead of once per layer.

The object exposes the slice of the DGL graph API that the reference layers
touch (``ndata``, ``edata``, ``edges()``, ``num_nodes()``, ``in_degrees()``,
``batch_num_nodes()``, ``local_var()``, ``local_scope()``) so the layers accept
it wherever they accepted a DGLGraph, and ``as_page_graph_batch`` converts any
DGL-like object (``g.edges()``, ``g.num_nodes()``, ``g.ndata/edata``) once and
caches the result on it.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import GteError


class PageGraphBatch:
    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int,
                 batch_num_nodes: Optional[Sequence[int]] = None, batch_num_edges: Optional[Sequence[int]] = None):
        if src.dtype != torch.int32 or dst.dtype != torch.int32:
            raise GteError("PageGraphBatch: edge ids must be int32 (builder.py:425)")
        if not src.is_cuda:
            raise GteError("PageGraphBatch lives on a CUDA device (no CPU path); use .to('cuda') on the inputs")
        self._src = src.contiguous()
        self._dst = dst.contiguous()
        self._n = int(num_nodes)
        self._bn = list(batch_num_nodes) if batch_num_nodes is not None else [self._n]
        self._be = list(batch_num_edges) if batch_num_edges is not None else [int(src.numel())]
        self.ndata: Dict[str, torch.Tensor] = {}
        self.edata: Dict[str, torch.Tensor] = {}
        self._cache: Dict[str, object] = {}  # shared by local_var() clones: formats are structure-only

    # ------------------------------------------------------ constructors --
    @classmethod
    def from_pages(cls, pages, device="cuda", pin: bool = True) -> "PageGraphBatch":
        """``dgl.batch(pages).to(device)``: list-order concatenation with node-id
        offsets (model_train.py:297).  ``pages`` are host ``synth.PageGraph``-like
        objects (``num_nodes, src, dst, weight, feat, label``)."""
        host = batch_pages_host(pages, pin=pin)
        return cls.from_host(host, device)

    @classmethod
    def from_host(cls, host: Dict[str, torch.Tensor], device="cuda") -> "PageGraphBatch":
        dev = torch.device(device)
        nb = host["src"].is_pinned()
        g = cls(host["src"].to(dev, non_blocking=nb), host["dst"].to(dev, non_blocking=nb), int(host["num_nodes"]),
                host["batch_num_nodes"], host["batch_num_edges"])
        if "weight" in host:
            g.edata["feat"] = host["weight"].to(dev, non_blocking=nb)
        if "feat" in host:
            g.ndata["feat"] = host["feat"].to(dev, non_blocking=nb)
        if "label" in host:
            g.ndata["label"] = host["label"].to(dev, non_blocking=nb)
        return g

    # ------------------------------------------------- DGL-like surface --
    def num_nodes(self) -> int:
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self) -> int:
        return int(self._src.numel())

    number_of_edges = num_edges

    def edges(self):
        return self._src, self._dst

    @property
    def device(self):
        return self._src.device

    def batch_num_nodes(self):
        return torch.tensor(self._bn, dtype=torch.int64)

    def batch_num_edges(self):
        return torch.tensor(self._be, dtype=torch.int64)

    @property
    def batch_size(self) -> int:
        return len(self._bn)

    def in_degrees(self) -> torch.Tensor:
        """Multiplicity-counting in-degrees in the id dtype (models.py:75)."""
        indptr = self.csc()[0]
        return indptr[1:] - indptr[:-1]

    def local_var(self) -> "PageGraphBatch":
        g = PageGraphBatch.__new__(PageGraphBatch)
        g._src, g._dst, g._n, g._bn, g._be = self._src, self._dst, self._n, self._bn, self._be
        g.ndata, g.edata = dict(self.ndata), dict(self.edata)
        g._cache = self._cache
        return g

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = dict(self.ndata), dict(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    def to(self, device) -> "PageGraphBatch":
        dev = torch.device(device)
        if dev == self.device:
            return self
        if dev.type != "cuda":
            raise GteError("PageGraphBatch.to: only CUDA devices are supported (no CPU path)")
        g = PageGraphBatch(self._src.to(dev), self._dst.to(dev), self._n, self._bn, self._be)
        g.ndata = {k: v.to(dev) for k, v in self.ndata.items()}
        g.edata = {k: v.to(dev) for k, v in self.edata.items()}
        return g

    # ------------------------------------------------- sparse formats ----
    def csc(self):
        """(indptr, indices=src, eid) compressed over destinations -- forward."""
        c = self._cache.get("csc")
        if c is None:
            c = ops.csx_from_coo(self._dst, self._src, self._n, bad=self._bad_ids())
            self._cache["csc"] = c
        return c

    def csr(self):
        """(indptr, indices=dst, eid) compressed over sources -- backward."""
        c = self._cache.get("csr")
        if c is None:
            c = ops.csx_from_coo(self._src, self._dst, self._n, bad=self._bad_ids())
            self._cache["csr"] = c
        return c

    def _bad_ids(self) -> torch.Tensor:
        """device flag raised by the format builders when a node id lies outside [0, num_nodes)"""
        f = self._cache.get("bad_ids")
        if f is None:
            f = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._cache["bad_ids"] = f
        return f

    def set_formats(self, csc=None, csr=None, w_csc=None, w_csr=None, w_src: Optional[torch.Tensor] = None):
        """Install pre-built formats (``PagePool.batch``)."""
        if csc is not None:
            self._cache["csc"] = csc
        if csr is not None:
            self._cache["csr"] = csr
        if w_src is not None:
            key = (w_src.data_ptr(), w_src._version, int(w_src.numel()))
            if w_csc is not None:
                self._cache["w_csc"] = (key, w_csc, w_src)
            if w_csr is not None:
                self._cache["w_csr"] = (key, w_csr, w_src)

    def norm(self, mode: int = _lib.GTE_NORM_INV_DEG_ZERO) -> torch.Tensor:
        """1/in_degree with inf -> 0 (``get_norm``, models.py:74-78), [N] fp32."""
        k = f"norm{mode}"
        v = self._cache.get(k)
        if v is None:
            v = ops.degree_norm(self.csc()[0], mode)
            self._cache[k] = v
        return v

    def pages(self):
        """(page_off int32 [P+1] on device, P, max page nodes, max page edges) for the staged SpMM, or None
        when the batch has no page structure worth staging (a single huge graph)."""
        v = self._cache.get("pages", 0)
        if v == 0:
            v = page_table(self._bn, self._be, self._n, self.device)
            self._cache["pages"] = v
        return v

    def _weights(self, which: str, w: torch.Tensor) -> torch.Tensor:
        w = _edge_weight_1d(w, self.num_edges())
        key = (w.data_ptr(), w._version, int(w.numel()))
        hit = self._cache.get(which)
        if hit is not None and hit[0] == key:
            return hit[1]
        eid = (self.csc() if which == "w_csc" else self.csr())[2]
        out = ops.gather_f32(w, eid)
        self._cache[which] = (key, out, w)  # holding `w` keeps its address from being recycled
        return out

    def prepare(self, w: Optional[torch.Tensor]) -> bool:
        """Batch assembly in one kernel (``gte_build_page_formats``): CSC, CSR, the degree normaliser and the packed
        edges of both directions for the edge weights ``w`` -- everything the layers ask this object for.  Used when
        the batch has a page table, the largest page fits in shared memory and nothing was built yet; otherwise the
        individual builders run lazily as before.  Returns True when the formats are in place."""
        if "csc" in self._cache or "csr" in self._cache:
            return False
        pages = self.pages()
        if w is None or not ops.page_formats_supported(pages) or len(self._be) != pages[1] \
                or sum(self._be) != self.num_edges():
            return False
        w = _edge_weight_1d(w, self.num_edges())
        if len(pages) < 5:
            return False
        eoff = pages[4]
        csc, csr, norm, pk_csc, pk_csr, bad = ops.build_page_formats(self._src, self._dst, w, pages[0], eoff, pages[1], self._n,
                                                                     pages[2], pages[3])
        key = (w.data_ptr(), w._version, int(w.numel()))
        self._cache["csc"], self._cache["csr"] = csc, csr
        self._cache[f"norm{_lib.GTE_NORM_INV_DEG_ZERO}"] = norm
        self._cache["pk_csc"] = (key, pk_csc, w)
        self._cache["pk_csr"] = (key, pk_csr, w)
        self._cache["bad"] = bad  # device flag: 1 = an edge leaves its page (check_page_structure() reads it)
        return True

    def check_page_structure(self) -> None:
        """Host check (synchronises) of the flags the format builders raise on the device: node ids outside
        [0, num_nodes) (DGL raises when such a graph is created; here the offending edges are dropped / clamped so the
        kernels stay memory safe, and THIS call reports it) and, for the one-kernel batch assembly, an edge that
        leaves its page (the page table does not describe the graph)."""
        bad_ids = self._cache.get("bad_ids")
        if bad_ids is not None and int(bad_ids.item()) != 0:
            raise GteError(f"PageGraphBatch: edge endpoints outside [0, {self._n}) (num_nodes too small or corrupt ids)")
        bad = self._cache.get("bad")
        if bad is not None and int(bad.item()) != 0:
            raise GteError("PageGraphBatch: an edge leaves its page (batch_num_nodes / batch_num_edges do not describe "
                           "this graph); build it without page sizes to use the generic kernels")

    validate = check_page_structure

    def packed_edges(self, which: str, w: torch.Tensor) -> "ops.PackedEdges":
        """Edges of the CSC (``which='csc'``, forward) or of the CSR with the source-side scale
        ``norm[dst]`` folded in (``'csr'``, backward), packed once per batch for the page kernel."""
        w = _edge_weight_1d(w, self.num_edges())
        key = (w.data_ptr(), w._version, int(w.numel()))
        name = "pk_" + which
        hit = self._cache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        indptr, indices, eid = self.csc() if which == "csc" else self.csr()
        pk = ops.paged_pack_edges(indptr, indices, w, self.pages(), eid=eid,
                                  pre_scale=self.norm() if which == "csr" else None)
        self._cache[name] = (key, pk, w)
        return pk

    def weights_csc(self, w: torch.Tensor) -> torch.Tensor:
        """Edge weights permuted into CSC row order (``edata['feat'][eid]``)."""
        return self._weights("w_csc", w)

    def weights_csr(self, w: torch.Tensor) -> torch.Tensor:
        return self._weights("w_csr", w)


def _edge_weight_1d(w: torch.Tensor, num_edges: int) -> torch.Tensor:
    if w.dtype != torch.float32:
        w = w.float()
    if w.dim() != 1:
        if w.dim() == 2 and w.shape[1] == 1:
            w = w.reshape(-1)
        else:
            raise GteError(f"edge weights must be [E] (loader.py:344), got {tuple(w.shape)}")
    if w.numel() != num_edges:
        raise GteError(f"edge weights: {w.numel()} values for {num_edges} edges")
    return w.contiguous()


def page_table(batch_num_nodes, batch_num_edges, num_nodes: int, device):
    """(page_off int32 [P+1] device tensor, P, max page nodes, max page edges, edge_off int32 [P+1] device tensor)
    or None (no useful page structure).  Pages are closed under edges, so the per-page edge counts hold for the CSC and the CSR."""
    bn = list(batch_num_nodes) if batch_num_nodes is not None else []
    be = list(batch_num_edges) if batch_num_edges is not None else []
    mx = max(bn) if bn else 0
    if not bn or len(be) != len(bn) or sum(bn) != num_nodes or mx > 1600:
        return None
    off = np.zeros(len(bn) + 1, dtype=np.int32)
    np.cumsum(np.asarray(bn, dtype=np.int64), out=off[1:])
    eoff = np.zeros(len(be) + 1, dtype=np.int64)
    np.cumsum(np.asarray(be, dtype=np.int64), out=eoff[1:])
    if eoff[-1] >= 2 ** 31:
        return None
    # [4] = edge offsets of the pages (dgl.batch concatenates edges in page order): the one-kernel batch assembly
    return (torch.from_numpy(off).to(device), len(bn), int(mx), int(max(be)), torch.from_numpy(eoff.astype(np.int32)).to(device))


def batch_pages_host(pages, pin: bool = True) -> Dict[str, torch.Tensor]:
    """Host-side ``dgl.batch``: concatenated COO with node offsets + features,
    weights and labels, in (optionally pinned) host tensors ready for one H2D copy."""
    n_tot = sum(p.num_nodes for p in pages)
    e_tot = sum(int(p.src.shape[0]) for p in pages)
    src = np.empty(e_tot, dtype=np.int32)
    dst = np.empty(e_tot, dtype=np.int32)
    w = np.empty(e_tot, dtype=np.float32)
    f = pages[0].feat.shape[1] if pages else 0
    feat = np.empty((n_tot, f), dtype=np.float32)
    label = np.empty(n_tot, dtype=np.float32)
    no = eo = 0
    bn, be = [], []
    for p in pages:
        e, n = int(p.src.shape[0]), int(p.num_nodes)
        np.add(p.src, no, out=src[eo:eo + e], casting="unsafe")
        np.add(p.dst, no, out=dst[eo:eo + e], casting="unsafe")
        w[eo:eo + e] = p.weight
        feat[no:no + n] = p.feat
        label[no:no + n] = p.label
        bn.append(n)
        be.append(e)
        no += n
        eo += e
    out = {"src": torch.from_numpy(src), "dst": torch.from_numpy(dst), "weight": torch.from_numpy(w),
           "feat": torch.from_numpy(feat), "label": torch.from_numpy(label)}
    if e_tot < 2 ** 31:  # the page table travels with the batch (captured steps refresh it per batch)
        po = np.zeros(len(bn) + 1, dtype=np.int32)
        eo = np.zeros(len(be) + 1, dtype=np.int32)
        np.cumsum(np.asarray(bn, dtype=np.int64), out=po[1:])
        np.cumsum(np.asarray(be, dtype=np.int64), out=eo[1:])
        out["page_off"], out["edge_off"] = torch.from_numpy(po), torch.from_numpy(eo)
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    out["num_nodes"] = n_tot
    out["batch_num_nodes"] = bn
    out["batch_num_edges"] = be
    return out


def as_page_graph_batch(g) -> PageGraphBatch:
    """Accept a ``PageGraphBatch`` or any DGL-like graph; convert once, cache on the object."""
    if isinstance(g, PageGraphBatch):
        return g
    cached = getattr(g, "_gte_batch", None)
    if cached is not None:
        pg = cached
    else:
        src, dst = g.edges()
        if not src.is_cuda:
            raise GteError("graph is on the CPU: move it to a CUDA device first (g.to('cuda')); there is no CPU path")
        try:
            bn = [int(v) for v in g.batch_num_nodes().tolist()]
            be = [int(v) for v in g.batch_num_edges().tolist()]
        except Exception:
            bn = be = None
        pg = PageGraphBatch(src.to(torch.int32), dst.to(torch.int32), int(g.num_nodes()), bn, be)
        try:
            g._gte_batch = pg
        except Exception:
            pass
    # feature dicts are re-read every time: callers overwrite ndata between steps
    pg = pg.local_var()
    pg.ndata = dict(g.ndata)
    pg.edata = dict(g.edata)
    return pg
