"""Synthetic PubLayNet-shaped page graphs (host side, numpy only).

The reference builds one graph per PDF page: nodes are word boxes, edges come
from a spatial k-NN search, the edge weight is ``1 - d/max_page(d)`` with ``d``
the rectangle distance, node features are the 13 "BBOX" numbers.  There is no
network / dataset here, so this module generates pages with the same shapes
and value ranges.  It is input data only -- no arithmetic of the hot path lives
here.

Reference behaviour restated (file:line under /root/reference):
  * rectangle distance ............ src/components/graphs/utils.py:56-88
  * k-NN edge direction (nbr->node) src/components/graphs/builder.py:286-290
  * edge weight 1 - d/max(d) ...... src/components/graphs/loader.py:332-344
  * BBOX features (9 geom + 4 hist) src/components/nlp/bbox.py:49-111,117-122
  * page size 1700x2200 px ........ 612x792 pt / SCALE_FACTOR 0.36 (src/utils/const.py:69)
  * labels stored as float32 ...... src/components/graphs/loader.py:348-354
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

PAGE_W = 1700
PAGE_H = 2200
NUM_CLASSES = 9
BBOX_FEATS = 13
# TEXT-dominant skewed class prior (9 converted classes)
_CLASS_PRIOR = np.array([0.55, 0.08, 0.06, 0.05, 0.02, 0.04, 0.12, 0.05, 0.03])


@dataclass
class PageGraph:
    """One page: COO edges (int32, sorted by (src, dst), unique), fp32 weights,
    [n, 13] fp32 features, float32 labels (the reference stores them as float)."""

    num_nodes: int
    src: np.ndarray  # int32 [E]
    dst: np.ndarray  # int32 [E]
    weight: np.ndarray  # float32 [E]
    feat: np.ndarray  # float32 [n, 13]
    label: np.ndarray  # float32 [n]
    bboxs: np.ndarray  # int64 [n, 4]

    @property
    def num_edges(self) -> int:
        return int(self.src.shape[0])


def rect_distance_matrix(b: np.ndarray) -> np.ndarray:
    """All-pairs rectangle distance D[a, b] = distance(rectA=b[a], rectB=b[b]).

    Vectorised restatement of ``distance`` (graphs/utils.py:56-88), including
    its branch order and the ``int(sqrt(.))`` truncation on the corner cases.
    Returns float64 (``inf`` where the reference returns ``inf``).
    """
    b = np.asarray(b, dtype=np.int64)
    A0, A1, A2, A3 = (b[:, i][:, None] for i in range(4))
    B0, B1, B2, B3 = (b[:, i][None, :] for i in range(4))
    left = (B2 - A0) <= 0
    bottom = (A3 - B1) <= 0
    right = (A2 - B0) <= 0
    top = (B3 - A1) <= 0
    vp = (A0 <= B2) & (B0 <= A2)
    hp = (A1 <= B3) & (B1 <= A3)
    inter = vp & hp

    def corner(dx, dy):
        return np.floor(np.sqrt((dx * dx + dy * dy).astype(np.float64)))

    n = b.shape[0]
    D = np.full((n, n), np.inf, dtype=np.float64)
    # apply branches in REVERSE priority so earlier branches overwrite later ones
    D = np.where(top, (A1 - B3).astype(np.float64), D)
    D = np.where(bottom, (B1 - A3).astype(np.float64), D)
    D = np.where(right, (B0 - A2).astype(np.float64), D)
    D = np.where(left, (A0 - B2).astype(np.float64), D)
    D = np.where(right & top, corner(B0 - A2, B3 - A1), D)
    D = np.where(bottom & right, corner(B0 - A2, B1 - A3), D)
    D = np.where(left & bottom, corner(B2 - A0, B1 - A3), D)
    D = np.where(top & left, corner(B2 - A0, B3 - A1), D)
    D = np.where(inter, 0.0, D)
    return D


def _layout_boxes(rng: np.random.Generator, n: int) -> np.ndarray:
    """Integer pixel word boxes laid out as text lines on a 1700x2200 page."""
    boxes = np.empty((n, 4), dtype=np.int64)
    x_left, x_right = 150, PAGE_W - 150
    y = 160
    x = x_left
    line_h = int(rng.integers(25, 36))
    for i in range(n):
        w = int(rng.integers(20, 151))
        if x + w > x_right:
            x = x_left
            y += line_h + int(rng.integers(8, 30))
            line_h = int(rng.integers(25, 36))
            if y + 40 > PAGE_H - 100:  # wrap to a "second column" jittered start
                y = 160 + int(rng.integers(0, 20))
        h = line_h - int(rng.integers(0, 4))
        boxes[i] = (x, y, x + w, y + h)
        # ~3% of words overlap the next one (rect distance 0 -> edge weight exactly 1)
        x += w + (int(rng.integers(8, 22)) if rng.random() > 0.03 else -int(rng.integers(2, 10)))
    return boxes


def _bbox_features(rng: np.random.Generator, boxes: np.ndarray) -> np.ndarray:
    """[w, h, cx, cy, w*h, x0, y0, x1, y1] + 4-bin char histogram (bbox.py:49-111)."""
    x0, y0, x1, y1 = (boxes[:, i] for i in range(4))
    w = x1 - x0
    h = y1 - y0
    cx = x1 - (w / 2).astype(np.int64)  # bbox[2] - int(width/2)
    cy = y1 - (h / 2).astype(np.int64)
    shape = np.stack([w, h, cx, cy, w * h, x0, y0, x1, y1], axis=1).astype(np.float64)
    n = boxes.shape[0]
    kind = rng.choice(4, size=n, p=[0.70, 0.12, 0.15, 0.03])
    hist = np.zeros((n, 4), dtype=np.float64)
    hist[kind == 0, 0] = 1.0  # pure literals
    hist[kind == 1, 1] = 1.0  # pure numbers
    mixed = kind == 2
    m = rng.dirichlet([4.0, 2.0, 1.0], size=int(mixed.sum()))
    hist[mixed, :3] = m
    hist[kind == 3, 3] = 1.0  # empty text
    return np.concatenate([shape, hist], axis=1).astype(np.float32)


def make_page(seed: int, n: int = 300, k: int = 10, bidirectional: bool = False) -> PageGraph:
    """One synthetic page graph.

    ``bidirectional=False`` (headline): directed k-NN, every node receives an
    edge from each of its k nearest boxes => in-degree exactly k, structure not
    symmetric.  ``bidirectional=True``: the k-NN edges are symmetrised and
    de-duplicated like ``to_simple`` + ``to_bidirected`` (loader.py:319-320).
    Edges are unique and sorted by (src, dst) as ``to_simple`` leaves them.
    """
    rng = np.random.default_rng(seed)
    boxes = _layout_boxes(rng, n)
    D = rect_distance_matrix(boxes)
    Dk = D.copy()
    np.fill_diagonal(Dk, np.inf)
    kk = min(k, n - 1)
    nbr = np.argsort(Dk, axis=1, kind="stable")[:, :kk]  # [n, kk] nearest boxes of each node
    dst = np.repeat(np.arange(n, dtype=np.int64), kk)
    src = nbr.reshape(-1).astype(np.int64)
    if bidirectional:
        s2 = np.concatenate([src, dst])
        d2 = np.concatenate([dst, src])
        key = np.unique(s2 * n + d2)
        src, dst = key // n, key % n
    else:
        order = np.argsort(src * n + dst, kind="stable")
        src, dst = src[order], dst[order]
    d = D[src, dst]
    m = d.max() if d.size else 1.0
    if not np.isfinite(m) or m <= 0:
        m = 1.0
    weight = (1.0 - d / m).astype(np.float32)
    feat = _bbox_features(rng, boxes)
    label = rng.choice(NUM_CLASSES, size=n, p=_CLASS_PRIOR / _CLASS_PRIOR.sum()).astype(np.float32)
    return PageGraph(n, src.astype(np.int32), dst.astype(np.int32), weight, feat, label, boxes)


def page_sizes(num_pages: int, base_seed: int = 42, ragged: bool = False, n: int = 300) -> List[int]:
    if not ragged:
        return [n] * num_pages
    rng = np.random.default_rng(base_seed + 10_000_019)
    s = np.clip(np.round(rng.normal(300, 80, size=num_pages)), 40, 900).astype(int)
    return [int(v) for v in s]


def make_pages(
    num_pages: int,
    base_seed: int = 42,
    n: int = 300,
    k: int = 10,
    bidirectional: bool = False,
    ragged: bool = False,
    distinct: Optional[int] = None,
) -> List[PageGraph]:
    """``num_pages`` pages with seeds ``base_seed + i``.

    ``distinct`` bounds the number of *different* pages generated (the rest are
    repeats in round-robin order) so that very large batches can be built
    quickly; the tensors are still materialised at full size.
    """
    sizes = page_sizes(num_pages, base_seed, ragged, n)
    cache = {}
    out = []
    for i in range(num_pages):
        j = i if distinct is None else i % max(1, distinct)
        if j not in cache:
            cache[j] = make_page(base_seed + j, sizes[j], k, bidirectional)
        out.append(cache[j])
    return out


def random_multigraph(seed: int, n: int, e: int, with_isolated: bool = True):
    """Unstructured directed multigraph (duplicates + self loops allowed, random
    edge order, optional zero-in-degree nodes) for the edge-case tests."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, size=e).astype(np.int32)
    hi = max(1, n - (n // 5 if with_isolated else 0))
    dst = rng.integers(0, hi, size=e).astype(np.int32)
    w = rng.random(e).astype(np.float32)
    if e >= 2:
        w[0] = 0.0
        w[1] = 1.0
    return src, dst, w
