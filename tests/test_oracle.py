"""CPU tests: the oracle against the golden vectors produced by the REFERENCE's own
models.py (tests/golden/make_golden.py), one unit test per assumed DGL semantic
(S1..S8 of oracle/dgl_shim.py), and the integer oracle for the sparse formats."""
import math

import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, TOL, load_golden, rel_err, sub
from helpers import oracle_graph_from_golden, oracle_graph_from_pages
from gnn_tableextraction_b200 import synth
from oracle import csx, dgl_shim
from oracle import sage_oracle as so


# ---------------------------------------------------------------- golden ----
@pytest.mark.parametrize("name", ["gcnsage_default_knn", "gcnsage_bidir_classw", "gcnsage_multigraph",
                                  "gcnsage_default_2400", "gcnsage_ragged_2400"])
def test_oracle_matches_reference_gcnsage(name):
    d = load_golden(name)
    inf, hid, ncls, nl = (int(v) for v in d["config"])
    model = so.OracleGcnSAGE(inf, hid, ncls, nl, F.relu, 0)
    state = sub(d, "state")
    assert set(model.state_dict().keys()) == set(state.keys())  # state_dict key parity with the reference
    model.load_state_dict(state)
    g = oracle_graph_from_golden(d)
    logits = model(g)
    cw = torch.from_numpy(d["class_w"]) if "class_w" in d else None
    loss = torch.nn.CrossEntropyLoss(weight=cw)(logits, g.ndata["label"].long())
    loss.backward()
    # same torch ops in the same order as the reference over the shim -> essentially bit-equal
    assert rel_err(logits, d["logits"]) < 1e-6
    assert abs(loss.item() - float(d["loss"])) < 1e-6 * max(1.0, abs(float(d["loss"])))
    grads = sub(d, "grad")
    for k, p in model.named_parameters():
        assert rel_err(p.grad, grads[k]) < 1e-6, k


def test_oracle_matches_reference_layers():
    d = load_golden("gcnsage_layers")
    g = oracle_graph_from_golden(d)
    cfg = {"ln_relu": (32, F.relu, True, True), "plain": (10, None, True, False), "nobias_relu": (16, F.relu, False, False)}
    for vn, (fo, act, bias, ln) in cfg.items():
        layer = so.OracleGcnSAGELayer(13, fo, act, 0.0, bias=bias, use_lynorm=ln)
        layer.load_state_dict(sub(d, f"{vn}.state"))
        h = torch.from_numpy(d["feat"]).clone().requires_grad_(True)
        y = layer(g, h)
        (y * torch.from_numpy(d[f"{vn}.upstream"])).sum().backward()
        assert rel_err(y, d[f"{vn}.out"]) < 1e-6, vn
        assert rel_err(h.grad, d[f"{vn}.dh"]) < 1e-6, vn
        for k, p in layer.named_parameters():
            assert rel_err(p.grad, d[f"{vn}.grad.{k}"]) < 1e-6, (vn, k)
    layer = so.OracleGcnSAGELayer(13, 12, F.relu, 0.0, use_pp=True)
    layer.load_state_dict(sub(d, "pp.state"))
    assert rel_err(layer(g, torch.from_numpy(d["pp.in"])), d["pp.out"]) < 1e-6


def test_oracle_matches_reference_meansage():
    d = load_golden("meansage")
    inf, hid, ncls, nl = (int(v) for v in d["config"])
    model = so.OracleMeanSAGE(inf, hid, ncls, nl)
    model.load_state_dict(sub(d, "state"))
    assert len(model.layers) == nl + 1  # models.py:158-162
    g = oracle_graph_from_golden(d)
    h = torch.from_numpy(d["feat"]).clone().requires_grad_(True)
    y = model(g, h, torch.from_numpy(d["weight"]))
    (y * torch.from_numpy(d["upstream"])).sum().backward()
    assert rel_err(y, d["out"]) < 1e-6
    assert rel_err(h.grad, d["dh"]) < 1e-6
    grads = sub(d, "grad")
    for k, p in model.named_parameters():
        assert rel_err(p.grad, grads[k]) < 1e-6, k


def test_oracle_fp32_vs_fp64_dense():
    """fp64 dense-adjacency tie-breaker: the fp32 oracle sits within 1e-5 of it."""
    d = load_golden("gcnsage_default_knn")
    model = so.OracleGcnSAGE(13, 218, 9, 3, F.relu, 0)
    model.load_state_dict(sub(d, "state"))
    g = oracle_graph_from_golden(d)
    ref64 = so.gcn_sage_forward_fp64(model, *g.edges(), g.edata["feat"], g.ndata["feat"])
    assert rel_err(model(g), ref64) < TOL


# ------------------------------------------------ assumed DGL semantics -----
def _tiny():
    # 5 nodes; node 4 has no incoming edge; edge (0->1) duplicated; self loop on 2
    src = torch.tensor([0, 0, 2, 3, 2, 1], dtype=torch.int32)
    dst = torch.tensor([1, 1, 2, 0, 3, 0], dtype=torch.int32)
    w = torch.tensor([0.5, 0.25, 1.0, 0.0, 2.0, 3.0])
    h = torch.arange(10, dtype=torch.float32).reshape(5, 2) + 1
    return src, dst, w, h


def test_S1_S2_u_mul_e_sum_broadcast_and_zero_rows():
    src, dst, w, h = _tiny()
    out = so.u_mul_e_sum(src, dst, w, h, 5)
    exp = torch.zeros(5, 2)
    for e in range(6):
        exp[dst[e]] += h[src[e]] * w[e]  # [E] weight broadcast over the feature axis
    assert torch.equal(out, exp)
    assert torch.all(out[4] == 0)  # zero in-degree -> exactly 0


def test_S3_mean_clamps_degree():
    src, dst, w, h = _tiny()
    s = so.u_mul_e_sum(src, dst, w, h, 5)
    m = so.u_mul_e_mean(src, dst, w, h, 5)
    deg = torch.tensor([2.0, 2.0, 1.0, 1.0, 1.0]).unsqueeze(1)  # node 4: clamp(0, 1) = 1
    assert torch.allclose(m, s / deg)


def test_S4_in_degrees_multiplicity_and_dtype():
    src, dst, _, _ = _tiny()
    deg = so.in_degrees(dst, 5)
    assert deg.dtype == torch.int32 and deg.tolist() == [2, 2, 1, 1, 0]
    g = dgl_shim.graph((src, dst), num_nodes=5)
    assert g.in_degrees().tolist() == [2, 2, 1, 1, 0]


def test_S5_batch_offsets_and_concat():
    pages = [synth.make_page(1, n=20, k=3), synth.make_page(2, n=30, k=3)]
    s, d, w, noff, eoff = csx.batch_coo(pages)
    assert noff.tolist() == [0, 20, 50] and eoff.tolist() == [0, 60, 150]
    assert np.array_equal(s[:60], pages[0].src) and np.array_equal(s[60:], pages[1].src + 20)
    assert np.array_equal(d[60:], pages[1].dst + 20) and np.array_equal(w[60:], pages[1].weight)
    gs = []
    for p in pages:
        g = dgl_shim.graph((torch.from_numpy(p.src), torch.from_numpy(p.dst)), num_nodes=p.num_nodes)
        g.ndata["feat"] = torch.from_numpy(p.feat)
        g.edata["feat"] = torch.from_numpy(p.weight)
        gs.append(g)
    bg = dgl_shim.batch(gs)
    assert bg.batch_num_nodes().tolist() == [20, 30] and bg.batch_num_edges().tolist() == [60, 90]
    assert np.array_equal(bg.edges()[0].numpy(), s) and np.array_equal(bg.edges()[1].numpy(), d)
    assert bg.ndata["feat"].shape == (50, 13)


def test_S6_csc_rows_keep_edge_order():
    src, dst, _, _ = _tiny()
    indptr, indices, eid = csx.csx_from_coo(dst.numpy(), src.numpy(), 5)
    assert indptr.tolist() == [0, 2, 4, 5, 6, 6]
    assert eid.tolist() == [3, 5, 0, 1, 2, 4]  # stable: ties by original edge position
    assert indices.tolist() == [3, 1, 0, 0, 2, 2]


def test_S7_backward_is_reverse_graph_spmm_no_edge_grad():
    src, dst, w, h = _tiny()
    h = h.clone().requires_grad_(True)
    w_ = w.clone()  # does not require grad
    up = torch.randn(5, 2, generator=torch.Generator().manual_seed(0))
    (so.u_mul_e_sum(src, dst, w_, h, 5) * up).sum().backward()
    rev = so.u_mul_e_sum(dst, src, w_, up, 5)  # same reduction over reversed edges
    assert torch.allclose(h.grad, rev)
    assert w_.grad is None


def test_S8_local_var_isolates_caller():
    src, dst, w, h = _tiny()
    g = dgl_shim.graph((src, dst), num_nodes=5)
    g.ndata["feat"] = h
    g.edata["feat"] = w
    lg = g.local_var()
    lg.ndata["h"] = h * 2
    lg.update_all(dgl_shim.u_mul_e("h", "feat", "m"), dgl_shim._sum("m", "h"))
    assert "h" not in g.ndata
    with g.local_scope():
        g.ndata["tmp"] = h
    assert "tmp" not in g.ndata


def test_get_norm_inf_to_zero():
    src, dst, w, h = _tiny()
    layer = so.OracleGcnSAGELayer(2, 3, None, 0.0)
    g = csx.OracleGraph(src, dst, 5, w, h)
    assert layer.get_norm(g).squeeze(1).tolist() == [0.5, 0.5, 1.0, 1.0, 0.0]  # models.py:75-76


def test_reset_parameters_range_and_rng_order():
    torch.manual_seed(0)
    layer = so.OracleGcnSAGELayer(13, 218, F.relu, 0.0)
    stdv = 1.0 / math.sqrt(26)
    assert layer.linear.weight.abs().max() <= stdv and layer.linear.bias.abs().max() <= stdv
    # the product modules consume the RNG in the same order (nn.Linear ctor, then reset_parameters)
    from gnn_tableextraction_b200.nn import GcnSAGE

    torch.manual_seed(5)
    a = so.OracleGcnSAGE(13, 24, 9, 3, F.relu, 0)
    torch.manual_seed(5)
    b = GcnSAGE(13, 24, 9, 3, F.relu, 0)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


# -------------------------------------------------------- integer oracle ----
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_csx_oracle_properties(seed):
    src, dst, _ = synth.random_multigraph(seed, 50, 400)
    indptr, indices, eid = csx.csx_from_coo(dst, src, 50)
    assert indptr[0] == 0 and indptr[-1] == 400 and np.all(np.diff(indptr) >= 0)
    assert sorted(eid.tolist()) == list(range(400))  # a permutation
    assert np.array_equal(indices, src[eid])
    for v in range(50):
        seg = eid[indptr[v]:indptr[v + 1]]
        assert np.all(dst[seg] == v) and np.all(np.diff(seg) > 0)


def test_csx_empty_and_ragged():
    indptr, indices, eid = csx.csx_from_coo(np.zeros(0, np.int32), np.zeros(0, np.int32), 4)
    assert indptr.tolist() == [0, 0, 0, 0, 0] and indices.size == 0 and eid.size == 0
    pages = synth.make_pages(6, ragged=True, k=4)
    assert len({p.num_nodes for p in pages}) > 1
    s, d, w, noff, eoff = csx.batch_coo(pages)
    assert s.max() < noff[-1] and len(s) == eoff[-1]


# ------------------------------------------------------ synthetic inputs ----
def _ref_distance(rectA, rectB):
    """scalar restatement of graphs/utils.py:56-88 for the test"""
    from math import inf, sqrt

    left = (rectB[2] - rectA[0]) <= 0
    bottom = (rectA[3] - rectB[1]) <= 0
    right = (rectA[2] - rectB[0]) <= 0
    top = (rectB[3] - rectA[1]) <= 0
    vp = rectA[0] <= rectB[2] and rectB[0] <= rectA[2]
    hp = rectA[1] <= rectB[3] and rectB[1] <= rectA[3]
    if vp and hp:
        return 0
    elif top and left:
        return int(sqrt((rectB[2] - rectA[0]) ** 2 + (rectB[3] - rectA[1]) ** 2))
    elif left and bottom:
        return int(sqrt((rectB[2] - rectA[0]) ** 2 + (rectB[1] - rectA[3]) ** 2))
    elif bottom and right:
        return int(sqrt((rectB[0] - rectA[2]) ** 2 + (rectB[1] - rectA[3]) ** 2))
    elif right and top:
        return int(sqrt((rectB[0] - rectA[2]) ** 2 + (rectB[3] - rectA[1]) ** 2))
    elif left:
        return rectA[0] - rectB[2]
    elif right:
        return rectB[0] - rectA[2]
    elif bottom:
        return rectB[1] - rectA[3]
    elif top:
        return rectA[1] - rectB[3]
    return inf


def test_synth_distance_matches_scalar_reference():
    p = synth.make_page(3, n=40)
    D = synth.rect_distance_matrix(p.bboxs)
    b = p.bboxs.tolist()
    for i in range(40):
        for j in range(40):
            assert D[i, j] == _ref_distance(b[i], b[j]), (i, j)


def test_synth_page_shape():
    p = synth.make_page(42)
    assert p.num_nodes == 300 and p.num_edges == 3000 and p.feat.shape == (300, 13)
    assert np.all(np.bincount(p.dst, minlength=300) == 10)  # in-degree exactly k
    key = p.src.astype(np.int64) * 300 + p.dst
    assert np.all(np.diff(key) > 0)  # unique, sorted by (src, dst) like to_simple
    assert (p.weight == 0).sum() >= 1 and p.weight.max() <= 1.0 and (p.weight == 1).sum() >= 1
    assert p.label.dtype == np.float32 and p.src.dtype == np.int32
    pb = synth.make_page(42, k=5, bidirectional=True)
    k2 = set(zip(pb.src.tolist(), pb.dst.tolist()))
    assert all((d, s) in k2 for s, d in k2)  # symmetric structure


# ------------------------------------------- either side of the layers ----
def test_bbox_oracle_matches_reference_golden():
    """oracle/bbox_oracle.py == the reference's own get_shape / get_histogram (tests/golden/make_bbox_golden.py),
    bit for bit, including the 75 rows that take the "keep sum 1" branch and the host-side class counting"""
    from oracle import bbox_oracle as bo

    d = np.load(os.path.join(GOLDEN, "bbox_features.npz"), allow_pickle=True)
    feat = bo.bbox_features(d["boxes"], d["counts"])
    assert feat.dtype == np.float32 and np.array_equal(feat, d["feat"])
    counts = np.asarray([bo.text_class_counts(str(t)) for t in d["texts"]], dtype=np.int32)
    assert np.array_equal(counts, d["counts"])
    assert np.allclose(feat[:, 9:].sum(1), 1.0, atol=1e-6)
    assert np.array_equal(bo.bbox_features(d["boxes"][:0], d["counts"][:0]), np.zeros((0, 13), np.float32))
    # the product's host-side counter agrees with the oracle's
    from gnn_tableextraction_b200.features import text_class_counts

    assert np.array_equal(text_class_counts([str(t) for t in d["texts"]]), d["counts"])


def test_page_predictions_oracle():
    from oracle import bbox_oracle as bo

    rng = np.random.default_rng(0)
    logits = rng.normal(size=(10, 4)).astype(np.float32)
    logits[3] = [1.0, 2.0, 2.0, 0.0]  # tie: first maximal index
    labels = logits.argmax(1).astype(np.float32)
    labels[0] = (labels[0] + 1) % 4
    preds, accs, mean = bo.page_predictions(logits, labels, [4, 6])
    assert preds[3] == 1 and accs == [0.75, 1.0] and abs(mean - 0.875) < 1e-12
