"""Generate the golden vectors under tests/golden/ from the REFERENCE's own code.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

What it does: registers ``oracle/dgl_shim.py`` as ``dgl`` (DGL itself is an
un-vendored, un-pinned dependency that cannot be installed here), imports
``/root/reference/src/components/graphs/models.py`` UNMODIFIED, instantiates the
reference modules, runs forward + CrossEntropy + backward on small seeded
graphs, and stores inputs, parameters, outputs and gradients as ``.npz``.

The fixtures pin the reference's layer composition (concat order, norm,
LayerNorm/activation placement, init); the DGL primitives underneath are the
shim's restatement (see oracle/__init__.py -- "parity unpinned" at that level).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import dgl_shim  # noqa: E402
from gnn_tableextraction_b200 import synth  # noqa: E402

REF_MODELS = "/root/reference/src/components/graphs/models.py"


def load_reference_models():
    dgl_shim.install_as_dgl()
    spec = importlib.util.spec_from_file_location("ref_models", REF_MODELS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def shim_batch(pages):
    gs = []
    for p in pages:
        g = dgl_shim.graph((torch.from_numpy(p.src), torch.from_numpy(p.dst)), num_nodes=p.num_nodes,
                           idtype=torch.int32)
        g.ndata["feat"] = torch.from_numpy(p.feat).float()
        g.ndata["label"] = torch.from_numpy(p.label)
        g.edata["feat"] = torch.from_numpy(p.weight)
        gs.append(g)
    return dgl_shim.batch(gs)


def pack(prefix, d):
    return {f"{prefix}.{k}": v.detach().cpu().numpy() for k, v in d.items()}


def run_gcnsage(ref, name, g, in_feats, hidden, classes, n_layers, seed, class_w=None):
    torch.manual_seed(seed)
    model = ref.GcnSAGE(in_feats, hidden, classes, n_layers, F.relu, 0)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    logits = model(g)
    labels = g.ndata["label"].type(torch.long)
    loss = torch.nn.CrossEntropyLoss(weight=class_w)(logits, labels)
    loss.backward()
    src, dst = g.edges()
    out = {
        "src": src.numpy().astype(np.int32), "dst": dst.numpy().astype(np.int32),
        "num_nodes": np.int64(g.num_nodes()), "weight": g.edata["feat"].numpy(), "feat": g.ndata["feat"].numpy(),
        "label": g.ndata["label"].numpy(), "logits": logits.detach().numpy(), "loss": loss.detach().numpy(),
        "batch_num_nodes": g.batch_num_nodes().numpy(),
        "config": np.array([in_feats, hidden, classes, n_layers], dtype=np.int64),
    }
    if class_w is not None:
        out["class_w"] = class_w.numpy()
    out.update(pack("state", state))
    out.update(pack("grad", {k: p.grad for k, p in model.named_parameters()}))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return {"file": name + ".npz", "kind": "GcnSAGE", "nodes": int(g.num_nodes()), "edges": int(src.numel())}


def multigraph_shim(seed, n, e, f):
    src, dst, w = synth.random_multigraph(seed, n, e)
    rng = np.random.default_rng(seed + 1)
    g = dgl_shim.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n, idtype=torch.int32)
    g.ndata["feat"] = torch.from_numpy(rng.standard_normal((n, f)).astype(np.float32))
    g.ndata["label"] = torch.from_numpy(rng.integers(0, 5, size=n).astype(np.float32))
    g.edata["feat"] = torch.from_numpy(w)
    return g


def main():
    ref = load_reference_models()
    manifest = []

    # 1. repo-default model (13 -> 218 -> 218 -> 9) on 3 small k-NN pages (directed, in-degree 10)
    pages = [synth.make_page(42 + i, n=n) for i, n in enumerate((48, 64, 40))]
    manifest.append(run_gcnsage(ref, "gcnsage_default_knn", shim_batch(pages), 13, 218, 9, 3, seed=0))

    # 2. bidirectional (to_simple + to_bidirected) variant, 4 layers, narrow hidden, class weights
    pages = [synth.make_page(7 + i, n=n, k=5, bidirectional=True) for i, n in enumerate((56, 72))]
    cw = torch.tensor([1.0] * 6 + [2.0] + [1.0] * 2)  # 'default' class weights (model_train.py:113-116)
    manifest.append(run_gcnsage(ref, "gcnsage_bidir_classw", shim_batch(pages), 13, 24, 9, 4, seed=1, class_w=cw))

    # 3. edge cases: multigraph with duplicate edges, self loops, zero-in-degree nodes, weights 0 and 1
    g = multigraph_shim(3, n=97, e=700, f=7)
    manifest.append(run_gcnsage(ref, "gcnsage_multigraph", g, 7, 20, 5, 3, seed=2))

    # 4. single layers: with / without LayerNorm, no bias, use_pp
    g = multigraph_shim(5, n=61, e=400, f=13)
    out = {"src": g.edges()[0].numpy().astype(np.int32), "dst": g.edges()[1].numpy().astype(np.int32),
           "num_nodes": np.int64(61), "weight": g.edata["feat"].numpy(), "feat": g.ndata["feat"].numpy()}
    torch.manual_seed(3)
    variants = {
        "ln_relu": ref.GcnSAGELayer(13, 32, F.relu, 0.0, bias=True, use_pp=False, use_lynorm=True),
        "plain": ref.GcnSAGELayer(13, 10, None, 0.0, bias=True, use_pp=False, use_lynorm=False),
        "nobias_relu": ref.GcnSAGELayer(13, 16, F.relu, 0.0, bias=False, use_pp=False, use_lynorm=False),
    }
    for vn, layer in variants.items():
        h = g.ndata["feat"].clone().requires_grad_(True)
        y = layer(g, h)
        up = torch.from_numpy(np.random.default_rng(11).standard_normal(tuple(y.shape)).astype(np.float32))
        (y * up).sum().backward()
        out[f"{vn}.out"] = y.detach().numpy()
        out[f"{vn}.upstream"] = up.numpy()
        out[f"{vn}.dh"] = h.grad.numpy()
        out.update(pack(f"{vn}.state", layer.state_dict()))
        out.update(pack(f"{vn}.grad", {k: p.grad for k, p in layer.named_parameters()}))
    # use_pp: input already [N, 2*in]
    layer = ref.GcnSAGELayer(13, 12, F.relu, 0.0, use_pp=True)
    hpp = torch.from_numpy(np.random.default_rng(12).standard_normal((61, 26)).astype(np.float32))
    out["pp.in"] = hpp.numpy()
    out["pp.out"] = layer(g, hpp).detach().numpy()
    out.update(pack("pp.state", layer.state_dict()))
    np.savez_compressed(os.path.join(HERE, "gcnsage_layers.npz"), **out)
    manifest.append({"file": "gcnsage_layers.npz", "kind": "GcnSAGELayer variants", "nodes": 61, "edges": 400})

    # 5. MeanSAGE (WeightedMeanSAGELayer x (n_layers + 1), relu + F.normalize between layers)
    g = multigraph_shim(9, n=83, e=600, f=13)
    torch.manual_seed(4)
    model = ref.MeanSAGE(13, 20, 9, 2)
    h = g.ndata["feat"].clone().requires_grad_(True)
    w = g.edata["feat"]
    y = model(g, h, w)
    up = torch.from_numpy(np.random.default_rng(13).standard_normal(tuple(y.shape)).astype(np.float32))
    (y * up).sum().backward()
    out = {"src": g.edges()[0].numpy().astype(np.int32), "dst": g.edges()[1].numpy().astype(np.int32),
           "num_nodes": np.int64(83), "weight": w.numpy(), "feat": g.ndata["feat"].numpy(),
           "out": y.detach().numpy(), "upstream": up.numpy(), "dh": h.grad.numpy(),
           "config": np.array([13, 20, 9, 2], dtype=np.int64)}
    out.update(pack("state", model.state_dict()))
    out.update(pack("grad", {k: p.grad for k, p in model.named_parameters()}))
    np.savez_compressed(os.path.join(HERE, "meansage.npz"), **out)
    manifest.append({"file": "meansage.npz", "kind": "MeanSAGE", "nodes": 83, "edges": 600})

    # 6. tensor-core route pin: the repo-default model on 8 pages of 300 nodes (N = 2400 >= 1024 rows, the size from
    #    which the B200 layers take the tcgen05 3xTF32 kernels), class weights on; reference-produced logits/gradients
    pages = [synth.make_page(1000 + i, n=300) for i in range(8)]
    cw = torch.tensor([1.0] * 6 + [2.0] + [1.0] * 2)
    manifest.append(run_gcnsage(ref, "gcnsage_default_2400", shim_batch(pages), 13, 218, 9, 3, seed=5, class_w=cw))

    # 7. same size, ragged pages (40..900 nodes) and the symmetrised k=5 variant: 4 layers, hidden 96
    sizes = (900, 40, 333, 128, 517, 77, 260, 145)
    pages = [synth.make_page(2000 + i, n=n, k=5, bidirectional=True) for i, n in enumerate(sizes)]
    manifest.append(run_gcnsage(ref, "gcnsage_ragged_2400", shim_batch(pages), 13, 96, 9, 4, seed=6))

    with open(os.path.join(HERE, "MANIFEST.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": REF_MODELS,
                   "dgl": "oracle/dgl_shim.py (DGL not installable; see oracle/__init__.py)",
                   "torch": torch.__version__, "fixtures": manifest}, fh, indent=1)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
