"""Golden vectors for the BBOX node features, produced by the REFERENCE's own code:
`get_shape` and `get_histogram` are lifted verbatim (via ast, at generation time) out of
/root/reference/src/components/nlp/bbox.py (the module itself cannot be imported here: it needs dgl, attrdict,
pdf2image), then driven exactly like `Bbox.__call__` drives them (bbox.py:115-124) and cast `.float()` like
model_train.py:295.  Run in the build container:  python tests/golden/make_bbox_golden.py"""
import ast, os, random, string, textwrap
import numpy as np
import torch

REF = "/root/reference/src/components/nlp/bbox.py"
src = open(REF).read()
tree = ast.parse(src)
fns = {}
for node in ast.walk(tree):
    if isinstance(node, ast.FunctionDef) and node.name in ("get_shape", "get_histogram"):
        fns[node.name] = textwrap.dedent(ast.get_source_segment(src, node))
ns = {}
exec(fns["get_shape"], ns)
exec(fns["get_histogram"], ns)
get_shape, get_histogram = ns["get_shape"], ns["get_histogram"]

rng = random.Random(7)
alphabet = string.ascii_letters * 3 + string.digits * 2 + ".,;:-()%$@#/&" + " " * 4 + "àéñüß€"
boxes, texts = [], []
for i in range(4000):
    x0, y0 = rng.randint(0, 1600), rng.randint(0, 2100)
    w, h = rng.randint(0, 400), rng.randint(0, 60)          # includes degenerate 0-width boxes and odd sizes
    boxes.append([x0, y0, x0 + w, y0 + h])
    L = rng.choice([0, 1, 2, 3, 5, 7, 9, 12, 20, 33])
    texts.append("".join(rng.choice(alphabet) for _ in range(L)))
texts[0], texts[1], texts[2], texts[3] = "", "   ", "abc", "1/3"   # empty, blanks only, one class, thirds (sum != 1.0 path)
emb_shape = list(map(get_shape, boxes))
emb_hist = list(map(get_histogram, texts))
feat = torch.tensor(np.append(emb_shape, emb_hist, 1)).float().numpy()   # bbox.py:122 + model_train.py:295
counts = []
for t in texts:
    t = t.replace(" ", "")
    counts.append([sum(ch.isalpha() for ch in t), sum(ch.isdigit() for ch in t),
                   sum((not ch.isalpha()) and (not ch.isdigit()) for ch in t)])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bbox_features.npz")
np.savez_compressed(out, boxes=np.asarray(boxes, dtype=np.int32), counts=np.asarray(counts, dtype=np.int32), feat=feat,
                    texts=np.asarray(texts, dtype=object))
print("wrote", out, feat.shape, "rows with hist sum != 1 before fix:",
      sum(1 for c in counts if sum(c) and (c[0] / sum(c) + c[1] / sum(c) + c[2] / sum(c)) != 1.0))
