"""Train-step time on RAGGED page batches (page sizes ~ clip(N(300, 80), 40, 900), SURVEY 8d robustness case):
eager steps, CUDA events; reports which aggregation kernels ran."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import gnn_tableextraction_b200 as gte
from gnn_tableextraction_b200 import synth
from gnn_tableextraction_b200.graph import batch_pages_host
P = int(os.environ.get("PAGES", "512"))
pages = synth.make_pages(P, ragged=True, k=10, distinct=128)
hb = batch_pages_host(pages)
torch.manual_seed(0)
model = gte.GcnSAGE(13, 218, 9, 3, F.relu, 0).cuda()
tr = gte.SageTrainer(model)
g = gte.PageGraphBatch.from_host(hb, "cuda")
pg = g.pages()
print("pages", P, "nodes", g.num_nodes(), "edges", g.num_edges(), "max page nodes/edges", pg[2], pg[3],
      "packed kernel usable at F=218:", gte.ops.paged_packed_supported(pg, 218), "one-kernel assembly:", gte.ops.page_formats_supported(pg))
for _ in range(3):
    tr.train_step(gte.PageGraphBatch.from_host(hb, "cuda"))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tr.train_step(gte.PageGraphBatch.from_host(hb, "cuda"))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"ragged eager step (H2D included): {ms:.3f} ms  -> {P / ms * 1e3:.0f} graphs/s, {g.num_nodes() / ms * 1e3 / 1e6:.1f} M nodes/s")
