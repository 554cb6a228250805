"""Config-2 conv-layer launches (512 pages x 300 nodes, in-degree 10, F = 218) for ncu: forward (CSC, norm)
and backward (CSR order emulated by the same structure, with addend).  No timing here."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tableextraction_b200 import _lib, ops

DEV = "cuda"
PAGE, deg, pages, f = 300, 10, int(os.environ.get("PAGES", "512")), int(os.environ.get("F", "218"))
n, e = PAGE * pages, PAGE * pages * deg
gen = torch.Generator(device=DEV).manual_seed(0)
indptr = (torch.arange(n + 1, device=DEV, dtype=torch.int64) * deg).to(torch.int32)
base = (torch.arange(n, device=DEV, dtype=torch.int32) // PAGE * PAGE).repeat_interleave(deg)
idx = (torch.randint(0, PAGE, (e,), device=DEV, generator=gen, dtype=torch.int32) + base).contiguous()
w = torch.rand(e, device=DEV, generator=gen)
norm = ops.degree_norm(indptr)
page_off = torch.arange(pages + 1, device=DEV, dtype=torch.int32) * PAGE
pg = (page_off, pages, PAGE, PAGE * deg)
x = ops.empty_padded(n, f, DEV); x.normal_(generator=gen)
add = ops.empty_padded(n, f, DEV); add.normal_(generator=gen)
y = ops.empty_padded(n, f, DEV)
pk = ops.paged_pack_edges(indptr, idx, w, pg)
pk2 = ops.paged_pack_edges(indptr, idx, w, pg, pre_scale=norm)
for _ in range(int(os.environ.get("REPS", "2"))):
    ops.spmm_packed(indptr, pk, x, pg, mode=_lib.GTE_AGG_SUM_NORM, row_norm=norm, out=y)
    ops.spmm_packed(indptr, pk2, x, pg, mode=_lib.GTE_AGG_SUM, addend=add, out=y)
    if os.environ.get("OLD"):
        ops.spmm(indptr, idx, w, x, mode=_lib.GTE_AGG_SUM_NORM, row_norm=norm, out=y, pages=pg)
torch.cuda.synchronize()
print("done")
