"""BASELINE.json configs[4]: conv-layer (aggregation kernel) microbench sweep.
hidden dim 32..512, average degree 5..40, 1M..50M edges per batch; block-diagonal batches of 300-node pages
(graph built directly on the device as a CSC with fixed in-degree), plus unstructured random graphs.
Reports achieved algorithmic HBM GB/s (8*N*F + 8*E + 4*N bytes per launch) against the measured
copy bandwidth (MEASURED_PEAKS.json) and the nominal 8 TB/s.  Timing: CUDA events, 3 warm-up + 10 timed launches,
operands larger than L2 except for the smallest cases (flagged)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tableextraction_b200 import _lib, ops

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
PAGE = 300
out = []
gen = torch.Generator(device=DEV).manual_seed(0)

def time_ms(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

E_TARGETS = [int(v) for v in os.environ.get("SWEEP_E", "1000000,5000000,20000000,50000000").split(",")]
DEGS = [int(v) for v in os.environ.get("SWEEP_DEG", "5,10,20,40").split(",")]
FS = [int(v) for v in os.environ.get("SWEEP_F", "32,64,128,218,256,512").split(",")]
for E_target in E_TARGETS:
    for deg in DEGS:
        n = max(PAGE, (E_target // deg) // PAGE * PAGE)
        e = n * deg
        pages = n // PAGE
        indptr = (torch.arange(n + 1, device=DEV, dtype=torch.int64) * deg).to(torch.int32)
        page_base = (torch.arange(n, device=DEV, dtype=torch.int32) // PAGE * PAGE).repeat_interleave(deg)
        idx_paged = (torch.randint(0, PAGE, (e,), device=DEV, generator=gen, dtype=torch.int32) + page_base).contiguous()
        idx_rand = torch.randint(0, n, (e,), device=DEV, generator=gen, dtype=torch.int32)
        del page_base
        w = torch.rand(e, device=DEV, generator=gen)
        norm = ops.degree_norm(indptr)
        page_off = torch.arange(pages + 1, device=DEV, dtype=torch.int32) * PAGE
        for f in FS:
            if n * f * 4 * 2 > 60e9:
                continue
            x = ops.empty_padded(n, f, DEV); x.normal_(generator=gen)
            y = ops.empty_padded(n, f, DEV)
            alg = 8 * n * f + 8 * e + 4 * n
            rec = {"E": e, "N": n, "deg": deg, "F": f, "alg_bytes": alg, "fits_l2": alg < 126e6}
            for name, idx, pg in (("paged", idx_paged, (page_off, pages, PAGE, PAGE * deg)), ("generic_blockdiag", idx_paged, None),
                                  ("generic_unstructured", idx_rand, None)):
                ms = time_ms(lambda: ops.spmm(indptr, idx, w, x, mode=_lib.GTE_AGG_SUM_NORM, row_norm=norm, out=y, pages=pg))
                rec[name] = {"ms": round(ms, 4), "GBs": round(alg / ms / 1e6, 1), "frac_measured": round(alg / ms / 1e6 / pk, 3),
                             "frac_8TBs": round(alg / ms / 1e6 / 8000, 3)}
            pg = (page_off, pages, PAGE, PAGE * deg)
            if ops.paged_packed_supported(pg, f):  # the persistent double-buffered kernel on pre-packed edges
                pke = ops.paged_pack_edges(indptr, idx_paged, w, pg)
                ms = time_ms(lambda: ops.spmm_packed(indptr, pke, x, pg, mode=_lib.GTE_AGG_SUM_NORM, row_norm=norm, out=y))
                rec["packed"] = {"ms": round(ms, 4), "GBs": round(alg / ms / 1e6, 1), "frac_measured": round(alg / ms / 1e6 / pk, 3),
                                 "frac_8TBs": round(alg / ms / 1e6 / 8000, 3)}
                rec["pack_ms"] = round(time_ms(lambda: ops.paged_pack_edges(indptr, idx_paged, w, pg)), 4)
                del pke
            out.append(rec)
            print(json.dumps(rec), flush=True)
            del x, y
        del indptr, idx_paged, idx_rand, w, norm
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"peak_hbm_gbs_measured": pk, "rows": out}, open(os.path.join(ROOT, "gpurun_out", os.environ.get("SWEEP_OUT", "conv_sweep.json")), "w"), indent=0)
