"""Role accounting of the tensor-core weight-gradient kernel (needs the -DGTE_EXPERIMENTS build: GTE_LIB=...libgte_b200_exp.so)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_tableextraction_b200 import ops, lib
DEV = "cuda"
def run(tag, n, fo, k1, k2):
    dz = ops.empty_padded(n, fo, DEV); dz.normal_()
    x1 = ops.empty_padded(n, k1, DEV); x1.normal_()
    x2 = ops.empty_padded(n, k2, DEV) if k2 else None
    if k2: x2.normal_()
    dW = torch.empty(fo, k1 + k2, device=DEV)
    for _ in range(3): ops.umma_linear_bwd_weight(dz, x1, x2, dW, None)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.umma_linear_bwd_weight(dz, x1, x2, dW, None)
    e1.record(); torch.cuda.synchronize()
    buf = np.zeros(148 * 16, dtype=np.int64)
    lib().gte_umma_debug_times(2, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    t = buf.reshape(148, 16)
    med = np.median(t, axis=0).astype(np.int64)
    st = max(int(med[8]), 1)
    print(f"== {tag}: {e0.elapsed_time(e1)/5:.4f} ms; per CTA (median): span {med[0]} cycles over {st} stages = {med[0]//st} / stage")
    print(f"   producer waits on empty {med[1]//st}/stage | split waits on full {med[2]//st} work {med[3]//st} | MMA waits on ready {med[4]//st} on tempty {med[5]//st} | epilogue waits on tfull {med[6]//st} work {med[7]//st}")
run("hidden dW 218 x (218+218)", 153600, 218, 218, 218)
run("input dW 218 x (13+13)", 153600, 218, 13, 13)
