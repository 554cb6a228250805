"""Role accounting of the CTA-pair projection kernel (exp build: GTE_LIB=...libgte_b200_exp.so, GTE_UMMA_DBG=1)."""
import os, sys, ctypes
os.environ.setdefault("GTE_UMMA_DBG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_tableextraction_b200 import ops, lib
DEV = "cuda"
n = 153600
def pad(r, c):
    t = ops.empty_padded(r, c, DEV); t.normal_(); return t
h, ah, dz = pad(n, 218), pad(n, 218), pad(n, 218)
W = torch.randn(218, 436, device=DEV) * 0.05
b = torch.randn(218, device=DEV); g = torch.ones(218, device=DEV); be = torch.zeros(218, device=DEV)
pack = ops.umma_pack_weights(W, 218, 2)
xc = ops.comb_buffer(n, DEV); xc[:, :13].normal_(); xc[:, 16:29].normal_()
p0 = ops.umma_pack_weights(torch.randn(218, 26, device=DEV) * 0.2, 13, 2)
x = torch.randn(8192, 8192, device=DEV)
for _ in range(30): x @ x   # clocks up
def run(tag, fn):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    buf = np.zeros(148 * 16, dtype=np.int64)
    lib().gte_umma_debug_times(1, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    t = buf.reshape(148, 16)
    lead = t[0::2]   # leader CTAs (MMA issuer stats live there)
    med = np.median(t, axis=0).astype(np.int64); ml = np.median(lead, axis=0).astype(np.int64)
    kb = max(int(med[9]), 1)
    ms = e0.elapsed_time(e1) / 10
    print(f"== {tag}: {ms:.4f} ms; span {med[0]} cycles ({med[0] / ms / 1e6:.2f} GHz) over {kb} k-blocks = {med[0] // kb} / k-block")
    print(f"   producer waits on empty {med[1] // kb} | split waits on full {med[2] // kb} work {med[3] // kb} | "
          f"MMA (leader) waits on tempty {ml[4] // kb} on ready {ml[5] // kb} loop {ml[6] // kb} | epilogue waits on tfull {med[7] // kb} work {med[8] // kb}")
run("fwd 436->218 LN", lambda: ops.umma_linear_fwd(h, ah, 218, pack, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True))
run("bwd_data 218->436", lambda: ops.umma_linear_bwd_data(dz, pack, 218, 2))
run("fwd comb 32->218 LN", lambda: ops.umma_linear_fwd_comb(xc, 13, p0, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True))
