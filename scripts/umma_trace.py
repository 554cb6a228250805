"""Role timeline of the tensor-core projection kernel (GTE_UMMA_DBG=1): where does a tile's time go?"""
import os, sys
os.environ["GTE_UMMA_DBG"] = os.environ.get("GTE_UMMA_DBG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from gnn_tableextraction_b200 import ops, lib
DEV = "cuda"
WHICH = int(os.environ.get("GTE_UMMA_PAIR", "1"))  # 1 = CTA-pair kernel, 0 = single-CTA kernel
ops.set_tuning(0, WHICH)
def run(tag, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print("   event ms", e0.elapsed_time(e1))
    slots = 8
    buf = np.zeros(148 * 16 * slots, dtype=np.int64)
    lib().gte_umma_debug_times(WHICH, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    t = buf.reshape(148, 16, slots)
    print("==", tag)
    span = (t[:, :, 3].max(axis=1) - t[:, 0, 7])
    print("   per-CTA span cycles: median", int(np.median(span)), "max", int(span.max()), " tile period (cta0):", [int(t[0, i + 1, 3] - t[0, i, 3]) for i in range(6)])
    for cta in (0,):
        base = t[cta, 0, 7]
        for tile in range(1, 5):
            r = t[cta, tile] - base
            print(f" cta {cta} tile {tile}: prod_first_tma {r[7]:7d} | mma wait_tempty {r[4]:7d}->{r[5]:7d} issued {r[6]:7d} | epi wait {r[0]:7d} tfull {r[1]:7d} stats {r[2]:7d} done {r[3]:7d}"
                  f"   [mma {r[6]-r[5]:6d}  epi_stats {r[2]-r[1]:6d} epi_store {r[3]-r[2]:6d}]"
                  + (f" [sums {r[8]-r[2]:6d} bar {r[9]-r[8]:6d} merge {r[10]-r[9]:6d} norm+store {r[3]-r[10]:6d}]" if slots == 16 and r[8] > 0 else ""))
n = 153600
h = ops.empty_padded(n, 218, DEV); h.normal_()
ah = ops.empty_padded(n, 218, DEV); ah.normal_()
W = torch.randn(218, 436, device=DEV) * 0.05
b = torch.randn(218, device=DEV); g = torch.ones(218, device=DEV); be = torch.zeros(218, device=DEV)
pack = ops.umma_pack_weights(W, 218, 2)
run("fwd 436->218 LN", lambda: ops.umma_linear_fwd(h, ah, 218, pack, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True))
run("bwd_data 218->436", lambda: ops.umma_linear_bwd_data(h, pack, 218, 2))
f = ops.empty_padded(n, 13, DEV); f.normal_()
a0 = ops.empty_padded(n, 13, DEV); a0.normal_()
W0 = torch.randn(218, 26, device=DEV) * 0.2
p0 = ops.umma_pack_weights(W0, 13, 2)
run("fwd 26->218 LN", lambda: ops.umma_linear_fwd(f, a0, 13, p0, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True))
xc = ops.comb_buffer(n, DEV)
xc[:, :13].normal_(); xc[:, 16:29].normal_()
run("fwd comb 32->218 LN", lambda: ops.umma_linear_fwd_comb(xc, 13, p0, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True))
