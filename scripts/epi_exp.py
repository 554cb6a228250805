"""Event-timed epilogue experiments of the pair kernel at steady clocks (exp build: GTE_UMMA_DBG bits
2 = no output stores, 4 = no normalisation math, 8 = no statistics math)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tableextraction_b200 import ops, _lib
dev = torch.device("cuda:0")
n = 153600
def pad(r, c):
    t = ops.empty_padded(r, c, dev); t.normal_(); return t
def timeit(f, it=40):
    for _ in range(10): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
h, ah, dz = pad(n, 218), pad(n, 218), pad(n, 218)
W = torch.randn(218, 436, device=dev) * 0.05
b = torch.randn(218, device=dev); g = torch.rand(218, device=dev) + 0.5; be = torch.randn(218, device=dev)
pack = ops.umma_pack_weights(W, 218, 2)
xc = ops.comb_buffer(n, dev); xc[:, :13].normal_(); xc[:, 16:29].normal_()
pack0 = ops.umma_pack_weights(torch.randn(218, 26, device=dev) * 0.2, 13, 2)
# keep the outputs fixed (no allocator effects)
cases = {
    "fwd436": lambda: ops.umma_linear_fwd(h, ah, 218, pack, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True),
    "bwd_data": lambda: ops.umma_linear_bwd_data(dz, pack, 218, 2),
    "fwd_comb": lambda: ops.umma_linear_fwd_comb(xc, 13, pack0, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True),
}
# burn-in so that the clocks are up
x = torch.randn(8192, 8192, device=dev)
for _ in range(50): x @ x
torch.cuda.synchronize()
for dbg in (0, 2, 6, 14, 32, 34):
    os.environ["GTE_UMMA_DBG"] = str(dbg)
    print(dbg, {k: round(timeit(f), 4) for k, f in cases.items()}, flush=True)
