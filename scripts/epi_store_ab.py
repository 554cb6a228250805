"""A/B of the pair kernel's epilogue store path (TMA stores vs the warps' own global stores) on the three projection
shapes of config 2.  Device times with CUDA events; inputs larger than L2 in total."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tableextraction_b200 import ops, _lib

dev = torch.device("cuda:0")
n = 153600
gen = torch.Generator(device=dev).manual_seed(1)


def pad(r, c):
    t = ops.empty_padded(r, c, dev)
    t.copy_(torch.randn(r, c, device=dev, generator=gen))
    return t


def timeit(f, it=30):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


h, ah = pad(n, 218), pad(n, 218)
W = torch.randn(218, 436, device=dev) * 0.05
b = torch.randn(218, device=dev)
g, be = torch.rand(218, device=dev) + 0.5, torch.randn(218, device=dev)
pack = ops.umma_pack_weights(W, 218, 2)
xc = ops.comb_buffer(n, dev)
xc[:, :13] = torch.randn(n, 13, device=dev)
xc[:, 16:29] = torch.randn(n, 13, device=dev)
W0 = torch.randn(218, 26, device=dev) * 0.2
pack0 = ops.umma_pack_weights(W0, 13, 2)
dz = pad(n, 218)
res = {}
for mode in (0, 1):
    ops.set_tuning(_lib.GTE_TUNE_EPI_STORE, mode)
    r = {}
    r["fwd436"] = timeit(lambda: ops.umma_linear_fwd(h, ah, 218, pack, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True))
    r["fwd436_y"] = timeit(lambda: ops.umma_linear_fwd(h, ah, 218, pack, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True, want_z=False))
    r["bwd_data"] = timeit(lambda: ops.umma_linear_bwd_data(dz, pack, 218, 2))
    r["fwd_comb"] = timeit(lambda: ops.umma_linear_fwd_comb(xc, 13, pack0, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True))
    r["stacked"] = timeit(lambda: ops.umma_linear_fwd_stacked(h, 218, ops.umma_pack_weights(torch.randn(9, 436, device=dev), 218, 2) if False else pack9, b[:9], 9)) if False else None
    res[mode] = r
    outs = ops.umma_linear_fwd(h, ah, 218, pack, b, 218, gamma=g, beta=be, relu=True, fuse_ln=True)
    res[("out", mode)] = [o.clone() for o in outs]
    res[("dx", mode)] = [o.clone() for o in ops.umma_linear_bwd_data(dz, pack, 218, 2)]
for k in res[0]:
    print(k, res[0][k], res[1][k])
for a, c in zip(res[("out", 0)], res[("out", 1)]):
    print("equal", torch.equal(a[:, :218] if a.dim() == 2 else a, c[:, :218] if c.dim() == 2 else c))
for a, c in zip(res[("dx", 0)], res[("dx", 1)]):
    print("dx equal", torch.equal(a, c))
