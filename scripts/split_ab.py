"""A/B: separate cross-term accumulator (one accumulator stage) vs a single accumulator (two stages, the epilogue of tile
i overlaps the MMAs of tile i + 1) in the pair kernel: speed and error vs fp64 at the config-2 hidden shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from gnn_tableextraction_b200 import ops, _lib
dev = torch.device("cuda:0")
n = 153600
gen = torch.Generator().manual_seed(3)
def pad(t):
    o = ops.empty_padded(t.shape[0], t.shape[1], dev); o.copy_(t); return o
def rel(a, b):
    a, b = a.double().cpu(), b.double()
    return ((a - b).abs().max() / b.abs().max()).item()
def timeit(f, it=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
h = torch.relu(torch.randn(n, 218, generator=gen)); ah = torch.relu(torch.randn(n, 218, generator=gen)) * 0.7
W = (torch.rand(218, 436, generator=gen) - 0.5) * (2 / 436 ** 0.5)
b = torch.randn(218, generator=gen) * 0.1
g, be = torch.rand(218, generator=gen) + 0.5, torch.randn(218, generator=gen) * 0.1
dz = torch.randn(n, 218, generator=gen)
ns = 20000  # fp64 reference on a slice
z64 = torch.cat([h[:ns], ah[:ns]], 1).double() @ W.double().t() + b.double()
y64 = F.relu(F.layer_norm(z64, (218,), g.double(), be.double(), 1e-5))
dx64 = dz[:ns].double() @ W.double()
hd, ahd, dzd = pad(h), pad(ah), pad(dz)
pack = ops.umma_pack_weights(W.to(dev), 218, 2)
for split in (1, 0):
    ops.set_tuning(_lib.GTE_TUNE_UMMA_SPLIT, split)
    f1 = lambda: ops.umma_linear_fwd(hd, ahd, 218, pack, b.to(dev), 218, gamma=g.to(dev), beta=be.to(dev), relu=True, fuse_ln=True)
    f2 = lambda: ops.umma_linear_bwd_data(dzd, pack, 218, 2)
    z, y, _, _ = f1()
    d1, d2 = f2()
    dx = torch.cat([d1[:ns, :218], d2[:ns, :218]], 1)
    zs = z[:ns, :218].double().cpu()
    bias = ((zs - z64) * z64.sign()).mean().item() / z64.abs().max().item()
    print(f"split={split}: fwd {timeit(f1):.4f} ms  bwd_data {timeit(f2):.4f} ms | z err {rel(z[:ns, :218], z64):.2e} (signed mean {bias:+.2e}) "
          f"y err {rel(y[:ns, :218], y64):.2e} dx err {rel(dx, dx64):.2e}")
zf = ops.linear_fwd(hd, ahd, W.to(dev), b.to(dev))
print(f"ffma z err {rel(zf[:ns, :218], z64):.2e}")
