"""First-contact check of the CTA-pair tensor-core kernels (run under `timeout`: a protocol error would hang)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gnn_tableextraction_b200 import ops, _lib

DEV = "cuda"


def rel(a, b):
    return ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max()).item()


def padded(t):
    o = ops.empty_padded(t.shape[0], t.shape[1], DEV)
    o.copy_(t)
    return o


def fwd(n, fin, fo, ln, relu, two):
    gen = torch.Generator().manual_seed(n)
    x1 = torch.randn(n, fin, generator=gen)
    x2 = torch.randn(n, fin, generator=gen) * 3 if two else None
    nseg = 2 if two else 1
    W = (torch.rand(fo, nseg * fin, generator=gen) - 0.5) * (2.0 / (nseg * fin) ** 0.5)
    b = torch.randn(fo, generator=gen) * 0.1
    gamma = torch.rand(fo, generator=gen) + 0.5
    beta = torch.randn(fo, generator=gen) * 0.1
    X = torch.cat([x1, x2], 1) if two else x1
    z64 = X.double() @ W.double().t() + b.double()
    y64 = F.layer_norm(z64, (fo,), gamma.double(), beta.double(), 1e-5) if ln else z64
    if relu:
        y64 = F.relu(y64)
    pack = ops.umma_pack_weights(W.to(DEV), fin, nseg)
    out = {}
    for mode in (0, 1):
        ops.set_tuning(_lib.GTE_TUNE_UMMA_PAIR, mode)
        z, y, mean, rstd = ops.umma_linear_fwd(padded(x1), padded(x2) if two else None, fin, pack, b.to(DEV), fo,
                                               gamma=gamma.to(DEV), beta=beta.to(DEV), relu=relu, fuse_ln=ln, want_y=True)
        torch.cuda.synchronize()
        out[mode] = (rel(z, z64), rel(y, y64))
    print(f"fwd n={n} fin={fin} fo={fo} ln={ln} relu={relu} two={two}: single z/y {out[0][0]:.2e}/{out[0][1]:.2e}  "
          f"pair z/y {out[1][0]:.2e}/{out[1][1]:.2e}", flush=True)
    return max(out[1])


def bwd(n, fin, fo):
    gen = torch.Generator().manual_seed(n + 1)
    dz = torch.randn(n, fo, generator=gen)
    W = (torch.rand(fo, 2 * fin, generator=gen) - 0.5) * 0.2
    pack = ops.umma_pack_weights(W.to(DEV), fin, 2)
    res = {}
    for mode in (0, 1):
        ops.set_tuning(_lib.GTE_TUNE_UMMA_PAIR, mode)
        d1, d2 = ops.umma_linear_bwd_data(padded(dz), pack, fin, 2)
        torch.cuda.synchronize()
        res[mode] = (rel(d1, dz.double() @ W.double()[:, :fin]), rel(d2, dz.double() @ W.double()[:, fin:]))
    print(f"bwd n={n} fin={fin} fo={fo}: single {res[0][0]:.2e}/{res[0][1]:.2e} pair {res[1][0]:.2e}/{res[1][1]:.2e}", flush=True)
    return max(res[1])


worst = 0.0
for cfg in [(256, 32, 32, False, False, False), (1000, 218, 218, True, True, True), (129, 64, 64, False, False, True),
            (5000, 100, 130, True, False, True), (3001, 256, 256, True, True, True), (40000, 218, 218, True, True, True),
            (1500, 13, 218, True, True, True)]:
    worst = max(worst, fwd(*cfg))
for cfg in [(1000, 218, 218), (4097, 64, 96), (257, 256, 16)]:
    worst = max(worst, bwd(*cfg))
print("worst pair error", worst)
assert worst < 5e-6
print("PAIR SMOKE OK")
