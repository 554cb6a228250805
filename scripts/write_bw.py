"""How fast can this GPU absorb plain writes?  (context for the store-bound projection epilogues)"""
import torch
dev = torch.device("cuda:0")
def t(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for mb in (134, 268, 1024):
    n = mb * 1000 * 1000 // 4
    x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
    ms = t(lambda: x.zero_())
    print(f"fill {mb} MB: {ms:.4f} ms  {mb / ms:.1f} GB/s")
    ms = t(lambda: y.copy_(x))
    print(f"copy {mb} MB: {ms:.4f} ms  {2 * mb / ms:.1f} GB/s (read+write)")
    ms = t(lambda: torch.cuda.memset if False else x.fill_(1.5))
    print(f"fill_(1.5) {mb} MB: {ms:.4f} ms  {mb / ms:.1f} GB/s")
rows = 153600
z = torch.empty(rows, 224, device=dev)
ms = t(lambda: z[:, :218].zero_())
print(f"strided fill [153600, 218 of 224]: {ms:.4f} ms {rows * 218 * 4 / ms / 1e6:.1f} GB/s")
z2 = torch.empty(rows, 224, device=dev)
def two():
    z.zero_(); z2.zero_()
ms = t(two)
print(f"two fills of 137.6 MB: {ms:.4f} ms {2 * rows * 224 * 4 / ms / 1e6:.1f} GB/s")
