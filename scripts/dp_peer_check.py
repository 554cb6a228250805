"""Data-parallel step with the fused peer-memory exchange (one kernel: one-shot all-reduce over NVLink + Adam) against
the NCCL path, under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_peer_check.py

Checks: same loss statistics and parameters as the NCCL path (fp32 rounding of the summation order), replicas
bit-identical across ranks, whole step replayed from ONE CUDA graph; prints the step times of both paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F
import gnn_tableextraction_b200 as gte
from gnn_tableextraction_b200 import synth
from gnn_tableextraction_b200.graph import batch_pages_host

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pages = int(os.environ.get("PAGES", "512"))
hb = batch_pages_host(synth.make_pages(pages, base_seed=42 + 1000 * rank, distinct=32), pin=True)


def make(peer):
    os.environ["GTE_DP_PEER"] = "1" if peer else "0"
    torch.manual_seed(0)
    m = gte.GcnSAGE(13, 218, 9, 3, F.relu, 0).to(dev)
    return gte.SageTrainer(m, lr=0.01, weight_decay=5e-4)


t_nccl, t_peer = make(False), make(True)
assert t_peer._dp_peer is not None, "peer-memory exchange did not come up"
res = {}
for name, tr in (("nccl", t_nccl), ("peer", t_peer)):
    tr.capture(hb)
    single_graph = not isinstance(tr._graph, tuple)
    for _ in range(3):
        tr.load_batch(hb)
        st = tr.replay().clone()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        tr.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[name] = (st, tr.flat_param.clone(), ms.item(), single_graph)
# the two trainers ran the same 3 + 50 steps on the same data
sa, pa, ma, ga = res["nccl"]
sb, pb, mb, gb = res["peer"]
err_p = ((pa - pb).abs().max() / pa.abs().max()).item()
err_s = ((sa - sb).abs().max() / sa.abs().max()).item()
# replicas identical across ranks?
chk = pb.double().sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
full = [torch.zeros_like(pb) for _ in range(world)] if rank == 0 else None
dist.gather(pb, full, dst=0)
if rank == 0:
    bit_identical = all(torch.equal(full[0], f) for f in full)
    print(f"world {world} pages/GPU {pages}: NCCL path {ma:.4f} ms/step (single graph: {ga}) | peer-memory fused exchange {mb:.4f} ms/step "
          f"(single graph: {gb}) | params rel diff {err_p:.2e} stats rel diff {err_s:.2e} | replicas bit-identical: {bit_identical and same}")
    assert err_p < 1e-4 and err_s < 1e-5 and bit_identical and same and gb
    print("DP PEER CHECK OK")
dist.destroy_process_group()
