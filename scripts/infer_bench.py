"""BASELINE.json configs[2]: inference on 100k synthetic page graphs sharded by graph across the GPUs of one box.

    python scripts/infer_bench.py [--pages 100000] [--batch 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/infer_bench.py

Every rank takes pages [r*P/W, (r+1)*P/W), streams them in batches of `--batch` pages from pinned host memory
(H2D inside the timed region), runs the batched predict pass (CSC build, 3 layers, argmax + per-page accuracy) and
keeps the predictions on the device; no collective on the data path, one all-reduce of the accuracy counters at
the end.  Prints one JSON line (rank 0): graphs/s = all pages / max-over-ranks device time."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import gnn_tableextraction_b200 as gte
from gnn_tableextraction_b200 import synth
from gnn_tableextraction_b200.graph import batch_pages_host

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=100_000)
ap.add_argument("--batch", type=int, default=4096)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
lo, hi = rank * args.pages // world, (rank + 1) * args.pages // world
mine = hi - lo
base = synth.make_pages(min(args.batch, mine), base_seed=42 + lo, n=300, k=10, distinct=64)
full = batch_pages_host(base, pin=True)
tail_n = mine % args.batch
tail = batch_pages_host(base[:tail_n], pin=True) if tail_n else None
torch.manual_seed(0)
model = gte.GcnSAGE(13, 218, 9, 3, F.relu, 0).to(dev).eval()
tr = gte.SageTrainer(model)

def run(host):
    g = gte.PageGraphBatch.from_host(host, dev)
    return tr.predict_pages(g)

for _ in range(2):
    run(full)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acc_sum = torch.zeros(1, dtype=torch.float64, device=dev)
npred = 0
e0.record()
for b in range(mine // args.batch):
    preds, acc = run(full)
    acc_sum += acc.sum()
    npred += preds.numel()
if tail is not None:
    preds, acc = run(tail)
    acc_sum += acc.sum()
    npred += preds.numel()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
cnt = torch.tensor([float(npred)], dtype=torch.float64, device=dev)
if world > 1:
    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    torch.distributed.all_reduce(acc_sum)
    torch.distributed.all_reduce(cnt)
if rank == 0:
    print(json.dumps({"metric": "page-graphs/sec (batched inference, H2D included)", "value": args.pages / (ms.item() / 1e3),
                      "unit": "graphs/s", "n_gpus": world, "pages": args.pages, "batch_pages": args.batch, "ms": ms.item(),
                      "nodes_predicted": int(cnt.item()), "mean_page_accuracy": acc_sum.item() / args.pages,
                      "config": "configs[2]: GcnSAGE 13-218-218-9 inference, 300-node pages, kNN k=10, sharded by graph"}))
if world > 1:
    torch.distributed.destroy_process_group()
