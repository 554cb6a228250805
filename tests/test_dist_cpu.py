"""world_size-2 gloo test (CPU): the data-parallel recipe of gnn_tableextraction_b200.parallel /
SageTrainer -- shard pages by graph, ONE SUM all-reduce of the un-normalised gradients together with the
loss statistics, division by the GLOBAL label-weight sum -- reproduces the single-process gradients.  The arithmetic runs on the
oracle here (no GPU); the same sequence runs on the kernels in SageTrainer._step_impl."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from gnn_tableextraction_b200 import synth
from gnn_tableextraction_b200.parallel import all_reduce_sum_, shard_pages_balanced, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import oracle_graph_from_pages
    from oracle import sage_oracle as so

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    pages = synth.make_pages(6, n=40, k=4, ragged=False)
    cw = torch.tensor([1.0] * 6 + [2.0] + [1.0] * 2)
    torch.manual_seed(0)
    model = so.OracleGcnSAGE(13, 16, 9, 3, F.relu, 0)
    b, e = shard_range(len(pages), rank, world)
    g = oracle_graph_from_pages(pages[b:e])
    labels = g.ndata["label"].long()
    logits = model(g)
    nll = F.cross_entropy(logits, labels, weight=cw, reduction="sum")
    stats = torch.tensor([nll.item(), cw[labels].sum().item(), float((logits.argmax(1) == labels).sum())])
    # the recipe of SageTrainer (dp_fused): back-propagate the UN-normalised local sum, ONE all-reduce of
    # [gradients | sum w*nll, sum w, #correct], then divide by the global label-weight sum (Adam's grad_den)
    nll.backward()
    buf = torch.cat([p.grad.reshape(-1) for p in model.parameters()] + [stats])
    all_reduce_sum_(buf)
    flat, stats = buf[:-3] / buf[-2], buf[-3:]
    if rank == 0:
        q.put((stats.numpy(), flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_by_graph_matches_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, flat = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process reference
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import oracle_graph_from_pages
    from oracle import sage_oracle as so

    pages = synth.make_pages(6, n=40, k=4, ragged=False)
    cw = torch.tensor([1.0] * 6 + [2.0] + [1.0] * 2)
    torch.manual_seed(0)
    model = so.OracleGcnSAGE(13, 16, 9, 3, F.relu, 0)
    g = oracle_graph_from_pages(pages)
    labels = g.ndata["label"].long()
    logits = model(g)
    loss = F.cross_entropy(logits, labels, weight=cw)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    assert abs(stats[0] / stats[1] - loss.item()) < 1e-6
    assert stats[2] == float((logits.argmax(1) == labels).sum())
    assert np.abs(flat - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


def test_shard_helpers():
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_range(3, 3, 4) == (3, 3)
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)
    sizes = [900, 40, 300, 310, 290, 60, 500, 120]
    parts = shard_pages_balanced(sizes, 3)
    assert sorted(i for p in parts for i in p) == list(range(8))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 300 and all(p == sorted(p) for p in parts)
