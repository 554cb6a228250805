"""Train / inference step engine for ``GcnSAGE`` on the sm_100a kernels.

Restates the train-step envelope of the reference
(/root/reference/src/models/model_train.py:297,320-332):

    batch_graph = dgl.batch(train_batch).to(device)      -> H2D of one pinned host batch
    logits = model(batch_graph)                           -> layer kernels (layers.py)
    loss = CrossEntropyLoss(weight)(logits, labels.long())-> gte_cross_entropy_fwd/bwd
    optimizer.zero_grad(); loss.backward(); optimizer.step()
                                                          -> explicit backward kernels writing
                                                             one flat gradient buffer, one
                                                             (optional) all-reduce, fused Adam

without autograd bookkeeping, without the per-step ``.item()`` sync, and -- for
fixed-shape batches -- replayed from a CUDA graph.  Data parallel by graph
(pages are independent): every rank runs the same step on its own pages; the
loss statistics ride in the tail of the flat gradient buffer, so ONE SUM
all-reduce per step carries gradients and statistics; Adam divides by the
GLOBAL label-weight sum, which reproduces the single-GPU result (not a mean
of per-rank means).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import layers as L
from . import ops
from ._lib import GteError
from .graph import PageGraphBatch, page_table
from .nn import GcnSAGE, _is_relu


def _dropout_p(d) -> float:
    """dropout attribute of the reference modules: nn.Dropout or the float 0. (models.py:30-33)"""
    return float(d.p) if isinstance(d, nn.Dropout) else float(d or 0.0)


class SageTrainer:
    def __init__(self, model: GcnSAGE, lr: float = 0.01, weight_decay: float = 5e-4, betas=(0.9, 0.999),
                 eps: float = 1e-8, class_weights: Optional[torch.Tensor] = None, process_group=None):
        self.model = model
        params = list(model.parameters())
        if not params or not params[0].is_cuda:
            raise GteError("SageTrainer: move the model to a CUDA device first (no CPU path)")
        self.device = params[0].device
        for layer in model.layers:
            if layer.activation is not None and not _is_relu(layer.activation):
                raise GteError("SageTrainer: only activation=F.relu/None is fused; use the nn.Module path otherwise")
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.class_w = None if class_weights is None else class_weights.to(self.device, torch.float32).contiguous()
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)

        # Dropout (models.py:30-33,60-61,113): masks come from Philox keyed by (seed, offset) held on the DEVICE, so a
        # captured step draws fresh masks on every replay: every call site of a step uses base offset + its own share,
        # and the base advances once per step (gte_rng_advance, captured with the step).
        self._drop_p = [_dropout_p(model.dropout)] + [_dropout_p(layer.dropout) for layer in model.layers]
        self.has_dropout = any(p > 0 for p in self._drop_p)
        gen = torch.cuda.default_generators[self.device.index if self.device.index is not None else torch.cuda.current_device()]
        rank = torch.distributed.get_rank(process_group) if self.world > 1 else 0
        # ranks share the seed (torch.manual_seed) but must not share masks: disjoint regions of the counter space
        self.rng_dev = torch.tensor([int(gen.initial_seed()) & (2 ** 63 - 1), rank << 44], dtype=torch.int64, device=self.device)
        # one flat fp32 buffer for parameters, gradients and both Adam moments;
        # the nn.Parameters become views, so state_dict()/load_state_dict() keep working
        total = sum(p.numel() for p in params)
        # every tensor starts on a 16-byte boundary so the kernels can vectorise
        offs, o = [], 0
        for p in params:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        # Data parallel: ONE all-reduce per step.  The loss statistics [sum w*nll, sum w, #correct] live in the tail of
        # the flat gradient buffer, every rank back-propagates the UN-normalised sum, the single SUM all-reduce carries
        # gradients and statistics together, and Adam divides by the global label-weight sum it finds in the tail
        # (gte_adam_step grad_den).  GTE_DP_FUSED=1 forces this path on one GPU (tests).
        self.dp_fused = self.world > 1 or os.environ.get("GTE_DP_FUSED", "0") == "1"
        self._flat_len = o
        self.flat_param = torch.zeros(o, dtype=torch.float32, device=self.device)
        self.flat_grad = None
        self._dp_peer = None
        if self.world > 1 and os.environ.get("GTE_DP_PEER", "1") != "0":
            self._setup_peer_exchange(o + 4)
        if self.flat_grad is None:
            self.flat_grad = torch.zeros(o + 4, dtype=torch.float32, device=self.device)
        self.exp_avg = torch.zeros(o, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(o, dtype=torch.float32, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.num_params = total
        self._grad_views: Dict[int, torch.Tensor] = {}
        with torch.no_grad():
            for p, off in zip(params, offs):
                view = self.flat_param[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                gv = self.flat_grad[off:off + p.numel()].view_as(p)
                p.grad = gv
                self._grad_views[id(p)] = gv
        self.stats = self.flat_grad[o:o + 3] if self.dp_fused else torch.zeros(3, dtype=torch.float32, device=self.device)
        # peer exchange: the tail of flat_grad keeps the LOCAL statistics, the fused kernel writes the global ones here
        self.stats_global = torch.zeros(4, dtype=torch.float32, device=self.device) if self._dp_peer else None
        self._one = torch.ones(1, dtype=torch.float32, device=self.device)
        if self.world > 1:
            # every rank must start from the same parameters / optimiser state (DDP broadcasts from rank 0 too):
            # do not rely on identical seeding
            for t in (self.flat_param, self.exp_avg, self.exp_avg_sq, self.step_dev):
                torch.distributed.broadcast(t, src=torch.distributed.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                                            group=self.pg)
        self._graph = None
        self._static: Optional[Dict[str, torch.Tensor]] = None

    # ----------------------------------------------------------- pieces ----
    def _layer_args(self, layer):
        has_ln = isinstance(layer.lynorm, nn.LayerNorm)
        return (layer.linear.weight.data, None if layer.linear.bias is None else layer.linear.bias.data,
                layer.lynorm.weight.data if has_ln else None, layer.lynorm.bias.data if has_ln else None, has_ln,
                layer.lynorm.eps if has_ln else 1e-5, _is_relu(layer.activation))

    def forward(self, g: PageGraphBatch, keep_ctx: bool = True, layer_outputs: Optional[list] = None,
                training: bool = False):
        """Layer loop of ``GcnSAGE.forward`` (models.py:105-116).  ``layer_outputs`` (tests): a list that receives every
        layer's output tensor (the parity tests read the ReLU on/off patterns from it).  ``training``: dropout active."""
        h = g.ndata["feat"]
        w_edge = g.edata["feat"]
        ctxs: List[L.LayerCtx] = []
        used = 0  # Philox counters consumed so far in this step
        drop_on = training and self.has_dropout
        if drop_on and self._drop_p[0] > 0:  # models.py:113: dropout on the input features (they need no gradient)
            h, _ = ops.dropout_concat(h, None, self._drop_p[0], rng_dev=self.rng_dev, offset=used)
            used += ops.dropout_counters(h.shape[0], h.shape[1])
        # tf32 hi / lo tiles of every layer's W in ONE launch (the layers would otherwise pack one by one)
        packs: Dict[int, torch.Tensor] = {}
        n_rows = h.shape[0]
        todo = [(li, layer.linear.weight.data, layer.linear.weight.shape[1] // 2)
                for li, layer in enumerate(self.model.layers)
                if L.wants_pack(n_rows, layer.linear.weight.shape[1] // 2, layer.linear.weight.shape[0], layer.use_pp)]
        if len(todo) > 1:
            for (li, _, _), pk in zip(todo, ops.umma_pack_weights_batch([(W, fin, 2) for _, W, fin in todo])):
                packs[li] = pk
        for li, layer in enumerate(self.model.layers):
            W, b, gamma, beta, has_ln, eps, relu = self._layer_args(layer)
            drop = None
            if drop_on and self._drop_p[1 + li] > 0:
                drop = L.DropoutSpec(self._drop_p[1 + li], 0, used, self.rng_dev)
                used += ops.dropout_counters(h.shape[0], W.shape[1])
            h, ctx = L.sage_layer_forward(g, h, w_edge, W, b, gamma, beta, ln=has_ln, relu=relu, eps=eps, agg=L.GCN,
                                          use_pp=layer.use_pp, save_for_backward=keep_ctx, dropout=drop,
                                          pack=packs.get(li))
            ctxs.append(ctx if keep_ctx else None)
            if layer_outputs is not None:
                layer_outputs.append(h)
        self._rng_used = used
        return h, ctxs

    def backward(self, g: PageGraphBatch, ctxs: List[L.LayerCtx], dlogits: torch.Tensor,
                 dy_comb: Optional[torch.Tensor] = None):
        dy = dlogits
        layers = list(self.model.layers)
        for i in range(len(layers) - 1, -1, -1):
            layer = layers[i]
            W, b, gamma, beta, has_ln, eps, relu = self._layer_args(layer)
            gv = self._grad_views
            dy = L.sage_layer_backward(
                g, ctxs[i], dy, W, gamma, beta, gv[id(layer.linear.weight)],
                None if layer.linear.bias is None else gv[id(layer.linear.bias)],
                gv[id(layer.lynorm.weight)] if has_ln else None, gv[id(layer.lynorm.bias)] if has_ln else None,
                need_dh=(i > 0), accumulate=False, dy_comb=dy_comb if i == len(layers) - 1 else None)

    def _all_reduce(self, t: torch.Tensor):
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=self.pg)

    # The step in kernel-only stages (forward + loss, backward, exchange + update): with the peer-memory exchange they
    # are ONE CUDA graph; with the NCCL fallback the first two are captured and the collective + Adam run eagerly.
    def _stage_forward(self, g: PageGraphBatch, labels: torch.Tensor):
        logits, ctxs = self.forward(g, training=self.model.training)
        ops.cross_entropy_fwd(logits, labels, self.class_w, stats=self.stats)
        return logits, ctxs

    def _stage_backward(self, g: PageGraphBatch, labels: torch.Tensor, logits, ctxs):
        den = self._one if self.dp_fused else self.stats[1:2]
        dc = None
        if L.wants_class_grad_comb(ctxs[-1], logits.shape[0]):
            # the class layer consumes [d logits | A_hat^T d logits] as one operand: the loss backward writes its half in place
            dc = ops.cross_entropy_bwd_comb(logits, labels, self.class_w, den)
            dlogits = ops.comb_views(dc, logits.shape[1])[0]
        else:
            dlogits = ops.cross_entropy_bwd(logits, labels, self.class_w, den)
        self.backward(g, ctxs, dlogits, dy_comb=dc)
        if getattr(self, "_rng_used", 0) > 0:
            ops.rng_advance(self.rng_dev, self._rng_used)  # the next step (or graph replay) draws new masks

    def _stage_update(self):
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, lr=self.lr, beta1=self.betas[0],
                      beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, step_dev=self.step_dev,
                      grad_den=self.stats[1:2] if self.dp_fused else None, count=self._flat_len)

    def _setup_peer_exchange(self, numel: int):
        """Flat gradient buffer in symmetric memory (every rank can load every other rank's buffer over NVLink) for the
        fused all-reduce + Adam kernel.  Falls back to NCCL + gte_adam_step when the ranks cannot map each other."""
        try:
            import torch.distributed._symmetric_memory as symm_mem

            group = self.pg if self.pg is not None else torch.distributed.group.WORLD
            buf = symm_mem.empty(numel, dtype=torch.float32, device=self.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group)
            if int(getattr(hdl, "offset", 0)) != 0 or hdl.world_size != self.world or hdl.signal_pad_size < 2048:
                raise RuntimeError("unexpected symmetric-memory layout")
            # this library's words of every rank's signal pad: 1 KB into the pad (torch's own barriers use its first words)
            pads = torch.tensor([int(p) + 1024 for p in hdl.signal_pad_ptrs], dtype=torch.int64, device=self.device)
            torch.cuda.synchronize(self.device)
            torch.distributed.barrier(group)
            self.flat_grad = buf
            self._dp_peer = {"hdl": hdl, "grad_ptrs": int(hdl.buffer_ptrs_dev), "pads": pads, "rank": int(hdl.rank),
                             "local": torch.zeros(4, dtype=torch.int32, device=self.device)}
        except Exception as exc:  # no P2P mapping (or an older torch): the NCCL path still works
            import warnings

            warnings.warn(f"SageTrainer: peer-memory gradient exchange unavailable ({type(exc).__name__}: {exc}); using NCCL")
            self.flat_grad, self._dp_peer = None, None

    def _stage_exchange_update(self):
        """gradient all-reduce + Adam: ONE kernel over NVLink peer memory when the ranks share symmetric memory, else one
        NCCL all-reduce (gradients + loss statistics in one buffer) + gte_adam_step"""
        if self._dp_peer is not None:
            d = self._dp_peer
            ops.dp_allreduce_adam(d["grad_ptrs"], d["pads"].data_ptr(), d["rank"], self.world, self._flat_len, self._flat_len,
                                  self.flat_param, self.exp_avg, self.exp_avg_sq, self.stats_global, lr=self.lr,
                                  beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.wd,
                                  step_dev=self.step_dev, local_words=d["local"])
        else:
            self._all_reduce(self.flat_grad)  # dp_fused: gradients + [sum w*nll, sum w, #correct] in one collective
            self._stage_update()

    def _result(self) -> torch.Tensor:
        return self.stats_global[:3] if self._dp_peer is not None else self.stats

    def step_stats(self) -> torch.Tensor:
        """device tensor [sum w*nll, sum w, #correct] of the last step (global over ranks)"""
        return self._result()

    def _step_impl(self, g: PageGraphBatch, labels: torch.Tensor):
        logits, ctxs = self._stage_forward(g, labels)
        self._stage_backward(g, labels, logits, ctxs)
        self._stage_exchange_update()
        return logits

    # ------------------------------------------------------------- API -----
    def train_step(self, g: PageGraphBatch, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step.  Returns the device tensor
        ``[sum w*nll, sum w, #correct]`` (global over ranks); loss = [0]/[1].
        No host synchronisation happens here."""
        if labels is None:
            labels = g.ndata["label"]
        with torch.cuda.device(self.device):
            self._step_impl(g, labels)
        return self._result()

    @torch.no_grad()
    def predict(self, g: PageGraphBatch) -> torch.Tensor:
        """Batched inference logits (model_predict.py:144-146 does this page by page)."""
        with torch.cuda.device(self.device):
            logits, _ = self.forward(g, keep_ctx=False)
        return logits

    def _predict_impl(self, g: PageGraphBatch, labels: Optional[torch.Tensor]):
        """forward (nothing saved for a backward pass) + argmax + per-page correct counts, all on the device"""
        pages = g.pages()
        if pages is None:
            raise GteError("captured predict pass needs a page table (batch_num_nodes / batch_num_edges)")
        logits, _ = self.forward(g, keep_ctx=False)
        return ops.page_predictions(logits, pages[0], pages[1], labels)

    @torch.no_grad()
    def predict_pages(self, g: PageGraphBatch, labels: Optional[torch.Tensor] = None):
        """The predict loop of model_predict.py:130-154 for a whole batch of pages in one pass:
        returns (preds int32 [N] on device, per-page accuracy float64 tensor [P] on device or None).
        ``all_pred.extend(preds.tolist())`` and ``mean_test_acc += acc`` of the reference become
        ``preds.tolist()`` and ``acc.sum()`` over the batches."""
        if labels is None:
            labels = g.ndata.get("label")
        pages = g.pages()
        if pages is None:  # one huge graph: a single "page"
            pages = page_table([g.num_nodes()], [g.num_edges()], g.num_nodes(), g.device)
            if pages is None:
                off = torch.tensor([0, g.num_nodes()], dtype=torch.int32, device=g.device)
                pages = (off, 1, g.num_nodes(), g.num_edges())
        with torch.cuda.device(self.device):
            logits, _ = self.forward(g, keep_ctx=False)
            preds, correct = ops.page_predictions(logits, pages[0], pages[1], labels)
        acc = None
        if correct is not None:
            sizes = (pages[0][1:] - pages[0][:-1]).to(torch.float64)
            acc = correct.to(torch.float64) / sizes
        return preds, acc

    # ----------------------------------------------------- CUDA graphs -----
    def capture_predict(self, host_batch: Dict[str, torch.Tensor]):
        """Capture the batched predict pass of model_predict.py:130-154 (format build + forward without saved
        activations + argmax + per-page correct counts) for batches of this shape.  Feed batches with ``load_batch`` /
        ``prefetch_batch``; ``replay()`` / ``replay_prefetched()`` then return ``(preds int32 [N], page_correct int32 [P])``
        -- static device tensors of the input set that ran, valid until that set runs again."""
        return self.capture(host_batch, split=False, mode="predict")

    def capture(self, host_batch: Dict[str, torch.Tensor], split: Optional[bool] = None, mode: str = "train"):
        """Capture the whole step (format build + forward + loss + backward +
        optimiser) for batches with exactly this node / edge count.  Later
        batches are fed with ``load_batch`` + ``replay``.  ``split`` (default: data-parallel
        runs) captures the kernels in one graph and leaves the all-reduce and Adam outside it."""
        if split is None:
            split = self.world > 1 and self._dp_peer is None  # the peer-memory exchange is an ordinary kernel: one graph
        if mode not in ("train", "predict"):
            raise GteError(f"capture: unknown mode {mode}")
        self._mode = mode
        dev = self.device
        n, e = int(host_batch["num_nodes"]), int(host_batch["src"].numel())
        f = int(host_batch["feat"].shape[1])
        st = {
            "src": torch.empty(e, dtype=torch.int32, device=dev),
            "dst": torch.empty(e, dtype=torch.int32, device=dev),
            "weight": torch.empty(e, dtype=torch.float32, device=dev),
            "feat": torch.empty((n, f), dtype=torch.float32, device=dev),
            "label": torch.empty(n, dtype=torch.float32, device=dev),
        }
        self._static = st
        bn, be = list(host_batch["batch_num_nodes"]), list(host_batch["batch_num_edges"])
        self._static_meta = (n, e, bn, be)
        # The page table (node / edge offsets of the pages) is part of the INPUT: a later batch with the same totals
        # may order or size its pages differently, so every static input set owns device copies of the two offset
        # arrays and load_batch / prefetch_batch refresh them.  Only the maxima (shared-memory sizing of the page
        # kernels) and the page count are fixed at capture time.
        pages = page_table(bn, be, n, dev)
        if pages is not None:
            st["page_off"], st["edge_off"] = pages[0], pages[4]
            self._static_caps = (pages[1], pages[2], pages[3])  # page count, max page nodes, max page edges
        else:
            self._static_caps = None
        self._loaded_layout = [None, None]
        self.load_batch(host_batch)

        self._split = split
        # warm-up on a side stream (allocator + lazy module state), restoring the optimiser state afterwards
        snap = (self.flat_param.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone())
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                if mode == "train":
                    self._step_impl(self._static_graph(st), st["label"])
                else:
                    self._predict_impl(self._static_graph(st), st["label"])
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self._check_static_structure(st)
        self._graph = self._capture_over(st)
        self._statics, self._graphs = [st], [self._graph]
        self._stage_sets = None
        self.flat_param.copy_(snap[0])
        self.exp_avg.copy_(snap[1])
        self.exp_avg_sq.copy_(snap[2])
        self.step_dev.copy_(snap[3])
        return self._graph

    def _static_graph(self, st) -> PageGraphBatch:
        n = self._static_meta[0]
        g = PageGraphBatch(st["src"], st["dst"], n, self._static_meta[2], self._static_meta[3])
        caps = self._static_caps
        g._cache["pages"] = None if caps is None else (st["page_off"], caps[0], caps[1], caps[2], st["edge_off"])
        g.edata["feat"] = st["weight"]
        g.ndata["feat"] = st["feat"]
        return g

    def _check_static_structure(self, st):
        """Host check (one synchronisation, capture time only) that the batch in ``st`` honours its page table: the
        one-kernel batch assembly raises a device flag when an edge leaves its page (results would be undefined)."""
        g = self._static_graph(st)
        if g.prepare(st["weight"]):
            g.check_page_structure()

    def _page_layout(self, host_batch):
        """Validated (page_off, edge_off) host int32 tensors of a batch fed to a captured step, or None when the
        captured step has no page table.  Raises when the batch cannot run through the captured kernels."""
        n, e, bn0, be0 = self._static_meta
        if int(host_batch["num_nodes"]) != n or int(host_batch["src"].numel()) != e:
            raise GteError("captured step: batch node / edge totals differ from the captured ones")
        if self._static_caps is None:
            return None
        bn, be = list(host_batch["batch_num_nodes"]), list(host_batch["batch_num_edges"])
        p, max_n, max_e = self._static_caps
        if len(bn) != p or len(be) != p or sum(bn) != n or sum(be) != e:
            raise GteError("captured step: page count / page sizes do not add up to the captured batch shape")
        if max(bn) > max_n or max(be) > max_e:
            raise GteError(f"captured step: largest page ({max(bn)} nodes, {max(be)} edges) exceeds the captured maxima "
                           f"({max_n}, {max_e}) the page kernels were sized for; capture with the largest batch first")
        if "page_off" in host_batch and "edge_off" in host_batch:
            return host_batch["page_off"], host_batch["edge_off"], (bn, be)
        import numpy as np
        po = np.zeros(p + 1, dtype=np.int32)
        eo = np.zeros(p + 1, dtype=np.int32)
        np.cumsum(np.asarray(bn, dtype=np.int64), out=po[1:])
        np.cumsum(np.asarray(be, dtype=np.int64), out=eo[1:])
        return torch.from_numpy(po), torch.from_numpy(eo), (bn, be)

    def _copy_inputs(self, st, which: int, host_batch, layout):
        for k in ("src", "dst", "weight", "feat", "label"):
            st[k].copy_(host_batch[k], non_blocking=True)
        if layout is not None:
            po, eo, key = layout
            if self._loaded_layout[which] != key:  # fixed-size pages: the offsets never change, skip the copies
                st["page_off"].copy_(po, non_blocking=True)
                st["edge_off"].copy_(eo, non_blocking=True)
                self._loaded_layout[which] = key

    def _capture_over(self, st, pool=None):
        """Capture the step reading the static input set ``st`` (capturing launches nothing).  Under data parallelism
        the collective stays outside: [graph: batch assembly + forward + CE + backward] -> ONE all-reduce (flat
        gradient + loss statistics) -> Adam (eager, 2 launches)."""
        graph = torch.cuda.CUDAGraph()
        kw = {} if pool is None else {"pool": pool}
        with torch.cuda.graph(graph, **kw):
            if getattr(self, "_mode", "train") == "predict":
                st["preds"], st["correct"] = self._predict_impl(self._static_graph(st), st["label"])
            elif self._split:
                g = self._static_graph(st)
                logits, ctxs = self._stage_forward(g, st["label"])
                self._stage_backward(g, st["label"], logits, ctxs)
            else:
                self._step_impl(self._static_graph(st), st["label"])
        return (graph,) if self._split else graph

    def load_batch(self, host_batch: Dict[str, torch.Tensor]):
        """Copy a host batch into static input set 0.  The batch must have the captured node / edge totals and page
        count and no page larger than the captured maxima; page ORDER and SIZES may differ (the page table is
        refreshed with the batch)."""
        st = self._static
        if st is None:
            raise GteError("load_batch: call capture() first")
        self._copy_inputs(st, 0, host_batch, self._page_layout(host_batch))

    def _replay_graphs(self, which: int = 0):
        graph = self._graphs[which] if getattr(self, "_graphs", None) else self._graph
        if isinstance(graph, tuple):
            graph[0].replay()
            self._stage_exchange_update()
        else:
            graph.replay()

    def _outputs(self, which: int):
        if getattr(self, "_mode", "train") == "predict":
            st = self._statics[which]
            return st["preds"], st["correct"]
        return self._result()

    def replay(self):
        self._replay_graphs()
        return self._outputs(0)

    # -- pipelined input: the host->device copy of batch i+1 overlaps the step of batch i ----------------------
    def prefetch_batch(self, host_batch: Dict[str, torch.Tensor]):
        """Start the H2D copy of the NEXT batch on a side stream.  Two static input sets alternate, each with its own
        captured graph (sharing one memory pool), so the step reads the batch where the copy engine put it: no
        device-to-device staging copy.  ``replay_prefetched()`` consumes the sets in order."""
        if self._static is None:
            raise GteError("prefetch_batch: call capture() first")
        layout = self._page_layout(host_batch)
        if getattr(self, "_stage_sets", None) is None:
            torch.cuda.synchronize(self.device)
            self._copy_stream = torch.cuda.Stream(device=self.device)
            inputs = {k: v for k, v in self._static.items() if k not in ("preds", "correct")}
            second = {k: torch.empty_like(v) for k, v in inputs.items()}
            for k, v in inputs.items():
                second[k].copy_(v)  # valid contents for the capture below (capturing runs nothing)
            self._loaded_layout[1] = self._loaded_layout[0]
            pool = (self._graph[0] if isinstance(self._graph, tuple) else self._graph).pool()
            self._statics = [self._static, second]
            self._graphs = [self._graph, self._capture_over(second, pool=pool)]
            self._stage_sets = self._statics
            self._stage_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_w = self._stage_r = 0
            # set 0 may still be read by work queued before this call
            self._stage_free[0].record(torch.cuda.current_stream(self.device))
            self._stage_free[1].record(torch.cuda.current_stream(self.device))
        if self._stage_w - self._stage_r >= 2:
            raise GteError("prefetch_batch: both input sets hold unconsumed batches; call replay_prefetched() first")
        i = self._stage_w % 2
        cs = self._copy_stream
        cs.wait_event(self._stage_free[i])  # the step that last read input set i has finished
        with torch.cuda.stream(cs):
            self._copy_inputs(self._statics[i], i, host_batch, layout)
            self._stage_ready[i].record(cs)
        self._stage_w += 1

    def replay_set(self, which: int) -> torch.Tensor:
        """Run one captured step on static input set ``which`` (0 or 1) as it is: inputs resident in HBM, no copy.
        The sets are filled by ``prefetch_batch`` / ``load_batch``; the caller keeps track of what is in them."""
        if getattr(self, "_stage_sets", None) is None and which != 0:
            raise GteError("replay_set: the second input set exists after the first prefetch_batch()")
        self._replay_graphs(which)
        return self._outputs(which)

    def replay_prefetched(self):
        """Run one captured step on the oldest prefetched batch."""
        if getattr(self, "_stage_sets", None) is None or self._stage_r >= self._stage_w:
            raise GteError("replay_prefetched: no prefetched batch")
        i = self._stage_r % 2
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._stage_ready[i])
        self._replay_graphs(i)
        self._stage_free[i].record(cur)
        self._stage_r += 1
        return self._outputs(i)
