"""N x N image-text-matching scoring and ranking of `run_retrieval.py --do_test --do_rank`
(reference run_retrieval.py:126-145 pair enumeration, :192-217 scoring loop, :220-249 ranks, :283-295 R@K).

The reference pushes every (image i, caption j) pair through the whole model, Swin trunk included (N times per
image).  `Conv_layer` depends on the image only, so here the trunk runs once per image and the BERT joint encoder
runs once per pair, selecting the feature row through `img_index` inside the embedding kernel — same arithmetic per
pair, 161.5 -> 91.6 PFLOP at N = 2000.

Multi-GPU: the pair matrix is row-sharded by image (rank r scores rows [r*ceil(N/R), ...)), captions are replicated,
and ONE all-gather of the fp32 score slabs (torch.distributed / NCCL over NVLink) assembles [N, N] on every rank.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block of rank `rank`: equal ceil(n/world)-sized blocks, the tail ranks may be short or empty."""
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


@torch.no_grad()
def image_features(model, images: torch.Tensor, chunk: int = 64) -> torch.Tensor:
    """Conv_layer over `images` ([n,3,224,224], host or device) -> [n,49,768] in the activation dtype."""
    dev = next(model.parameters()).device
    outs = [model.conv(images[i:i + chunk].to(dev, non_blocking=True)) for i in range(0, images.shape[0], chunk)]
    return torch.cat(outs) if len(outs) > 1 else outs[0]


class _PairGraph:
    """One pair-batch of the scoring loop captured as a CUDA graph over static buffers: pair indices base .. base + PB
    (row-major enumeration of run_retrieval.py:133-145) -> img_index / caption gather -> joint encoder -> pooler -> head ->
    softmax -> prob[:, 1] scattered into the padded output; `base` advances by PB inside the graph, so the loop is
    `replay()` x n_batches with no host work in between (the eager loop spent arange / index_select / empty / slicing per batch)."""

    def __init__(self, model, feats: torch.Tensor, captions: torch.Tensor, pair_batch: int):
        dev = feats.device
        n_img, n_cap = feats.shape[0], captions.shape[0]
        self.P = n_img * n_cap
        self.PB = pair_batch
        self.n_batches = -(-self.P // pair_batch)
        self.out = torch.zeros(self.n_batches * pair_batch, device=dev, dtype=torch.float32)
        self.base = torch.zeros((), device=dev, dtype=torch.int64)
        ar = torch.arange(pair_batch, device=dev, dtype=torch.int64)
        bert = model.MVLBert

        def step():
            p = self.base + ar
            pc = p.clamp(max=self.P - 1)                       # the tail batch re-scores the last pair; its slots are padding
            img_index = torch.div(pc, n_cap, rounding_mode="floor").to(torch.int32)
            ids = captions.index_select(0, pc - img_index.to(torch.int64) * n_cap)
            hidden, shadow, B, S = bert.encode(ids, None, feats, None, False, img_index=img_index)
            prob = ops.softmax_rows(model.head_logits(bert.pool(hidden, shadow, B, S)))
            self.out.index_copy_(0, p, prob[:, 1].contiguous())
            self.base.add_(pair_batch)

        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            step()                                              # warm-up (packs weights, sizes the caching allocator)
            self.stream.synchronize()
            self.base.zero_()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                step()
            self.base.zero_()
        self.stream.synchronize()

    def run(self) -> torch.Tensor:
        with torch.cuda.stream(self.stream):
            for _ in range(self.n_batches):
                self.graph.replay()
        torch.cuda.current_stream(self.out.device).wait_stream(self.stream)
        return self.out[:self.P]


@torch.no_grad()
def score_pairs(model, feats: torch.Tensor, captions: torch.Tensor, pair_batch: int = 512, use_graph: bool = True) -> torch.Tensor:
    """prob[:,1] (run_retrieval.py:204) for every (feature row i, caption j), row-major -> fp32 [n_img, n_cap].
    use_graph: the pair-batch step is captured once as a CUDA graph and replayed (same kernels, same arithmetic)."""
    n_img, n_cap = feats.shape[0], captions.shape[0]
    dev = feats.device
    captions = captions.to(dev).contiguous()
    if use_graph and n_img * n_cap >= 2 * pair_batch:
        return _PairGraph(model, feats.contiguous(), captions, pair_batch).run().view(n_img, n_cap).clone()
    out = torch.empty(n_img * n_cap, device=dev, dtype=torch.float32)
    bert = model.MVLBert
    for p0 in range(0, n_img * n_cap, pair_batch):
        p = torch.arange(p0, min(p0 + pair_batch, n_img * n_cap), device=dev)
        img_index = (p // n_cap).to(torch.int32)
        ids = captions.index_select(0, p % n_cap)
        hidden, shadow, B, S = bert.encode(ids, None, feats, None, False, img_index=img_index)
        prob = ops.softmax_rows(model.head_logits(bert.pool(hidden, shadow, B, S)))
        out[p0:p0 + B] = prob[:, 1]
    return out.view(n_img, n_cap)


@torch.no_grad()
def score_matrix(model, images: torch.Tensor, captions: torch.Tensor, rank: int = 0, world: int = 1,
                 pair_batch: int = 512, image_chunk: int = 64) -> torch.Tensor:
    """This rank's slab of the score matrix: rows shard_rows(N, rank, world) -> fp32 [rows, n_cap] on the device."""
    lo, hi = shard_rows(images.shape[0], rank, world)
    dev = next(model.parameters()).device
    if hi <= lo:
        return torch.empty(0, captions.shape[0], device=dev, dtype=torch.float32)
    feats = image_features(model, images[lo:hi], image_chunk)
    return score_pairs(model, feats, captions, pair_batch)


def all_gather_scores(local: torch.Tensor, n_rows: int, world: int, group=None, timing: Optional[dict] = None) -> torch.Tensor:
    """The single collective of the path: equal-sized (padded) slabs -> [n_rows, n_cap] on every rank.
    timing (optional dict): receives 'allgather_events' = (start, end) CUDA events around the collective."""
    if world == 1:
        return local
    import torch.distributed as dist
    per = -(-n_rows // world)
    padded = local
    if local.shape[0] < per:
        padded = torch.zeros(per, local.shape[1], device=local.device, dtype=local.dtype)
        padded[:local.shape[0]] = local
    full = torch.empty(world * per, local.shape[1], device=local.device, dtype=local.dtype)
    padded = padded.contiguous()
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.all_gather_into_tensor(full, padded, group=group)
    if timing is not None:
        e1.record()
        timing["allgather_events"] = (e0, e1)
        timing["allgather_bytes_per_rank"] = padded.numel() * padded.element_size()
    return full[:n_rows]


def time_all_gather(local: torch.Tensor, n_rows: int, world: int, reps: int = 5, group=None) -> float:
    """Device time of the path's one collective on its own, in microseconds: the ranks are aligned with a barrier first (inside
    the job the interval around the all-gather also contains the wait for the slowest rank's scoring), then the all-gather of
    the real slabs is repeated `reps` times between two CUDA events; -> max over ranks of the mean."""
    if world == 1:
        return 0.0
    import torch.distributed as dist
    dist.barrier(group=group)
    torch.cuda.synchronize(local.device)
    all_gather_scores(local, n_rows, world, group)                # warm (buffers, NCCL channel set-up for this size)
    dist.barrier(group=group)
    torch.cuda.synchronize(local.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        all_gather_scores(local, n_rows, world, group)
    e1.record()
    torch.cuda.synchronize(local.device)
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], device=local.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.item()


def compute_ranks(scores, labels):
    """run_retrieval.py:220-249: per image row the position of the first matching caption in descending-score order, then
    per caption column. -> (i2t, t2i) as Python lists.  A CUDA score matrix (what `score_matrix` / `all_gather_scores`
    return) is ranked on the device (`mvlt_rank_first_positive`: two passes per line, no sort, no [N,N] D2H copy); host
    arrays take the reference's own numpy route."""
    if torch.is_tensor(scores) and scores.is_cuda:
        lab = labels if torch.is_tensor(labels) else torch.as_tensor(np.asarray(labels))
        rows, cols = ops.rank_first_positive(scores.float().contiguous(), lab)
        return rows.tolist(), cols.tolist()
    return compute_ranks_host(scores, labels)


def compute_ranks_host(scores, labels, kind=None):
    """The reference's numpy algorithm for host arrays (np.argsort reversed).  kind=None is the reference's call: numpy's default
    sort is NOT stable, so where other entries tie with the best positive's score the rank it reports is implementation-defined;
    kind="stable" fixes the order of equal scores (decreasing index after the reversal) — the rule the device kernel implements."""
    scores = np.asarray(scores.detach().cpu() if torch.is_tensor(scores) else scores)
    labels = np.asarray(labels.detach().cpu() if torch.is_tensor(labels) else labels)
    n = scores.shape[1]

    def ranks(sim, lab):
        out = []
        for s, l in zip(sim, lab):
            hit = np.nonzero(l[np.argsort(s, kind=kind)[::-1]] == 1)[0]
            out.append(int(hit[0]) if hit.size else n)
        return out

    return ranks(scores, labels), ranks(scores.T, labels.T)


def evaluate(scores, labels, ks=(1, 5, 10)) -> dict:
    """run_retrieval.py:283-295 -> {'i2t_retrieval': {'R@1':..}, 't2i_retrieval': {...}}"""
    i2t, t2i = compute_ranks(scores, labels)
    res = {"i2t_retrieval": {f"R@{k}": sum(r < k for r in i2t) / len(i2t) for k in ks}}
    if t2i:
        res["t2i_retrieval"] = {f"R@{k}": sum(r < k for r in t2i) / len(t2i) for k in ks}
    return res


def rank_task(model, images: torch.Tensor, captions: torch.Tensor, labels, rank: int = 0, world: int = 1,
              pair_batch: int = 512, group=None, timing: Optional[dict] = None) -> Tuple[torch.Tensor, Optional[dict]]:
    """Whole `--do_rank` job: shard, score, all-gather, rank on rank 0.  -> (scores [N,N], metrics or None)."""
    local = score_matrix(model, images, captions, rank, world, pair_batch)
    full = all_gather_scores(local, images.shape[0], world, group, timing)
    return full, (evaluate(full, labels) if rank == 0 else None)
