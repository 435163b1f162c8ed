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

from typing import Iterator, Optional, Tuple

import numpy as np
import torch

from . import ops


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block of rank `rank`: equal ceil(n/world)-sized blocks, the tail ranks may be short or empty."""
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def device_chunks(host: torch.Tensor, chunk: int, device) -> Iterator[torch.Tensor]:
    """Yields `host[i:i+chunk]` on `device`, the copy of chunk i+1 overlapping the consumer's work on chunk i: two pinned staging
    buffers and two device buffers, H2D on a private copy stream, events both ways.  A device tensor is sliced as is.  Replaces the
    per-batch `img.cuda()` of pageable loader output (run_retrieval.py:200-201)."""
    n = host.shape[0]
    if host.is_cuda:
        for i in range(0, n, chunk):
            yield host[i:i + chunk]
        return
    dev = torch.device(device)
    copy, main = torch.cuda.Stream(device=dev), torch.cuda.current_stream(dev)
    shape = (min(chunk, n),) + tuple(host.shape[1:])
    pinned = [host.new_empty(shape).pin_memory() for _ in range(2)] if not host.is_pinned() else None
    dbuf = [torch.empty(shape, device=dev, dtype=host.dtype) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]

    def issue(k: int):
        lo, b = k * chunk, k % 2
        m = min(chunk, n - lo)
        with torch.cuda.stream(copy):
            if k >= 2:
                copy.wait_event(consumed[b])                 # the consumer has finished with the device buffer
            src = host[lo:lo + m]
            if pinned is not None:
                if k >= 2:
                    staged[b].synchronize()                  # the previous H2D out of this staging buffer has completed
                pinned[b][:m].copy_(src)                     # pageable -> pinned on the host
                src = pinned[b][:m]
            dbuf[b][:m].copy_(src, non_blocking=True)
            staged[b].record(copy)
            ready[b].record(copy)

    chunks = -(-n // chunk)
    if chunks:
        issue(0)
    for k in range(chunks):
        if k + 1 < chunks:
            issue(k + 1)
        b = k % 2
        main.wait_event(ready[b])
        yield dbuf[b][:min(chunk, n - k * chunk)]
        consumed[b].record(main)


@torch.no_grad()
def image_features(model, images: torch.Tensor, chunk: int = 64) -> torch.Tensor:
    """Conv_layer over `images` ([n,3,224,224], host or device) -> [n,49,768] in the activation dtype.  Host images stream through
    `device_chunks` (pinned, double-buffered: the copy of the next chunk overlaps the trunk of this one)."""
    dev = next(model.parameters()).device
    outs = [model.conv(x) for x in device_chunks(images, chunk, dev)]
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


def pretokenize(dataset) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The reference's test split (run_retrieval.py:126-145) enumerates N^2 (image i, caption j) pairs and converts caption j's tokens
    to ids inside `__getitem__`, i.e. N times per caption, after unpickling / indexing the image N times per image.  Here every
    image is read once and every caption tokenised once:
        images  float32 [N,3,224,224] (pinned when CUDA is available),
        caption ids int64 [N, max_caption_len] (truncated / zero-padded exactly as :141-144),
        labels  int64 [N,N]: 1 where i == j or the two entries share a cap_id (:137-139).
    `dataset` is duck-typed on the reference's RetrievalDataset: img_num, get_data_by_idx(i) -> (image, caption, tokens, img_id,
    cap_id), tokenizer.convert_tokens_to_ids, max_caption_len."""
    n, L = int(dataset.img_num), int(dataset.max_caption_len)
    ids = torch.zeros(n, L, dtype=torch.int64)
    images, cap_ids = None, []
    for i in range(n):
        im, _, tokens, _, cap_id = dataset.get_data_by_idx(i)
        im = torch.as_tensor(np.asarray(im), dtype=torch.float32)
        if images is None:
            images = torch.empty((n,) + tuple(im.shape), dtype=torch.float32, pin_memory=torch.cuda.is_available())
        images[i] = im
        t = dataset.tokenizer.convert_tokens_to_ids(tokens)[:L]
        ids[i, :len(t)] = torch.as_tensor(t, dtype=torch.int64)
        cap_ids.append(cap_id)
    uniq = {c: k for k, c in enumerate(dict.fromkeys(cap_ids))}
    code = torch.tensor([uniq[c] for c in cap_ids])
    labels = ((code[:, None] == code[None, :]) | torch.eye(n, dtype=torch.bool)).to(torch.int64)
    return images, ids, labels


def testRetrieval(model, test_data, output_file: Optional[str] = None, rank: int = 0, world: int = 1, pair_batch: int = 2048,
                  group=None):
    """Drop-in for run_retrieval.py:192-217: scores all N^2 pairs and returns (and, like the reference, torch.save's) `[results,
    labels]` — two dicts keyed by the flat pair index i * N + j holding Python floats / ints, the input format of the reference's
    `compute_ranks(dataset, results)` / `evaluate`.  `test_data` is the reference's test-split dataset (or a DataLoader around it);
    the N^2-element loader is never iterated: inputs go through `pretokenize` + `rank_task` (trunk once per image, pinned
    double-buffered H2D, row-sharded over `world` ranks, one all-gather) and the dicts are built from one D2H copy instead of one
    `.item()` per pair (:205-209).  Every rank returns the dicts; rank 0 writes the file."""
    dataset = getattr(test_data, "dataset", test_data)
    images, ids, labels = pretokenize(dataset)
    was_training = model.training
    model.eval()
    scores, _ = rank_task(model, images, ids, labels, rank, world, pair_batch, group)
    flat = scores.reshape(-1).cpu().tolist()
    results = dict(enumerate(flat))
    labs = dict(enumerate(labels.reshape(-1).tolist()))
    if output_file and rank == 0:
        torch.save([results, labs], output_file)
    model.train(was_training)
    return results, labs


def rank_task(model, images: torch.Tensor, captions: torch.Tensor, labels, rank: int = 0, world: int = 1,
              pair_batch: int = 512, group=None, timing: Optional[dict] = None) -> Tuple[torch.Tensor, Optional[dict]]:
    """Whole `--do_rank` job: shard, score, all-gather, rank on rank 0.  -> (scores [N,N], metrics or None)."""
    local = score_matrix(model, images, captions, rank, world, pair_batch)
    full = all_gather_scores(local, images.shape[0], world, group, timing)
    return full, (evaluate(full, labels) if rank == 0 else None)
