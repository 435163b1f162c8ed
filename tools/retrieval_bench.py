"""Config 4 (BASELINE.json): image-text retrieval ranking over an N x N pair matrix, rows sharded by image across the ranks,
ONE all-gather of the fp32 score slabs over NCCL, ranking on rank 0.

    python tools/retrieval_bench.py --n-img 64 --n-cap 2000                      # one GPU
    python -m torch.distributed.run --nproc-per-node R --master-addr 127.0.0.1 tools/retrieval_bench.py --n-img 128 --n-cap 2000 --check

Times the job on the device (events around shard scoring + all-gather, max over ranks) and prints scored pairs/s (a
"pair" = one (image, caption) through the BERT joint encoder at 22.89 GFLOP, image features amortised: BASELINE.md §3).
--check: every rank also scores the FULL matrix alone and asserts the sharded + gathered matrix is bit-identical.
--check-rows K (for the full 2000 x 2000 job, where --check would double the work): rank 0 re-scores K sampled image rows
alone and compares them bit for bit with the gathered matrix.  The all-gather is timed on its own (CUDA events, max over ranks)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from medical_vision_langauge_transformer_b200 import retrieval, synth
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
ap = argparse.ArgumentParser()
ap.add_argument("--n-img", type=int, default=64); ap.add_argument("--n-cap", type=int, default=2000); ap.add_argument("--len", type=int, default=80)
ap.add_argument("--pair-batch", type=int, default=2048); ap.add_argument("--check", action="store_true")
ap.add_argument("--check-rows", type=int, default=0); ap.add_argument("--out", default="")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
model = M.MVLBertForRetrieval(C.offline_config("retrieval", max_length=a.len)).eval()
synth.load_synth(model, 0, "stress")
model = model.cuda().set_precision("bf16")
imgs = synth.synth_images(a.n_img, 1, 0.02); caps = synth.synth_token_ids(a.n_cap, a.len, 1)
labels = torch.zeros(a.n_img, a.n_cap); labels[torch.arange(a.n_img), torch.arange(a.n_img) % a.n_cap] = 1
retrieval.rank_task(model, imgs[:2 * world], caps[:64], labels[:2 * world, :64], rank, world, 64)   # warm-up (NCCL, kernels)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
timing = {}
e0.record()
full, metrics = retrieval.rank_task(model, imgs, caps, labels, rank, world, a.pair_batch, timing=timing)
e1.record(); torch.cuda.synchronize()
ag_us = timing["allgather_events"][0].elapsed_time(timing["allgather_events"][1]) * 1e3 if "allgather_events" in timing else 0.0
ms = torch.tensor([e0.elapsed_time(e1), ag_us], device="cuda")
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ag_job_us = ms[1].item(); ms = ms[:1]
lo, hi = retrieval.shard_rows(a.n_img, rank, world)
ag_us = retrieval.time_all_gather(full[lo:hi].contiguous(), a.n_img, world)     # the collective alone, ranks aligned first
ok = None
rows_ok = None
if a.check_rows > 0:
    g = torch.Generator().manual_seed(7)
    rows = torch.randperm(a.n_img, generator=g)[:a.check_rows].sort().values
    if rank == 0:
        alone = retrieval.score_matrix(model, imgs[rows], caps, 0, 1, a.pair_batch)
        rows_ok = bool(torch.equal(alone, full[rows.cuda()]))
if a.check:
    alone = retrieval.score_matrix(model, imgs, caps, 0, 1, a.pair_batch)
    ok = bool(torch.equal(alone, full))
    flag = torch.tensor([int(ok)], device="cuda")
    if world > 1: dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
if rank == 0:
    pairs = a.n_img * a.n_cap
    print(json.dumps({"workload": f"retrieval rank {a.n_img}x{a.n_cap} L={a.len} (config 4 shape, bounded image count)", "n_gpus": world,
                      "ms": ms.item(), "pairs_per_s": pairs / ms.item() * 1e3, "bert_tflops": pairs * 22.89e9 / ms.item() / 1e9,
                      "sharded_equals_single_rank": ok, "sampled_rows_bit_identical": rows_ok, "sampled_rows": a.check_rows,
                      "allgather_us": ag_us, "allgather_in_job_us_incl_rank_skew": ag_job_us,
                      "allgather_bytes_per_rank": timing.get("allgather_bytes_per_rank", 0),
                      "R@1_i2t": metrics["i2t_retrieval"]["R@1"], "R@5_i2t": metrics["i2t_retrieval"]["R@5"],
                      "R@1_t2i": metrics["t2i_retrieval"]["R@1"], "pair_batch": a.pair_batch,
                      "pair_loop": "CUDA graph replay per pair batch"}))
if world > 1: dist.destroy_process_group()
