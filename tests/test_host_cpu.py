"""CPU suite, part 2: host logic, the C ABI surface, and the world_size-2 sharding path (gloo)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_cabi_library_loads_and_exports_every_declared_symbol():
    from medical_vision_langauge_transformer_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mvlt_b200.h")).read()
    declared = set(re.findall(r"^int (mvlt_\w+)\(", header, flags=re.M))
    assert len(declared) >= 13
    lib = _lib.load()                                    # dlopen only: no CUDA context, no compute
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mvlt_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    sizes = set(re.findall(r"^long long (mvlt_\w+)\(", header, flags=re.M))
    assert sizes == set(_lib.SIZE_QUERIES) and all(hasattr(lib, n) for n in sizes)
    declared |= sizes
    assert lib.mvlt_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mvlt_\w+)", out))
    assert declared <= exported


def test_header_is_plain_c_and_links_from_a_c_host(tmp_path):
    """include/mvlt_b200.h is the binding a non-Python host compiles against: a C11 translation unit that takes the address
    of every declared entry point must compile with gcc (no C++, no CUDA headers) and link against libmvlt_b200.so; it runs
    without a GPU (only mvlt_abi_version is called)."""
    from medical_vision_langauge_transformer_b200 import _lib
    _lib.load()
    header = open(os.path.join(ROOT, "include", "mvlt_b200.h")).read()
    names = sorted(set(re.findall(r"^int (mvlt_\w+)\(", header, flags=re.M)))
    src = tmp_path / "host.c"
    src.write_text('#include <stdio.h>\n#include "mvlt_b200.h"\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t fns[] = {'
                   + ", ".join(f"(fn_t){n}" for n in names)
                   + '};\n  printf("%d %d\\n", (int)(sizeof fns / sizeof fns[0]), mvlt_abi_version());\n  return fns[0] == 0;\n}\n')
    exe = tmp_path / "host"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:libmvlt_b200.so", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == [str(len(names)), "1"], (out.stdout, out.stderr)


def test_ops_fail_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from medical_vision_langauge_transformer_b200 import _lib, ops
    with pytest.raises((_lib.MvltNativeError, AssertionError)):
        ops.softmax_rows(torch.zeros(2, 2))


@pytest.mark.parametrize("task", ["vqa", "retrieval", "pretrain", "caption"])
def test_state_dict_layout_matches_reference(task):
    from medical_vision_langauge_transformer_b200.modules import config as C, model as M
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))[task]
    cls = {"vqa": M.MVLBertForVQA, "retrieval": M.MVLBertForRetrieval, "pretrain": M.MVLBertForPretraining,
           "caption": M.MVLBertForImageCaption}[task]
    sd = cls(C.offline_config(task)).state_dict()
    assert list(sd) == list(ref)                         # same keys, same order
    assert all(list(sd[k].shape) == ref[k] for k in ref)


@pytest.mark.parametrize("conv", ["resnet101", "resnet50", "linear", "vit"])
def test_resnet_state_dict_layout_matches_reference(conv):
    """torchvision ResNet key layout under `conv.conv.0.` (incl. the never-applied fc and the BatchNorm buffers) + resnet_fc."""
    from medical_vision_langauge_transformer_b200.modules import config as C, model as M
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))[f"retrieval_{conv}"]
    sd = M.MVLBertForRetrieval(C.offline_config("retrieval", conv=conv)).state_dict()
    assert list(sd) == list(ref)
    assert all(list(sd[k].shape) == ref[k] for k in ref)


def test_resnet_bn_folding_matches_batchnorm_eval():
    """Host logic of the ResNet path: conv + eval BatchNorm == conv with folded weights + bias (packing is CPU-checkable)."""
    import torch.nn as nn
    import torch.nn.functional as F
    from medical_vision_langauge_transformer_b200.modules.visual_feature_extractor import _fold_bn
    torch.manual_seed(0)
    conv, bn = nn.Conv2d(8, 12, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(12).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5), bn.bias.normal_(), bn.running_mean.normal_(), bn.running_var.uniform_(0.5, 2.0)
        x = torch.randn(2, 8, 9, 9)
        w, b = _fold_bn(conv, bn)
        assert torch.allclose(F.conv2d(x, w, b, stride=2, padding=1), bn(conv(x)), atol=1e-5)


def test_config_classes_mirror_reference_defaults():
    from medical_vision_langauge_transformer_b200.modules import config as C
    v, p, r = C.MVLBertConfigforVQA(), C.MVLBertPretrainConfig(), C.MVLBertRetrieval()
    assert (v.type_vocab_size, v.result_num, v.hidden_dropout_prob, v.conv) == (3, 224, 0.1, "resnet101")
    assert (p.MLM_task, p.ITM_task, p.max_length) == (True, False, 150)
    assert (r.max_length, r.lr, r.hidden_dropout_prob, r.attention_probs_dropout_prob) == (80, 1e-6, 0.0, 0.1)
    assert (v.hidden_size, v.num_hidden_layers, v.intermediate_size, v.layer_norm_eps) == (768, 12, 3072, 1e-12)


def test_synth_is_deterministic_and_shaped():
    from medical_vision_langauge_transformer_b200 import synth
    a, b = synth.synth_token_ids(4, 80, 1), synth.synth_token_ids(4, 80, 1)
    assert torch.equal(a, b) and a.dtype == torch.int64
    for row in a:
        n = int((row > 0).sum())
        assert 10 <= n <= 80 and row[n - 1] == 104 and (row[n:] == 0).all() and (row[:n - 1] >= 1000).all()
    m, l = synth.synth_mlm_labels(a, 1)
    assert ((l != -100).sum(1) <= 10).all() and ((m == 103) == (l != -100)).all()
    assert torch.equal(synth.synth_tensor("x.weight", (4, 8), 0), synth.synth_tensor("x.weight", (4, 8), 0))
    assert not torch.equal(synth.synth_tensor("x.weight", (4, 8), 0), synth.synth_tensor("y.weight", (4, 8), 0))


def test_shard_rows_partitions_exactly():
    from medical_vision_langauge_transformer_b200.retrieval import shard_rows
    for n in (0, 1, 7, 8, 2000, 2001):
        for world in (1, 2, 3, 8):
            blocks = [shard_rows(n, r, world) for r in range(world)]
            covered = [i for lo, hi in blocks for i in range(lo, hi)]
            assert covered == list(range(n))
            assert max(hi - lo for lo, hi in blocks) <= -(-n // world) if n else True


def test_compute_ranks_matches_reference_semantics():
    from medical_vision_langauge_transformer_b200 import retrieval
    from oracle import mvlt_oracle as O
    g = torch.load(os.path.join(GOLDEN, "rank6.pt"))
    assert retrieval.compute_ranks(g["scores"], g["labels"]) == O.compute_ranks(g["scores"].numpy(), g["labels"].numpy())
    rng = np.random.default_rng(0)
    s = rng.random((9, 9)).astype(np.float32)
    s[3, 4] = s[3, 5]                                    # a tie: numpy argsort order must decide, as in the reference
    lab = np.eye(9, dtype=np.int64)
    lab[2] = 0                                           # a row with no positive -> rank N (run_retrieval.py:231)
    mine, ref = retrieval.compute_ranks(s, lab), O.compute_ranks(s, lab)
    assert mine == ref and mine[0][2] == 9
    ev = retrieval.evaluate(s, lab)
    assert ev["i2t_retrieval"]["R@10"] == sum(r < 10 for r in ref[0]) / 9


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from medical_vision_langauge_transformer_b200.retrieval import shard_rows, all_gather_scores
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
N, NC = 7, 5
full = torch.arange(N * NC, dtype=torch.float32).view(N, NC)
lo, hi = shard_rows(N, dist.get_rank(), 2)
out = all_gather_scores(full[lo:hi].clone(), N, 2)
assert torch.equal(out, full), out
dist.destroy_process_group()
print("ok")
"""


def test_row_sharded_all_gather_world2_gloo(tmp_path):
    """The N>1 path of the retrieval job on CPU: 2 processes, uneven shards (7 rows), one all-gather."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


# ---------------------------------------------------------------------------------------------------------------------
# HuggingFace load paths the reference's scripts use (run_vqa.py:256, run_retrieval.py:313, run_pretrain.py:190,:241) and
# whole-model pickles (run_vqa.py:114,:264) — construction only, no kernels
# ---------------------------------------------------------------------------------------------------------------------
def _task_model(task, seed=0):
    import torch
    from medical_vision_langauge_transformer_b200.modules import config as C, model as M
    torch.manual_seed(seed)
    cls = {"vqa": M.MVLBertForVQA, "retrieval": M.MVLBertForRetrieval, "pretrain": M.MVLBertForPretraining}[task]
    return cls(C.offline_config(task)).eval()


def test_fresh_construction_keeps_pytorch_default_init():
    """the reference never calls init_weights() (model.py:365): nn.Embedding ~ N(0,1), not N(0, 0.02)"""
    m = _task_model("retrieval")
    assert abs(m.MVLBert.word_embeddings.weight.std().item() - 1.0) < 0.02
    assert hasattr(m, "all_tied_weights_keys")          # post_init() ran


@pytest.mark.parametrize("task", ["vqa", "retrieval", "pretrain"])
def test_save_pretrained_from_pretrained_round_trip(tmp_path, task):
    import torch
    m = _task_model(task)
    m.save_pretrained(tmp_path)
    m2 = type(m).from_pretrained(tmp_path)
    sd1, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd1) == list(sd2)
    for k in sd1:
        assert torch.equal(sd1[k], sd2[k]), k


def test_pretraining_checkpoint_into_vqa_model(tmp_path):
    """run_vqa.py:253-256: keys the checkpoint lacks (final_mlp) are initialised as model.py:280-294 does, the rest loads"""
    import torch
    from medical_vision_langauge_transformer_b200.modules import config as C, model as M
    pm = _task_model("pretrain", seed=1)
    pm.save_pretrained(tmp_path)
    v = M.MVLBertForVQA.from_pretrained(tmp_path, config=C.offline_config("vqa"))
    sdp, sdv = pm.state_dict(), v.state_dict()
    shared = [k for k in sdv if k in sdp]
    assert len(shared) > 500 and all(torch.equal(sdp[k], sdv[k]) for k in shared)
    w, b = sdv["final_mlp.1.weight"], sdv["final_mlp.1.bias"]
    assert torch.isfinite(w).all() and 0.015 < w.std().item() < 0.025 and b.abs().max().item() == 0


def test_whole_model_pickle_round_trip(tmp_path):
    import torch
    m = _task_model("vqa")
    torch.save(m, tmp_path / "model.pt")
    m2 = torch.load(tmp_path / "model.pt", weights_only=False)
    assert type(m2) is type(m)
    sd1, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd1) == list(sd2) and all(torch.equal(sd1[k], sd2[k]) for k in sd1)


def test_train_mode_is_refused_not_silently_eval():
    import torch
    m = _task_model("vqa")
    m.train()
    with pytest.raises(NotImplementedError, match="eval"):
        m.conv.conv[0].forward_features(torch.zeros(1, 3, 224, 224))
    with pytest.raises(NotImplementedError, match="eval"):
        m.MVLBert.encode(torch.zeros(1, 80, dtype=torch.long), None, torch.zeros(1, 49, 768))


def test_vit_mlp_accepts_old_torchvision_keys():
    import torch
    from medical_vision_langauge_transformer_b200.modules.visual_feature_extractor import _ViTMLP
    mlp = _ViTMLP(8, 16)
    sd = {"linear_1.weight": torch.randn(16, 8), "linear_1.bias": torch.randn(16), "linear_2.weight": torch.randn(8, 16),
          "linear_2.bias": torch.randn(8)}
    mlp.load_state_dict(dict(sd), strict=True)
    assert torch.equal(mlp[0].weight, sd["linear_1.weight"]) and torch.equal(mlp[3].bias, sd["linear_2.bias"])


# ---------------------------------------------------------------------------------------------------------------------
# Drop-in at the import level: the reference's OWN scripts, imported unmodified with the three module
# aliases in place, must pick up this package's classes and build their models through their own code path
# ---------------------------------------------------------------------------------------------------------------------
_DROPIN_CHILD = r'''
import inspect, json, os, sys
root, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
mode = sys.argv[3]
out = {}
if mode == "reference":
    from oracle import ref_shims
    m, c, v = ref_shims.import_reference()
else:
    # INTEGRATION.md section 2, verbatim
    import medical_vision_langauge_transformer_b200.modules.model as m, medical_vision_langauge_transformer_b200.modules.config as c
    import medical_vision_langauge_transformer_b200.modules.visual_feature_extractor as v
    sys.modules["modules.model"], sys.modules["modules.config"], sys.modules["modules.visual_feature_extractor"] = m, c, v
    sys.path.insert(0, ref)
    os.chdir(ref)
    sys.argv = sys.argv[:1]
    import run_vqa, run_retrieval, run_pretrain                       # the reference's scripts, unmodified
    out["identity"] = {
        "run_vqa.MVLBertForVQA": run_vqa.MVLBertForVQA is m.MVLBertForVQA,
        "run_vqa.MVLBertConfigforVQA": run_vqa.MVLBertConfigforVQA is c.MVLBertConfigforVQA,
        "run_retrieval.MVLBertForRetrieval": run_retrieval.MVLBertForRetrieval is m.MVLBertForRetrieval,
        "run_retrieval.MVLBertRetrieval": run_retrieval.MVLBertRetrieval is c.MVLBertRetrieval,
        "run_pretrain.MVLBertForPretraining": run_pretrain.MVLBertForPretraining is m.MVLBertForPretraining,
    }
    # model construction as run_retrieval.RetrievalTask does it (run_retrieval.py:300-312), minus the network fetch of
    # bert-base-uncased (BertConfig defaults are bert-base) — the tokenizer is the reference's own vocab.txt
    from transformers import BertTokenizer
    config = run_retrieval.MVLBertRetrieval()
    config.conv = "swintransformer"
    tokenizer = BertTokenizer.from_pretrained("./dataset/bert-base-uncased")
    config.update_special_tokens(tokenizer)
    model = run_retrieval.MVLBertForRetrieval(config)
    out["built"] = {"class": type(model).__module__ + "." + type(model).__name__, "vocab_size": config.vocab_size,
                    "ids": [config.cls_token_id, config.sep_token_id, config.eos_token_id, config.mask_token_id],
                    "n_tensors": len(model.state_dict())}
    vcfg = run_vqa.MVLBertConfigforVQA(); vcfg.conv = "swintransformer"; vcfg.update_special_tokens(tokenizer); vcfg.result_num = 7
    vm = run_vqa.MVLBertForVQA(vcfg)
    out["built"]["vqa_head"] = list(vm.final_mlp[1].weight.shape)
    # the reference's own ranking (run_retrieval.py:218-249: dataset with img_num, flat per-pair results) against this package's
    # compute_ranks on the same N x N matrix, ties and a row without a positive included
    import numpy as np, torch
    from medical_vision_langauge_transformer_b200 import retrieval
    rng = np.random.RandomState(0)
    N = 17
    sim = np.round(rng.rand(N, N), 2)                 # two decimals: plenty of exact ties
    lab = np.eye(N); lab[3, 3] = 0; lab[5, 9] = 1
    class _DS:
        img_num = N
        def __len__(self): return N * N
    ref_ranks = run_retrieval.compute_ranks(_DS(), [list(sim.reshape(-1)), list(lab.reshape(-1))])
    mine = retrieval.compute_ranks_host(torch.from_numpy(sim), torch.from_numpy(lab))
    out["compute_ranks"] = {"reference": [list(map(int, r)) for r in ref_ranks], "ours": [list(map(int, r)) for r in mine]}
sig = {}
for cls, meths in (("MVLBert", ["forward"]), ("Conv_layer", ["forward"]), ("MVLBertForVQA", ["forward"]),
                   ("MVLBertForRetrieval", ["forward"]), ("MVLBertForPretraining", ["forward"])):
    for meth in meths:
        sig[f"{cls}.{meth}"] = [(n, repr(p.default) if p.default is not inspect._empty else None)
                                for n, p in inspect.signature(getattr(getattr(m, cls), meth)).parameters.items()]
sig["SwinTransformer.__init__"] = [n for n in inspect.signature(v.SwinTransformer.__init__).parameters][:17]
out["signatures"] = sig
print("JSON" + json.dumps(out))
'''


def _run_dropin_child(mode):
    ref = os.environ.get("MVLT_REFERENCE_ROOT", "/root/reference")
    r = subprocess.run([sys.executable, "-c", _DROPIN_CHILD, ROOT, ref, mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("JSON")][-1]
    return json.loads(line[4:])


def test_dropin_reference_scripts_import_and_build_with_module_aliases():
    """run_vqa.py / run_retrieval.py / run_pretrain.py of the reference, imported UNMODIFIED after the sys.modules aliases of
    INTEGRATION.md §2: they bind this package's classes, build the task models through their own construction code, and the
    forward signatures equal the reference's (parameter names and defaults)."""
    ref = os.environ.get("MVLT_REFERENCE_ROOT", "/root/reference")
    if not os.path.isfile(os.path.join(ref, "run_vqa.py")):
        pytest.skip("reference tree not present (GPU box)")
    ours, theirs = _run_dropin_child("ours"), _run_dropin_child("reference")
    assert all(ours["identity"].values()), ours["identity"]
    assert ours["built"]["class"] == "medical_vision_langauge_transformer_b200.modules.model.MVLBertForRetrieval"
    assert ours["built"]["vocab_size"] == 30522 and ours["built"]["ids"] == [101, 102, 104, 103]
    assert ours["built"]["n_tensors"] > 500 and ours["built"]["vqa_head"] == [7, 768]
    assert ours["compute_ranks"]["ours"] == ours["compute_ranks"]["reference"]
    for name, sig in theirs["signatures"].items():
        mine = ours["signatures"][name]
        if name == "SwinTransformer.__init__":
            assert mine == sig, (name, mine, sig)
        else:
            assert mine[:len(sig)] == sig, (name, mine, sig)      # same names / defaults; extra trailing kwargs are allowed


class _FakeTokenizer:
    eos_token = "[END]"

    def convert_tokens_to_ids(self, tokens):
        return [1000 + (sum(map(ord, t)) % 20000) for t in tokens]


class _FakeRetrievalDataset:
    """The fields of the reference's RetrievalDataset (run_retrieval.py:25-77) that the test split touches."""

    def __init__(self, n=7, max_caption_len=12):
        g = torch.Generator().manual_seed(4)
        self.img_num, self.max_caption_len, self.tokenizer, self.split = n, max_caption_len, _FakeTokenizer(), "test"
        self.images = torch.randn(n, 3, 8, 8, generator=g).numpy()
        lens = [3, 12, 20, 1, 7, 12, 5][:n]
        self.tokens = [[f"w{i}_{k}" for k in range(l)] + ["[END]"] for i, l in enumerate(lens)]
        self.cap_ids = [10, 11, 12, 10, 14, 15, 12][:n]              # duplicates: different images sharing a caption id

    def get_data_by_idx(self, i):
        return self.images[i], "caption", self.tokens[i], f"img{i}", self.cap_ids[i]

    def __len__(self):
        return self.img_num ** 2

    def __getitem__(self, index):                                   # restatement of run_retrieval.py:126-145 (test split)
        img_idx, cap_idx = index // self.img_num, index % self.img_num
        img1, _, _, _, cap_id1 = self.get_data_by_idx(img_idx)
        _, _, tok2, _, cap_id2 = self.get_data_by_idx(cap_idx)
        label = 1 if img_idx == cap_idx or cap_id1 == cap_id2 else 0
        cap = np.array(self.tokenizer.convert_tokens_to_ids(tok2), dtype=np.int64)
        new = np.zeros(self.max_caption_len, dtype=np.int64)
        new[:min(cap.shape[0], self.max_caption_len)] = cap[:min(cap.shape[0], self.max_caption_len)]
        return img1, new, label


def test_pretokenize_matches_the_reference_pair_enumeration():
    """retrieval.pretokenize reads every image and tokenises every caption ONCE; the N^2 (image, ids, label) triples the reference's
    test split would have produced one by one (run_retrieval.py:126-145) are exactly images[i], ids[j], labels[i, j]."""
    from medical_vision_langauge_transformer_b200 import retrieval
    ds = _FakeRetrievalDataset()
    images, ids, labels = retrieval.pretokenize(ds)
    n = ds.img_num
    assert images.shape == (n, 3, 8, 8) and ids.shape == (n, ds.max_caption_len) and ids.dtype == torch.int64 and labels.shape == (n, n)
    for index in range(len(ds)):
        img, cap, lab = ds[index]
        i, j = divmod(index, n)
        assert np.array_equal(images[i].numpy(), img) and np.array_equal(ids[j].numpy(), cap) and labels[i, j].item() == lab
    assert labels[0, 3] == 1 and labels[2, 6] == 1 and labels[0, 1] == 0      # shared cap_id pairs are positives
