"""CPU suite, part 2: host-side logic — the C-ABI library loads and exports every symbol the header
declares, the drop-in modules keep the reference's names/shapes, the synthetic generator is
deterministic, the metric table and the world-size-2 sharding + all-reduce (gloo) agree with a single
process, and ops refuse to run without CUDA (no silent CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib, synthetic as syn
from few_shot_seg_cwt_b200.episodic import IoUTable


def _header_functions(name="cwt_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cwt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    names = _header_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cwt_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes binding table and header disagree"
    assert _lib.load().cwt_version() >= 100
    # the product header holds no developer entry points; those live in cwt_b200_debug.h (and are exported too)
    assert not [n for n in names if n.startswith("cwt_debug_")]
    dbg = _header_functions("cwt_b200_debug.h")
    assert set(dbg) == set(_lib.DEBUG_SIGNATURES) and all(n.startswith("cwt_debug_") for n in dbg)
    for n in dbg:
        assert hasattr(lib, n), f"{n} declared in include/cwt_b200_debug.h but not exported"


def test_library_is_blackwell_native():
    """sm_100a cubin only (no PTX-JIT for other arches, no multi-backend dispatch)."""
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    arches = set(re.findall(r"sm_\d+a?", out.stdout))
    assert arches == {"sm_100a"}, arches


def test_hot_kernels_use_the_blackwell_units_they_claim():
    """SASS of the shipped library: the resident fit keeps its tile in tensor memory (STTM / LDTM, no MMA), stages it with
    tensor-map TMA copies (UTMALDG) and reduces through 64-bit L2 atomics (REDG.E.ADD.64); the multi-shot fit does the same for
    the first tile of every CTA and streams the others (UBLKCP); the K projection runs on the 5th-gen tensor cores
    (UTCHMMA, operands by UTMALDG, accumulator read with LDTM); the streaming kernels are TMA-fed with packed fp32 FMAs."""
    lst = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if lst.returncode != 0:
        pytest.skip("cuobjdump not available")
    kernels, cur = {}, None
    for line in lst.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = set()
        elif cur is not None:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                kernels[cur].add(m.group(1).split(".")[0] if not m.group(1).startswith("REDG") else m.group(1))

    def ops_of(pattern):
        found = [v for k, v in kernels.items() if re.search(pattern, k)]
        assert found, pattern
        return found

    # tensor-memory resident fit: template arguments <512, 1, 512, 20, 5, 60, 60, PROF, TM = true, NA = 0>
    for ops in ops_of(r"k_fit_residentILi512ELi1ELi512ELi20ELi5ELi60ELi60ELb[01]ELb1ELi0E"):
        assert {"STTM", "LDTM", "UTMALDG", "FFMA2"} <= ops and any(o.startswith("REDG.E.ADD.64") for o in ops), sorted(ops)
        assert "UTCHMMA" not in ops                         # tensor memory as a scratchpad: no MMA in this kernel
    # multi-shot fit: first tile of every CTA in tensor memory, the others re-laid out with bulk stores and streamed with bulk /
    # tensor-map copies through the shared-memory ring
    for ops in ops_of(r"k_fit_l2ILi[1-4]E"):
        assert {"STTM", "LDTM", "UTMALDG", "FFMA2"} <= ops and any(o.startswith("REDG.E.ADD.64") for o in ops), sorted(ops)
    for ops in ops_of(r"k_fit_l2ILi[2-4]E"):                   # (one tile per CTA: nothing is streamed after the first sweep)
        assert "UBLKCP" in ops, sorted(ops)
    for ops in ops_of(r"k_kproj_scores"):
        assert {"UTCHMMA", "UTMALDG", "LDTM"} <= ops, sorted(ops)
    for name in ("k_logits_iou_stream", "k_rtf_stream", "k_ftr_stream"):
        for ops in ops_of(name):
            assert {"UTMALDG", "FFMA2"} <= ops, (name, sorted(ops))


def test_workspace_size_queries_need_no_gpu():
    lib = _lib.load()
    # at least the packed label cells (16 B per low-res cell) and the per-step buffers of the streaming algorithm
    assert lib.cwt_fit_workspace_bytes(4, 1, 512, 60, 60, 473, 473) > 4 * 60 * 60 * (16 + 8)
    # ... and it grows with the batch (per-episode accumulator words of the on-chip all-reduce)
    assert lib.cwt_fit_workspace_bytes(64, 1, 512, 60, 60, 473, 473) > lib.cwt_fit_workspace_bytes(4, 1, 512, 60, 60, 473, 473)
    assert lib.cwt_transformer_workspace_bytes(4, 2, 4, 512, 3600, 0) > 0
    assert lib.cwt_logits_iou_workspace_bytes(4, 2, 512, 60, 60, 473, 473) > 0


def test_no_cpu_fallback():
    f = torch.zeros(1, 1, 8, 2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cwt.fit_classifier(f, torch.zeros(1, 1, 9, 9, dtype=torch.uint8), torch.zeros(1, 2, 8), 0.1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cwt.batch_intersectionAndUnionGPU(torch.zeros(1, 1, 2, 2, 2), torch.zeros(1, 1, 9, 9, dtype=torch.long), 2)
    m = cwt.MultiHeadAttentionOne(1, 8, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        k = torch.zeros(1, 8, 2, 2)
        m.eval()(torch.zeros(1, 2, 8), k, k)


def test_module_state_dict_matches_reference_names_and_shapes():
    m = cwt.MultiHeadAttentionOne(4, 512, 512, 512, dropout=0.5)
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd == {"w_qkvs.weight": (2048, 512), "layer_norm.weight": (512,), "layer_norm.bias": (512,),
                  "fc.weight": (512, 2048), "fc.bias": (512,)}
    assert len(list(m.parameters())) == 5
    m.load_state_dict(syn.make_transformer_params(4, 512))
    m.train(); assert m.training
    m.eval(); assert not m.training
    # init statistics of src/model/transformer.py:44-51
    assert abs(float(m.w_qkvs.weight.std()) - np.sqrt(2.0 / 1024)) < 2e-3
    with pytest.raises(NotImplementedError):
        cwt.MultiHeadAttentionOne(1, 512, 256, 256)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present")
def test_module_loads_reference_checkpoint_format():
    sys.path.insert(0, "/root/reference")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.transformer import MultiHeadAttentionOne as RefMHA
    ref = RefMHA(4, 512, 512, 512, dropout=0.5)
    ckpt = {"epoch": 1, "state_dict": ref.state_dict()}          # src/train.py:138-163
    m = cwt.MultiHeadAttentionOne(4, 512, 512, 512, dropout=0.5)
    m.load_state_dict(ckpt["state_dict"])
    assert torch.equal(m.fc.weight, ref.fc.weight)


def test_checkpoint_format_is_the_references(tmp_path):
    """save_transformer_checkpoint writes {'epoch', 'state_dict', 'optimizer'} (src/train.py:138-163) and the reference's own
    MultiHeadAttentionOne loads it; a checkpoint written from the reference module loads into ours (src/test.py:82-89);
    the directory layout is src/util.py:167-179."""
    from oracle import ref_episode as R
    mods = R.load_reference_modules(prefer_live=False) or R.load_reference_modules(prefer_live=True)
    if mods is None:
        pytest.skip("reference modules not available (oracle/_ref not made)")
    RefMHA = mods[0]

    class A:
        model_dir, train_name, train_split, shot, arch, layers = str(tmp_path), "pascal", 0, 1, "resnet", 50
        main_optim, momentum, weight_decay, nesterov = "SGD", 0.9, 1e-4, True
        heads, bottleneck_dim = 4, 64
    assert cwt.get_model_dir_trans(A).endswith(os.path.join("pascal", "split=0", "model", "shot_1", "transformer_resnet50"))
    mine = cwt.MultiHeadAttentionOne(4, 64, 64, 64, dropout=0.5)
    opt = cwt.get_optimizer(A, [dict(params=mine.parameters(), lr=0.0025)])
    assert isinstance(opt, torch.optim.SGD) and opt.defaults["nesterov"] and opt.defaults["momentum"] == 0.9
    path = os.path.join(cwt.get_model_dir_trans(A), "best.pth")
    cwt.save_transformer_checkpoint(path, 3, mine, opt)
    ck = torch.load(path)
    assert set(ck) == {"epoch", "state_dict", "optimizer"} and ck["epoch"] == 3
    ref = RefMHA(4, 64, 64, 64, dropout=0.5)
    ref.load_state_dict(ck["state_dict"])                                   # strict: same names and shapes
    assert all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), mine.state_dict().values()))
    # the other direction: written from the reference's module and optimizer
    ref2 = RefMHA(4, 64, 64, 64, dropout=0.5)
    ropt = torch.optim.SGD(ref2.parameters(), lr=0.0025, momentum=0.9, weight_decay=1e-4, nesterov=True)
    torch.save({"epoch": 7, "state_dict": ref2.state_dict(), "optimizer": ropt.state_dict()}, str(tmp_path / "ref.pth"))
    mine2 = cwt.MultiHeadAttentionOne(4, 64, 64, 64, dropout=0.5)
    assert cwt.load_transformer_checkpoint(str(tmp_path / "ref.pth"), mine2, opt) == 7
    assert torch.equal(mine2.fc.weight, ref2.fc.weight) and torch.equal(mine2.w_qkvs.weight, ref2.w_qkvs.weight)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present")
def test_coscls_mirror_matches_reference_module():
    """few_shot_seg_cwt_b200.CosCls vs the reference's CosCls (src/model/pspnet.py:290-315): same state-dict keys and the
    same forward for every flag combination that does not re-parametrise the weight."""
    sys.path.insert(0, "/root/reference")
    import contextlib, io, warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.pspnet import CosCls as RefCosCls
    x = torch.relu(torch.randn(2, 16, 5, 6, generator=torch.Generator().manual_seed(3)))
    for cls_type in ("oooo", "0000", "00b0", "000t", "0n00", "0nbt"):
        with contextlib.redirect_stdout(io.StringIO()):
            ref = RefCosCls(in_dim=16, n_classes=2, cls_type=cls_type)
        mine = cwt.CosCls(16, 2, cls_type)
        assert set(mine.state_dict().keys()) == set(ref.state_dict().keys())
        mine.load_state_dict(ref.state_dict())
        assert torch.equal(mine(x), ref(x)), cls_type
        assert mine.plain == (cls_type in ("oooo", "0000"))
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefCosCls(in_dim=16, n_classes=2, cls_type="r000")            # WeightNorm: weight_g / weight_v
    assert set(cwt.CosCls(16, 2, "r000").state_dict().keys()) == set(ref.state_dict().keys())


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU and prints one JSON line with the
    keys of the contract."""
    import json, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-episodes", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "episodes/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["metric"] == "episodes_per_sec_head_pascal_1shot_60x60x512" and "workload" in d["config"]


def test_synthetic_generator_deterministic_and_well_formed():
    a = syn.make_episode(5, shot=2, C=32, h=12, w=12, H=89, W=89)
    b = syn.make_episode(5, shot=2, C=32, h=12, w=12, H=89, W=89)
    assert torch.equal(a.f_s, b.f_s) and torch.equal(a.q_label, b.q_label) and torch.equal(a.w0, b.w0)
    assert a.f_s.shape == (2, 32, 12, 12) and a.s_label.shape == (2, 89, 89) and a.s_label.dtype == torch.uint8
    assert set(torch.unique(a.s_label).tolist()) <= {0, 1, 255}
    assert (a.s_label == 1).sum() > 0 and (a.q_label == 1).sum() > 0
    assert float(a.f_s.min()) >= 0.0 and 0.2 < float((a.f_s == 0).float().mean()) < 0.8
    assert float(a.w0.abs().max()) <= 1 / np.sqrt(32) + 1e-7
    assert a.subcls == 5 % 5 + 1
    assert syn.shard_indices(10, 1, 4) == [1, 5, 9]
    batch = syn.make_batch([0, 1, 2], shot=1, C=32, h=12, w=12, H=89, W=89)
    assert batch.f_s.shape == (3, 1, 32, 12, 12) and batch.n_episodes == 3


def _fake_counts(idx):
    g = torch.Generator().manual_seed(idx)
    I = torch.randint(0, 1000, (2, 2), generator=g)
    U = I + torch.randint(1, 1000, (2, 2), generator=g)
    T = I + torch.randint(0, 500, (2, 2), generator=g)
    return torch.stack([I, U, T], -1)          # [V=2, class=2, 3]


def _table_for(indices, num_classes=5):
    t = IoUTable(num_classes, "cpu")
    if indices:
        counts = torch.stack([_fake_counts(i) for i in indices])
        sub = torch.tensor([i % num_classes + 1 for i in indices])
        t.update(counts, sub, torch.ones(len(indices), 2, 2, dtype=torch.float64))
    return t


def test_iou_table_matches_reference_accumulation():
    idx = list(range(23))
    t = _table_for(idx)
    # reference arithmetic: dict of per-class sums, IoU = I/(U+1e-10), mean over classes seen
    cI, cU = {}, {}
    for i in idx:
        c = i % 5 + 1
        k = _fake_counts(i)
        cI[c] = cI.get(c, 0) + int(k[0, 1, 0]); cU[c] = cU.get(c, 0) + int(k[0, 1, 1])
    ref = np.mean([cI[c] / (cU[c] + 1e-10) for c in cU])
    assert abs(t.miou(0) - ref) < 1e-12
    assert 0.0 < t.fb_iou(0) < 1.0 and int(t.n_episodes) == 23


def _worker(rank, world, port, n, q, n_steps=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = syn.shard_indices(n, rank, world)
    t = IoUTable(5, "cpu")
    per = (len(mine) + n_steps - 1) // n_steps
    for s_ in range(n_steps):                       # every rank reduces after every step (bench.py / HostPipeline pattern)
        part = mine[s_ * per:(s_ + 1) * per]
        if part:
            counts = torch.stack([_fake_counts(i) for i in part])
            t.update(counts, torch.tensor([i % 5 + 1 for i in part]), torch.ones(len(part), 2, 2, dtype=torch.float64))
        t.all_reduce()
    t.all_reduce()                                  # idempotent: a further reduce with nothing new changes nothing
    # plain lists: torch tensors in an mp.Queue travel through shared-memory files that vanish when the child exits
    q.put((rank, t.cls.tolist(), t.fb.tolist(), int(t.n_episodes), t.miou(0), t.miou(1), t.ce.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_steps", [1, 3])
def test_sharded_sweep_all_reduce_world2_gloo(n_steps):
    """episodes i -> rank i mod 2, integer all-reduce after every step (and once more at the end): identical table on both
    ranks, equal to the single-process table — however many times the table is reduced."""
    n, world = 17, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_steps
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q, n_steps)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in procs]
    single = _table_for(list(range(n)))
    for rank, cls, fb, ne, m0, m1, ce in res:
        assert cls == single.cls.tolist() and fb == single.fb.tolist() and ne == n
        assert m0 == single.miou(0) and m1 == single.miou(1)
        assert ce == single.ce.tolist()


def test_fit_status_words_raise_the_reference_errors():
    """ops.raise_for_status: the deferred form of the reference's ZeroDivisionError (src/test.py:174) / CrossEntropyLoss
    target check, raised from the status words that travel with the counts."""
    from few_shot_seg_cwt_b200 import ops
    ops.raise_for_status(torch.zeros(4, dtype=torch.int32))
    with pytest.raises(ZeroDivisionError, match="episode 6"):
        ops.raise_for_status(torch.tensor([0, 0, _lib.FIT_NO_FG | _lib.FIT_NONFINITE, 0], dtype=torch.int32), first_episode=4)
    with pytest.raises(ValueError):
        ops.raise_for_status(torch.tensor([_lib.FIT_BAD_LABEL], dtype=torch.int32))
    with pytest.raises(FloatingPointError):
        ops.raise_for_status(torch.tensor([0, _lib.FIT_NONFINITE], dtype=torch.int32))


def test_header_is_plain_c():
    """include/cwt_b200.h is a C ABI: it must compile as C (no C++/torch types in the signatures)."""
    for name in ("cwt_b200.h", "cwt_b200_debug.h"):
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                            os.path.join(ROOT, "include", name)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the product package (nor tools/) may import it anywhere
    — module level or inside a function; the sweep's oracle spot-check lives in tests/sweep_oracle_check.py."""
    import ast
    for sub in ("few_shot_seg_cwt_b200", "tools"):
        pkg = os.path.join(ROOT, sub)
        for fn in sorted(os.listdir(pkg)):
            if not fn.endswith(".py"):
                continue
            for node in ast.walk(ast.parse(open(os.path.join(pkg, fn)).read())):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), f"{sub}/{fn} imports the oracle"
    code = "import sys, few_shot_seg_cwt_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_host_pipeline_pieces_are_equal_and_cover_every_episode():
    """HostPipeline._sub_batches: the first host batch of a run goes through in the ramp sizes; every later batch in EQUAL pieces
    of at most ``sub_batch`` episodes (36 -> 18 + 18, not 32 + 4: a 4-episode launch leaves most groups of the persistent fit
    idle); every episode exactly once, in order; an empty batch still yields one (empty) piece so that the result list lines up."""
    from few_shot_seg_cwt_b200.episodic import HostPipeline

    class HB:
        def __init__(self, n):
            self.n_episodes = n

    hp = HostPipeline.__new__(HostPipeline)                     # the generator needs no device
    hp.sub_batch, hp.sub_batch_all, hp.ramp = 32, True, (8, 8, 16)
    for E in (0, 1, 5, 32, 33, 36, 64, 100):
        pieces = list(hp._sub_batches([HB(E), HB(E), HB(E)]))
        for bi in range(3):
            mine = [(lo, hi, last) for b, lo, hi, last, _, _ in pieces if b == bi]
            assert mine[0][0] == 0 and mine[-1][1] == E and mine[-1][2] and not any(l for _, _, l in mine[:-1])
            assert all(a[1] == b[0] for a, b in zip(mine, mine[1:]))            # contiguous, in order
            if bi > 0 and E > 0:
                sizes = [hi - lo for lo, hi, _ in mine]
                assert max(sizes) <= 32 and max(sizes) - min(sizes) < len(sizes), (E, sizes)     # equal up to the rounding
                assert len(sizes) == -(-E // 32)
    assert [(lo, hi) for b, lo, hi, *_ in hp._sub_batches([HB(36), HB(36)]) if b == 1] == [(0, 18), (18, 36)]
    hp.sub_batch_all = False                                     # later batches whole
    assert [(lo, hi) for b, lo, hi, *_ in hp._sub_batches([HB(64), HB(64)]) if b == 1] == [(0, 64)]


def test_kernel_counters_are_matched_per_kernel_source(tmp_path, monkeypatch):
    """bench.load_kernel_counters: a capture is 'of this build' when the fingerprint of the whole library matches OR the sources
    of the kernel the line quotes are unchanged (another kernel may have changed since); otherwise the newest file is
    returned and flagged as not matching."""
    import importlib
    import json
    from few_shot_seg_cwt_b200 import build as B
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    prof = tmp_path / "profiles"
    prof.mkdir()
    kf = B.kernel_fingerprints()
    assert set(kf) >= {"k_fit_resident", "k_fit_l2", "k_logits_iou_stream", "k_kproj_scores"}
    for k, files in B.KERNEL_SOURCES.items():                    # every listed source exists
        for f in files:
            assert os.path.exists(os.path.join(B.CSRC, f)), (k, f)
    json.dump({"build_fingerprint": "stale", "kernels": {}}, open(prof / "a_kernel_counters.json", "w"))
    json.dump({"build_fingerprint": "older build", "kernel_source_fingerprints": dict(kf, k_fit_l2="changed since"), "kernels": {}},
              open(prof / "b_kernel_counters.json", "w"))
    json.dump({"build_fingerprint": "stale too", "kernels": {}}, open(prof / "c_kernel_counters.json", "w"))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    d, name, ok = bench.load_kernel_counters("k_fit_resident")
    assert ok and name.endswith("b_kernel_counters.json")
    d, name, ok = bench.load_kernel_counters("k_fit_l2")         # that kernel's sources changed: newest file, not matching
    assert not ok and name.endswith("c_kernel_counters.json")
    json.dump({"build_fingerprint": B.fingerprint(), "kernels": {}}, open(prof / "0_kernel_counters.json", "w"))
    d, name, ok = bench.load_kernel_counters("k_fit_l2")
    assert ok and name.endswith("0_kernel_counters.json")


def test_zero_compressed_host_format_round_trip_on_cpu():
    """hostformat.compress_map: lossless on the BIT PATTERN (-0.0, NaN, denormals, Inf survive; +0.0 is the only value dropped),
    and what the device kernel consumes — bit mask + one prefix count per block of 32 mask words + packed values — rebuilds the
    tensor when read the way cwt_expand_zero_compressed_f32 reads it (numpy restatement of csrc/expand.cu: value index of an
    element = block prefix + set bits of the earlier words of the block + set bits below it in its word)."""
    from few_shot_seg_cwt_b200 import hostformat as HF
    g = torch.Generator().manual_seed(7)
    t = torch.relu(torch.randn(3, 2, 16, 8, 8, generator=g))                  # 2048 elements per episode = 64 mask words = 2 blocks
    flat = t.view(-1)
    flat[5], flat[77], flat[300], flat[301] = -0.0, float("nan"), 1e-42, float("inf")
    flat[2048:2048 + 40] = 0.0                                                  # a run of empty words at the start of episode 1
    c = HF.compress_map(t)
    assert c.mask.shape == (3, 64) and c.woff.shape == (3, 2) and len(c.val_start) == 4
    assert c.val_start[-1] == c.vals.numel() == int((t.view(torch.int32) != 0).sum())
    assert torch.equal(HF.expand_map_reference(c).view(torch.int32), t.view(torch.int32))
    # the device algorithm, word by word
    mask = c.mask.numpy().view(np.uint32)
    woff = c.woff.numpy().view(np.uint32)
    vals = c.vals.numpy()
    out = np.zeros(t.numel(), dtype=np.float32)
    for row in range(mask.shape[0]):
        for wi in range(mask.shape[1]):
            base = int(woff[row, wi // 32]) + sum(bin(int(m)).count("1") for m in mask[row, (wi // 32) * 32:wi])
            word = int(mask[row, wi])
            for b in range(32):
                if (word >> b) & 1:
                    out[(row * mask.shape[1] + wi) * 32 + b] = vals[base + bin(word & ((1 << b) - 1)).count("1")]
    assert np.array_equal(out.view(np.int32), t.numpy().reshape(-1).view(np.int32))
    # sub-batches are slices: the values of episodes lo..hi are vals[val_start[lo]:val_start[hi]], and nbytes counts exactly those
    assert c.nbytes(1, 3) == 2 * (64 + 2) * 4 + (c.val_start[3] - c.val_start[1]) * 4
    assert c.nbytes() < t.numel() * 4
    with pytest.raises(TypeError):
        HF.compress_map(t.double())
    with pytest.raises(ValueError):
        HF.compress_map(torch.zeros(2, 33))
    e = HF.compress_map(torch.zeros(0, 64))
    assert e.vals.numel() == 0 and e.val_start == [0]
