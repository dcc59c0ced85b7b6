"""CPU-only tests of everything around the kernels: the C ABI surface, flag -> mode mapping,
query sharding + tally reduction over a world_size-2 gloo group, the drop-in packages' import
surface, the rollout restatement, and the reference's caller running unchanged under the shims
(the last two need /root/reference and are skipped on the GPU box)."""
import ast
import ctypes
import json
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vit-reranking_b200")
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


# ---- C ABI --------------------------------------------------------------------------------------
def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vitrerank.h")).read()
    return sorted(set(re.findall(r"VR_API\s+[\w\s\*]+?\b(vr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vitrerank import _lib
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vitrerank.h but not exported"
    assert set(_lib.EXPORTS) == set(names), "ctypes binding and header disagree"
    assert _lib.lib.vr_abi_version() == 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    from vitrerank import _lib
    from vitrerank.engine import RerankEngine
    with pytest.raises(_lib.VitRerankError):
        RerankEngine.get()
    h = ctypes.c_void_p()
    assert _lib.lib.vr_create(0, ctypes.byref(h)) != 0
    assert len(_lib.lib.vr_last_error()) > 0
    import utilities.diml as D
    with pytest.raises(_lib.VitRerankError):
        D.Sinkhorn(torch.ones(1, 2, 2), torch.ones(1, 2) / 2, torch.ones(1, 2) / 2)
    with pytest.raises(_lib.VitRerankError):
        D.calc_similarity(None, torch.ones(4), None, torch.ones(3, 4), 0)


def test_product_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "rerank_oracle" in src:
                    bad.append(os.path.join(base, f))
    assert not bad, bad


# ---- flags -> marginal mode (branch order of evaluate / calc_similarity) ----------------------------
def test_flag_mapping_follows_the_reference_branches():
    from vitrerank.engine import OTParams
    f = OTParams.from_flags
    assert f(use_rollout=True, use_inverse=True, use_minus=True).mode == "rollout"   # eval_cvt_diml.py:344-351
    assert f(use_rollout=True, use_uniform=True).mode == "uniform"                   # diml.py:344-346
    assert f(use_inverse=True).mode == "inverse"
    assert f(use_inverse=True, use_minus=True).mode == "minus"                       # diml.py:80-81
    assert f(use_soft=True).mode == "soft"
    assert f().mode == "relu"
    p = f(use_rollout=True, ot_part=1.0, temperature=0.1)
    s = p.struct()
    assert (s.mode, s.max_iter) == (0, 100) and abs(s.thresh - 0.1) < 1e-8 and abs(s.ot_temp - 0.05) < 1e-8


def test_shard_partition_is_exact():
    from vitrerank.distributed import shard
    for n in (1, 7, 100, 8131):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                s, st, nq = shard(n, r, w)
                seen += list(range(s, n, st))[:nq]
                assert len(range(s, n, st)) == nq
            assert sorted(seen) == list(range(n))


# ---- world_size-2 gloo: sharded evaluation + all-reduce of the tallies ------------------------------
WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
    import numpy as np, torch, torch.distributed as dist
    from oracle import rerank_oracle as O
    from vitrerank import synth, distributed as vd
    from vitrerank.engine import OTParams

    class OracleEngine:   # stands in for the CUDA engine: same evaluate() contract, raw tallies
        device = None
        def __init__(self, g): self.g = g; self.bank = dict(n=g.patches.shape[0])
        def evaluate(self, truncs, params, q_start=0, q_stride=1, nq=None, want_niter=False):
            g = self.g; n = self.bank["n"]
            ids = list(range(q_start, n, q_stride))[:nq]
            out = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=list(truncs),
                                   use_rollout=True, ot_part=1.0, query_ids=ids, dump=True)
            t = np.zeros((len(truncs), 8)); sc = n / 100.0
            for i in range(len(truncs)):
                t[i, 0], t[i, 1], t[i, 2] = out["r1"][i] * sc, out["rp"][i] * sc, out["mapr"][i] * sc
                t[i, 3:7] = np.array(out["recall_at_1_2_4_8"][i]) * sc
                t[i, 7] = len(ids)
            nit = np.array([d["n_iter"] for d in out["dump"]], dtype=np.int32)
            return (t, nit) if want_niter else t

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    torch.set_num_threads(2)
    g = synth.make_gallery(48, 128, 49, classes=4, seed=9, sigma=0.6)
    t, nit = vd.evaluate_sharded(OracleEngine(g), [0, 10], OTParams(mode="rollout"))
    if dist.get_rank() == 0:
        json.dump(dict(t=t.tolist(), nq=int(len(nit))), open({out!r}, "w"))
    dist.barrier(); dist.destroy_process_group()
''')


def test_two_rank_gloo_sharding_matches_single_process(tmp_path):
    from oracle import rerank_oracle as O
    from vitrerank import synth
    out = str(tmp_path / "t.json")
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, pkg=PKG, port=port, out=out))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = json.load(open(out))
    g = synth.make_gallery(48, 128, 49, classes=4, seed=9, sigma=0.6)
    ref = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=[0, 10], use_rollout=True, ot_part=1.0)
    t = np.array(got["t"])
    assert got["nq"] == 24 and t[0, 7] == 48
    np.testing.assert_allclose(t[:, 0] / 0.48, ref["r1"], rtol=1e-12)
    np.testing.assert_allclose(t[:, 1] / 0.48, ref["rp"], rtol=1e-9)
    np.testing.assert_allclose(t[:, 2] / 0.48, ref["mapr"], rtol=1e-9)


# ---- drop-in packages keep the reference's import surface -------------------------------------------
def test_dropin_import_surface():
    import inspect
    import utilities.diml as D
    import evaluation.eval_cvt_diml as E
    import evaluation.eval_diml as E_res
    import evaluation.eval_attn_diml as E_attn
    import evaluation.eval_swin_diml as E_swin
    import evaluation.metrics as M
    import evaluation  # noqa: F401  (must not need faiss)
    for name in ("Sinkhorn", "Sinkhorn_partial", "calc_similarity", "calc_distance", "calc_similarity_vit",
                 "calc_similarity_cvt", "calc_similarity_cvt_rollout", "calc_similarity_featvit",
                 "calc_similarity_mhvit", "input_inv_transform"):
        assert callable(getattr(D, name))
    assert list(inspect.signature(D.calc_similarity).parameters) == [
        "anchor", "anchor_center", "fb", "fb_center", "stage", "use_uniform", "use_inverse", "temperature",
        "use_cls_token", "ot_temp", "use_minus", "ot_part", "use_soft"]
    assert list(inspect.signature(D.calc_similarity_cvt_rollout).parameters) == [
        "anchor_center", "anchor", "anchor_query", "fb_center", "fb", "fb_key", "stage", "use_uniform", "ot_temp",
        "use_ot", "ot_part", "device"]
    ev = list(inspect.signature(E.evaluate).parameters)
    assert ev[:17] == ["model", "dataset", "dataloader", "training", "trunc_nums", "use_uniform", "grid_size",
                       "use_inverse", "temperature", "use_cls_token", "attn_blk_ind", "use_ot", "ot_part", "to_submit",
                       "use_minus", "use_rollout", "plot_topk"]
    assert callable(M.get_metrics_rank) and callable(M.get_metrics) and callable(E.evaluate_patch_similarity)
    x = np.zeros((3, 4, 4), dtype=np.float32)
    assert D.input_inv_transform(x).shape == (4, 4, 3)


@needs_ref
def test_signatures_match_the_reference_source():
    """Positional / keyword names and defaults of every public function, read from the reference's source
    with ast (the reference modules themselves need CUDA / cv2 / matplotlib to import)."""
    import inspect
    import utilities.diml as D
    import evaluation.eval_cvt_diml as E
    import evaluation.eval_diml as E_res
    import evaluation.eval_attn_diml as E_attn
    import evaluation.eval_swin_diml as E_swin
    import evaluation.metrics as M

    def ref_sigs(path):
        tree = ast.parse(open(path).read())
        out = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef):
                a = node.args
                names = [x.arg for x in a.args]
                defaults = [ast.literal_eval(d) if not isinstance(d, (ast.Call, ast.Attribute)) else "<expr>"
                            for d in a.defaults]
                out[node.name] = (names, defaults)
        return out

    for mod, path, names in ((D, "utilities/diml.py", ["Sinkhorn", "Sinkhorn_partial", "calc_similarity",
                                                       "calc_distance", "calc_similarity_vit", "calc_similarity_cvt",
                                                       "calc_similarity_cvt_rollout", "calc_similarity_featvit",
                                                       "calc_similarity_mhvit", "input_inv_transform"]),
                             (E, "evaluation/eval_cvt_diml.py", ["evaluate", "evaluate_patch_similarity",
                                                                 "get_attention_rollout", "filter_attention_map",
                                                                 "resize_attn_map"]),
                             (E_res, "evaluation/eval_diml.py", ["evaluate"]),          # the sibling callers of the same loop
                             (E_attn, "evaluation/eval_attn_diml.py", ["evaluate"]),
                             (E_swin, "evaluation/eval_swin_diml.py", ["evaluate"]),
                             (M, "evaluation/metrics.py", ["get_metrics", "get_metrics_rank"])):
        sigs = ref_sigs(os.path.join(REF, path))
        for n in names:
            rn, rd = sigs[n]
            params = list(inspect.signature(getattr(mod, n)).parameters.values())
            mine = [p.name for p in params][:len(rn)]
            assert mine == rn, (n, mine, rn)
            mydef = [p.default for p in params[:len(rn)] if p.default is not inspect._empty]
            for a, b in zip(mydef, rd):
                if b != "<expr>":
                    assert a == b, (n, mydef, rd)


# ---- rollout restatement against the reference's own functions ---------------------------------------
@needs_ref
def test_attention_rollout_matches_reference():
    src = open(os.path.join(REF, "evaluation", "eval_cvt_diml.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and
            n.name in ("resize_attn_map", "filter_attention_map", "get_attention_rollout")]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "ref_rollout", "exec"), ns)
    sys.path.insert(0, os.path.join(PKG, "shims", "override"))
    try:
        import importlib
        archs = importlib.import_module("architectures")
    finally:
        sys.path.pop(0)
    import evaluation.eval_cvt_diml as E

    class Opt:
        arch, embed_dim, seed = "cvt_13_normalize", 128, 3
    model = archs.select("cvt_13_normalize", Opt()).eval()
    img = torch.randn(5, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    ref = ns["get_attention_rollout"](model.model, img, display_map=False)
    got = E.get_attention_rollout(model.model, img, display_map=False)
    assert len(ref) == len(got) == 2
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    assert got[-1].mean(1).shape == (5, 49)


# ---- the reference's caller, unchanged, under the shims (evaluate replaced by a recorder) ---------------
@needs_ref
def test_reference_caller_runs_unchanged_under_shims(tmp_path):
    rec = str(tmp_path / "call.json")
    driver = tmp_path / "driver.py"
    driver.write_text(textwrap.dedent(f'''
        import json, runpy, sys, torch
        torch.nn.Module.to = lambda self, *a, **k: self          # no GPU in this container
        import evaluation.eval_cvt_diml as E
        def fake_evaluate(model, dataset, dataloader, training, trunc_nums, **kw):
            assert hasattr(model, "pars") and hasattr(model.model, "head") and len(dataset) > 0
            label, img, _ = dataset[0]
            assert tuple(img.shape) == (3, 224, 224)
            json.dump(dict(trunc_nums=trunc_nums, training=training, **kw), open({rec!r}, "w"))
            return {{'r1': [1.0] * len(trunc_nums), 'rp': [2.0] * len(trunc_nums), 'mapr': [3.0] * len(trunc_nums)}}
        E.evaluate = fake_evaluate
        sys.argv = ["test_diml_cvt.py", "--dataset", "cub200", "--group", "t", "--arch", "cvt_13_normalize",
                    "--embed_dim", "128", "--bs", "16", "--samples_per_class", "2", "--not_pretrained",
                    "--use_ot", "--use_inverse", "--use_rollout", "--grid_size", "7", "--temperature", "0.1",
                    "--ot_part", "1.0", "--source_path", "/tmp", "--kernels", "0"]
        runpy.run_path("{REF}/test_diml_cvt.py", run_name="__main__")
    '''))
    env = dict(os.environ)
    env["PYTHONSAFEPATH"] = "1"
    env["VITRERANK_SHIM_N"] = "32"
    env["PYTHONPATH"] = os.pathsep.join([PKG, os.path.join(PKG, "shims", "override"), REF,
                                         os.path.join(PKG, "shims", "fallback")])
    r = subprocess.run([sys.executable, str(driver)], cwd=str(tmp_path), env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    call = json.load(open(rec))
    assert call["trunc_nums"] == [0, 100]                                  # test_diml_cvt.py:130
    assert call["use_rollout"] and call["use_inverse"] and call["use_ot"]
    assert call["grid_size"] == 7 and abs(call["temperature"] - 0.1) < 1e-12 and call["ot_part"] == 1.0
    csv = (tmp_path / "test_results" / "test_diml_cub200.csv").read_text()  # :158-161
    assert "method,r1,rp,mapr" in csv.replace(" ", "") and "ours (100)" in csv


def test_bank_file_roundtrip_and_rejects_damage(tmp_path):
    """vitrerank/bankfile.py: the on-disk bank (the cache the reference keeps disabled, evaluation/eval_diml.py:80-85,
    151-153) round-trips bit for bit, mmap or not, with and without the optional banks, and refuses damaged files."""
    from vitrerank import bankfile as B, synth
    g = synth.make_gallery(37, 16, 9, classes=3, seed=2)
    p = str(tmp_path / "a.vrbank")
    nbytes = B.save(p, g.patches, g.centers, g.rollout, g.labels)
    assert os.path.getsize(p) == nbytes
    for mm in (True, False):
        pt, ce, ro, la = B.load(p, mmap=mm)
        assert torch.equal(pt, g.patches) and torch.equal(ce, g.centers) and torch.equal(ro, g.rollout)
        assert torch.equal(la, g.labels)
    B.save(p, g.patches, g.centers)
    pt, ce, ro, la = B.load(p)
    assert ro is None and la is None and torch.equal(pt, g.patches)
    raw = bytearray(open(p, "rb").read())
    bad = str(tmp_path / "b.vrbank")
    for pos, what in ((3, "not a vitrerank bank"), (7, "version"), (200, "checksum")):
        dmg = bytearray(raw)
        dmg[pos] ^= 0x20
        open(bad, "wb").write(dmg)
        with pytest.raises(B.BankFileError, match=what):
            B.load(bad)
    open(bad, "wb").write(raw[:-64])
    with pytest.raises(B.BankFileError, match="header says"):
        B.load(bad)
    with pytest.raises(B.BankFileError):
        B.save(bad, g.patches, g.centers[:5])


def test_partial_ot_pad_is_the_reference_rounding():
    """diml.py:61 forms `K.new_tensor(1 - ot_part)`: a double subtraction of the Python float, rounded once to fp32.  The
    ABI carries ot_part as fp32; vr_partial_ot_pad recovers the decimal the caller typed.  (Host arithmetic: runs without a GPU.)"""
    import ctypes as C
    from vitrerank import _lib
    for i in range(0, 1000):
        part = i / 1000.0
        want = np.float32(1 - part)
        got = np.float32(_lib.lib.vr_partial_ot_pad(C.c_float(part)))
        assert got == want, (part, got, want)
    for part in (0.123456, 0.3333333, 1 / 3, 0.7071067811865476):
        got = np.float32(_lib.lib.vr_partial_ot_pad(C.c_float(part)))
        assert abs(float(got) - (1 - part)) < 1e-7


def test_rollout_workspace_is_the_fused_map():
    """vr_rollout_block_workspace_bytes (host arithmetic, no GPU): the fused [b, ht * wt] map, the per-image histograms and the
    union mask -- the select runs over the WHOLE map, cls row and column included (eval_cvt_diml.py:74-108 filters before :57-58 drops)."""
    from vitrerank import _lib
    f = _lib.lib.vr_rollout_block_workspace_bytes
    for b, ht, wt in ((1, 50, 50), (64, 197, 197), (8, 3136, 784)):
        n = f(b, ht, wt, 0)
        assert n == f(b, ht, wt, 1)
        assert b * ht * wt * 4 + ht * wt <= n <= b * ht * wt * 4 + ht * wt + b * 2048 * 4 + b * 16 + 2048
    assert f(0, 50, 50, 0) == 0
