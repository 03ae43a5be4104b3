"""CPU: the C-ABI library loads and exports every symbol include/ralf_b200.h declares (no compute calls),
and the drop-in classes keep the reference's state-dict contract."""
import json
import os
import re

import numpy as np
import pytest
import torch

from tests import helpers

ROOT = helpers.ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ralf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ralf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ralf_b200 import _lib

    lib = _lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ralf_b200.h but not exported"
    assert lib.ralf_version() >= 100


def test_product_path_fails_loudly_without_library(monkeypatch):
    from ralf_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libralf_b200.so")
    with pytest.raises(_lib.RalfError):
        _lib.lib()


def test_no_oracle_import_in_product_code():
    """The product package must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "ralf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "libknn_oracle" not in txt or f == "build.py", f


@pytest.mark.parametrize("name,is_ralf", [("ralf_cgl", True), ("autoreg_cgl", False)])
def test_state_dict_contract(name, is_ralf):
    """Key names, order, shapes and dtypes equal the reference class's state_dict (strict-load contract)."""
    from ralf_b200 import generator as G

    cls = G.ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg if is_ralf else G.ConcateAuxilaryTaskAutoreg
    tok = helpers.make_tokenizer()
    model = cls(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, db_dataset=None,
                retrieval_backbone="dreamsim", random_retrieval=False, top_k=16, saliency_k="None", auxilary_task="uncond")
    ref = helpers.load_schema(name)
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert list(v.shape) == ref[k]["shape"], k
        assert str(v.dtype).replace("torch.", "") == ref[k]["dtype"], k
    # a reference checkpoint loads strictly; frozen FIDNet stays frozen (retrieval_augmented_autoreg.py:150-154)
    model.load_state_dict(helpers.synth_weights(name, 3), strict=True)
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_all = sum(p.numel() for p in model.parameters())
    if is_ralf:
        assert (n_all, n_train) == (44386946, 42801282)
    else:
        assert n_all == 41224066
    with pytest.raises(RuntimeError):
        model.engine()  # CPU: no fallback


def test_uncond_constraint_sequence_matches_reference():
    from ralf_b200 import generator as G

    z, meta = helpers.load_golden("ralf_cgl_256")
    tok = helpers.make_tokenizer()
    pre = G.UnconditionalPreprocessor(tok)
    assert pre.N_total == 549
    out = pre(G.ConditionalInputs(image=torch.zeros(2, 4, 8, 8)))
    assert out["seq"].tolist() == z["seq_layout_const"].tolist()
    assert out["pad_mask"].tolist() == z["seq_layout_const_pad_mask"].tolist()


def test_optim_groups_match_reference():
    """models/common/base_model.py:207-347 through train.py:217-223's call: same groups, order, lr and weight decay
    (fixture dumped from the reference class by tests/golden/make_golden.py)."""
    import json
    import os

    from ralf_b200 import generator as G
    from tests import helpers

    with open(os.path.join(helpers.GOLDEN, "optim_groups_ralf_cgl.json")) as f:
        ref = json.load(f)
    model = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10)
    names = {id(p): n for n, p in model.named_parameters()}
    groups = model.optim_groups(base_lr=1e-4, weight_decay=1e-4, custom_lr={"encoder.extractor.body": 1e-5})
    mine = [{"lr": g["lr"], "weight_decay": g["weight_decay"], "params": [names[id(p)] for p in g["params"]]} for g in groups]
    assert mine == ref
    plain = model.optim_groups(base_lr=1e-4, weight_decay=1e-4)
    assert len(plain) == 2 and sum(len(g["params"]) for g in plain) == sum(len(g["params"]) for g in ref)


def test_state_dict_contract_pku():
    """BASELINE configs[2] names the PKU dataset (3 labels -> vocabulary 518, constraint vocabulary 548): key names, order,
    shapes and dtypes of the reference class built for PKU (tests/golden/schema_ralf_pku.json, dumped from the reference),
    and the PKU tokenizer's ids / per-position mask (tests/golden/tokenizer_pku.npz)."""

    from oracle import synth
    from ralf_b200 import generator as G

    tok = helpers.make_tokenizer("pku")
    model = G.RALF(features=None, tokenizer=tok, dataset_name="pku", max_seq_length=10, top_k=16, auxilary_task="uncond")
    ref = helpers.load_schema("ralf_pku")
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert list(v.shape) == ref[k]["shape"] and str(v.dtype).replace("torch.", "") == ref[k]["dtype"], k
    z = np.load(helpers.GOLDEN + "/tokenizer_pku.npz")
    b = synth.synth_batch(3, 8, 8, 10, 1, tok.N_label, seed=21)
    enc = tok.encode({k: b[k] for k in ["label", "mask", "center_x", "center_y", "width", "height"]})
    np.testing.assert_array_equal(enc["seq"].numpy(), z["seq"])
    np.testing.assert_array_equal(enc["mask"].numpy(), z["mask"])
    np.testing.assert_array_equal(tok.token_mask.numpy(), z["token_mask"])
    assert [tok.name_to_id("pad"), tok.name_to_id("bos"), tok.name_to_id("eos")] == z["special"].tolist()


def test_dynamic_top_k_reaches_existing_engines():
    """inference.py:290,346 reads and re-assigns model.top_k between runs."""
    import types

    from ralf_b200 import generator as G

    m = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10, top_k=16)
    assert m.top_k == 16
    m._engine = types.SimpleNamespace(top_k=16)
    m._train_engine = types.SimpleNamespace(infer=types.SimpleNamespace(top_k=16))
    m.top_k = 8
    assert m.top_k == 8 and m._engine.top_k == 8 and m._train_engine.infer.top_k == 8


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--hw", "64", "--gallery", "2000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "layouts/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("layouts/sec") and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "layouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # other ranks of a torchrun launch print nothing and exit 0
    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_postprocess_decodes_sequences_and_masked_logits():
    """BaseModel.postprocess (base_model.py:367-389)."""
    from ralf_b200 import generator as G

    tok = helpers.make_tokenizer()
    m = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10)
    z, _ = helpers.load_golden("ralf_cgl_256")
    seq = torch.from_numpy(z["gen_seq"])
    out = m.postprocess({"seq": seq})
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(out[k].numpy(), z["gen_" + k])
    logits = torch.from_numpy(z["gen_step_logits"])  # per-step logits of the greedy loop: argmax under the mask = gen_seq
    noisy = torch.where(torch.isfinite(logits), logits, torch.full_like(logits, 1e9))  # the mask must be re-applied here
    out2 = m.postprocess({"logits": noisy})
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(out2[k].numpy(), z["gen_" + k])


@pytest.mark.parametrize("cls_name,stats_key", [("RALF", "ralf_cgl"), ("ConcateAuxilaryTaskAutoreg", "autoreg_cgl")])
def test_initial_values_follow_the_reference_distributions(cls_name, stats_key):
    """A from-scratch training run must start where the reference's does: per state-dict entry, the freshly constructed
    drop-in class against the statistics of the freshly constructed reference class (tests/golden/init_stats.json).
    The ResNet50 trunk and the FIDNet layout encoder come from checkpoint files in both (next test)."""
    from ralf_b200 import generator as G

    with open(os.path.join(helpers.GOLDEN, "init_stats.json")) as f:
        ref = json.load(f)[stats_key]
    torch.manual_seed(99)
    m = getattr(G, cls_name)(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10)
    checked = 0
    for k, v in m.state_dict().items():
        if k not in ref or k.startswith(("encoder.extractor.body.", "layout_encoer.")):
            continue
        mean, std, lo, hi, numel = ref[k]
        v = v.double()
        assert v.numel() == numel, k
        if std == 0.0 and numel > 1:  # constants: LayerNorm gains / shifts, attention biases
            assert float(v.min()) == lo and float(v.max()) == hi, k
        elif numel >= 4096:
            assert abs(float(v.std()) / std - 1) < 0.03, (k, float(v.std()), std)
            assert abs(float(v.mean())) < 0.05 * std, k
            if hi < 4 * std:  # a uniform distribution: the bounds agree too
                assert abs(float(v.max()) / hi - 1) < 0.02 and abs(float(v.min()) / lo - 1) < 0.02, k
        else:  # small tensors (biases): inside the reference's support, spread of the right order
            bound = max(abs(lo), abs(hi))
            assert float(v.abs().max()) <= bound * 1.6 + 1e-12, k
            if numel >= 64:
                assert 0.6 < float(v.std()) / std < 1.6, (k, float(v.std()), std)
        checked += 1
    assert checked > 150
    # drawn from the global generator, like the reference: another seed, another model; same seed, same model
    torch.manual_seed(99)
    m2 = getattr(G, cls_name)(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10)
    assert torch.equal(m.state_dict()["decoder.emb.weight"], m2.state_dict()["decoder.emb.weight"])
    m3 = getattr(G, cls_name)(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10)
    assert not torch.equal(m.state_dict()["decoder.emb.weight"], m3.state_dict()["decoder.emb.weight"])


def test_constructor_reads_the_reference_weight_files(tmp_path, monkeypatch):
    """common/image.py:36-48,70-78 and fid/model.py:131-169: ImageNet ResNet50 (4th stem channel = mean of the three) and
    the frozen FIDNetV3 checkpoint, from the locations the reference looks in ("pku" -> "pku10")."""
    from ralf_b200 import generator as G

    g = torch.Generator().manual_seed(3)
    probe = G.RALF(features=None, tokenizer=helpers.make_tokenizer("pku"), dataset_name="pku", max_seq_length=10)
    own = probe.state_dict()
    resnet = {k[len("encoder.extractor.body."):]: torch.randn(v.shape, generator=g) if v.is_floating_point() else v.clone()
              for k, v in own.items() if k.startswith("encoder.extractor.body.")}
    resnet["conv1.weight"] = torch.randn(64, 3, 7, 7, generator=g)
    resnet["fc.weight"], resnet["fc.bias"] = torch.randn(1000, 2048, generator=g), torch.randn(1000, generator=g)
    fid = {k[len("layout_encoer."):]: torch.randn(v.shape, generator=g) if v.is_floating_point() else v.clone()
           for k, v in own.items() if k.startswith("layout_encoer.")}
    fid["pos_token"] = torch.randn(10, 1, 256, generator=g)  # parts of the checkpoint the feature extractor drops
    fid["fc_out_disc.weight"] = torch.randn(1, 256, generator=g)
    os.makedirs(tmp_path / "cache" / "PRECOMPUTED_WEIGHT_DIR" / "fidnet" / "pku10")
    torch.save(resnet, tmp_path / "cache" / "PRECOMPUTED_WEIGHT_DIR" / "resnet50_a1_0-14fe96d1.pth")
    torch.save({"state_dict": fid, "epoch": 7}, tmp_path / "cache" / "PRECOMPUTED_WEIGHT_DIR" / "fidnet" / "pku10" / "model_best.pth.tar")
    monkeypatch.chdir(tmp_path)
    m = G.RALF(features=None, tokenizer=helpers.make_tokenizer("pku"), dataset_name="pku", max_seq_length=10)
    sd = m.state_dict()
    w = sd["encoder.extractor.body.conv1.weight"]
    assert torch.equal(w[:, :3], resnet["conv1.weight"]) and torch.allclose(w[:, 3], resnet["conv1.weight"].mean(dim=1))
    for k, v in resnet.items():
        if k not in ("conv1.weight", "fc.weight", "fc.bias"):
            assert torch.equal(sd["encoder.extractor.body." + k], v), k
    for k, v in fid.items():
        if k not in ("pos_token", "fc_out_disc.weight"):
            assert torch.equal(sd["layout_encoer." + k], v), k
    assert not any(p.requires_grad for n, p in m.named_parameters() if n.startswith("layout_encoer."))
    # pretrained=False and the Autoreg class (no layout encoder) leave / skip as expected
    m0 = G.RALF(features=None, tokenizer=helpers.make_tokenizer("pku"), dataset_name="pku", max_seq_length=10, pretrained=False)
    assert not torch.equal(m0.state_dict()["encoder.extractor.body.layer1.0.conv1.weight"], resnet["layer1.0.conv1.weight"])
    ar = G.ConcateAuxilaryTaskAutoreg(features=None, tokenizer=helpers.make_tokenizer("pku"), dataset_name="pku")
    assert torch.equal(ar.state_dict()["encoder.extractor.body.layer1.0.conv1.weight"], resnet["layer1.0.conv1.weight"])


def test_constructor_rejects_configurations_that_are_not_built():
    """Reference constructor options that change the architecture are refused, never silently ignored."""
    from ralf_b200 import generator as G

    tok = helpers.make_tokenizer()
    ok = dict(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, db_dataset=None, weight_init=True,
              use_reference_image=False, layout_backbone="feature_extractor", freeze_layout_encoder=True, decoder_d_model=256,
              shared_embedding=False, global_task_embedding=False)
    G.RALF(**ok)
    for bad in [{"use_reference_image": True}, {"saliency_k": "dynamic"}, {"decoder_d_model": 512}, {"d_model": 128},
                {"global_task_embedding": True}, {"shared_embedding": True}, {"use_flag_embedding": False}]:
        with pytest.raises(NotImplementedError):
            G.RALF(**{**ok, **bad})
    with pytest.raises(TypeError):
        G.RALF(**ok, no_such_option=1)


def test_knn_entry_points_validate_before_touching_the_device():
    """Error behaviour of the C ABI (include/ralf_b200.h: RalfStatus): bad arguments come back as negative status codes
    from the argument checks, before any CUDA call -- so this runs without a GPU.  NULL -> RALF_ERR_NULL (-3), shapes ->
    RALF_ERR_SHAPE (-1), row pitch -> RALF_ERR_ALIGN (-2), small workspace -> RALF_ERR_WORKSPACE (-7)."""
    import ctypes as C

    from ralf_b200 import _lib

    L = _lib.lib()
    assert L.ralf_version() >= 100
    L.ralf_knn_workspace_bytes.restype = C.c_size_t
    ws = L.ralf_knn_workspace_bytes(1_000_000, 512, 128, 16)
    assert 0 < ws < (1 << 30) and L.ralf_knn_workspace_bytes(1_000_000, 512, 128, 33) >= ws
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)
    z = C.c_float(0.0)
    for fn, extra in ((L.ralf_knn_topk, True), (L.ralf_knn_topk_exact, False)):
        def call(g, n, d, q, nq, k, oi, os_, w, wb):
            if extra:
                return fn(g, n, d, q, nq, k, 0, z, oi, os_, None, w, wb, None)
            return fn(g, n, d, q, nq, k, 0, oi, os_, w, wb, None)

        assert call(None, 10, 512, p, 1, 16, p, p, p, ws) == -3
        assert call(p, 10, 512, p, 1, 16, None, p, p, ws) == -3
        assert call(p, 0, 512, p, 1, 16, p, p, p, ws) == -1
        assert call(p, 10, 512, p, 0, 16, p, p, p, ws) == -1
        assert call(p, 10, 512, p, 1, 64, p, p, p, ws) == -1        # k beyond the 48 the candidate filter keeps
        assert call(p, 10, 512, p, 1, 16, p, p, p, 8) == -7
        assert call(p, 10, 512, p, 1, 16, p, p, None, ws) == -7
    assert L.ralf_knn_topk(p, 10, 510, p, 1, 16, 0, z, p, p, None, p, ws, None) == -2  # TMA needs 16-byte rows
    assert L.ralf_knn_merge(None, p, 2, 1, 16, p, p, None) == -3
    assert L.ralf_knn_merge(p, p, 0, 1, 16, p, p, None) == -1
    if not torch.cuda.is_available():  # no device: a status code and a message, not a crash
        assert L.ralf_check_device(0) < 0
        L.ralf_last_cuda_error.restype = C.c_char_p
        assert len(L.ralf_last_cuda_error()) > 0


def test_round2_entry_points_validate_before_touching_the_device():
    """ralf_knn_fixup_exact / ralf_decode_chain / ralf_attention_decode_kv16: argument checks return RalfStatus codes before
    any CUDA call (runs without a GPU)."""
    import ctypes as C

    from ralf_b200 import _lib

    L = _lib.lib()
    buf = (C.c_char * 4096)()
    p = C.cast(buf, C.c_void_p)
    ws = L.ralf_knn_workspace_bytes(10, 512, 1, 16)
    assert L.ralf_knn_fixup_exact(p, 10, 512, p, 1, 16, 0, None, p, p, p, ws, None) == -3        # certified is required
    assert L.ralf_knn_fixup_exact(p, 10, 512, p, 1, 64, 0, p, p, p, p, ws, None) == -1           # k beyond the filter's 48
    assert L.ralf_knn_fixup_exact(p, 10, 512, p, 1, 16, 0, p, p, p, p, 8, None) == -7            # workspace too small
    assert L.ralf_attention_decode_kv16(None, 256, p, 10, 10, 1, 8, C.c_float(1.0), p, 0, 256, None) == -3
    assert L.ralf_attention_decode_kv16(p, 256, p, 10, 10, 1, 4, C.c_float(1.0), p, 0, 256, None) == -1   # row format: 8 heads
    st = (_lib.ChainStage * 2)()
    assert L.ralf_decode_chain(p, 256, 4, None, 1, None) == -3
    assert L.ralf_decode_chain(p, 256, 4, st, 5, None) == -1                                     # at most 4 stages
    st[0].W, st[0].ldw, st[0].n_out, st[0].k_in, st[0].in_mode = p.value, 256, 256, 512, 1       # K must be 256 or 1024
    assert L.ralf_decode_chain(p, 256, 4, st, 1, None) == -1
    st[0].k_in = 256                                                                             # LayerNorm input without gamma / beta
    assert L.ralf_decode_chain(p, 256, 4, st, 1, None) == -3
    st[0].gamma, st[0].beta, st[0].out_operand = p.value, p.value, 1                             # operand output needs n_out = 1024 and a next stage
    assert L.ralf_decode_chain(p, 256, 4, st, 1, None) == -1


def test_late_round2_entry_points_validate_before_touching_the_device():
    """ralf_conv_gemm_strided / ralf_gemm_res_ln: shape, null and alignment checks return RalfStatus codes before any CUDA
    call (runs without a GPU)."""
    import ctypes as C

    from ralf_b200 import _lib

    L = _lib.lib()
    buf = (C.c_char * 4096)()
    p = (C.addressof(buf) + 255) & ~255  # 256-byte aligned scratch address (never dereferenced)
    g = _lib.GemmArgs()
    g.A, g.W, g.lda, g.ldw, g.npass = p, p, 64, 576, 3
    g.M, g.N, g.K = 2 * 32 * 32, 64, 9 * 64
    assert L.ralf_conv_gemm_strided(C.byref(g), 2, 64, 64, 64, 3, 3, 3, None) == -1          # stride 1 or 2 only
    assert L.ralf_conv_gemm_strided(C.byref(g), 2, 64, 64, 64, 3, 3, 1, None) == -1          # M must be B * Ho * Wo of THAT stride
    g.M = 2 * 32 * 31
    assert L.ralf_conv_gemm_strided(C.byref(g), 2, 64, 64, 64, 3, 3, 2, None) == -1
    g.A = None
    g.M = 2 * 32 * 32
    assert L.ralf_conv_gemm_strided(C.byref(g), 2, 64, 64, 64, 3, 3, 2, None) == -3
    r = _lib.GemmArgs()
    r.A, r.W, r.lda, r.ldw, r.npass, r.M, r.N, r.K = p, p, 256, 256, 3, 8, 256, 256
    r.res, r.res_ld, r.out_f32, r.out_ld = p, 256, p, 256
    f = C.c_float(1e-5)
    assert L.ralf_gemm_res_ln(C.byref(r), None, p, f, p, 8 * 256, None) == -3                  # gamma / beta / ln_split are required
    r.N = 512
    assert L.ralf_gemm_res_ln(C.byref(r), p, p, f, p, 8 * 256, None) == -1                     # the LayerNorm row is d_model = 256
    r.N, r.K = 256, 200
    assert L.ralf_gemm_res_ln(C.byref(r), p, p, f, p, 8 * 256, None) == -1                     # K in whole k-blocks
    r.K, r.act = 256, 1
    assert L.ralf_gemm_res_ln(C.byref(r), p, p, f, p, 8 * 256, None) == -1                     # no activation in this epilogue
    r.act, r.res_ld = 0, 255
    assert L.ralf_gemm_res_ln(C.byref(r), p, p, f, p, 8 * 256, None) == -2                     # rows must allow 16-byte accesses
    r.res_ld, r.res = 256, None
    assert L.ralf_gemm_res_ln(C.byref(r), p, p, f, p, 8 * 256, None) == -3


def test_bench_line_assembly_with_stub_measurements():
    """The block of bench.py that turns the measurements into the contract's JSON line, executed here with stub numbers (the
    measurements themselves need a B200): every key the driver reads is present and the arithmetic holds together."""
    import textwrap
    import types

    import bench

    src = open(os.path.join(ROOT, "bench.py")).read()
    a = src.index("    if rank == 0:\n        peaks = {}")
    b = src.index("    else:\n        line = None\n", a)
    block = textwrap.dedent(src[a:b])

    class Ev:
        def elapsed_time(self, other):
            return 2.9  # ms for the 8 k-NN passes of a step

    sized = lambda n: types.SimpleNamespace(numel=lambda: n, shape=(1_000_000, 512))
    ns = dict(vars(bench))
    ns.update(rank=0, world=1, B=1024, HW=256, S=60, ms=930.0, ms_e2e=940.0, knn_ev=[(Ev(), Ev())] * 5, launches=28910,
              args=types.SimpleNamespace(steps=5, warmup=3, precision="bf16x3", micro_batch=128, gallery=1_000_000, elems=12,
                                         overlap=False, decode_ways=1, no_cpu_baseline=True, hw=256),
              retr=types.SimpleNamespace(emb=sized(1)), model=types.SimpleNamespace(special_token_ids={}),
              img_h=sized(1024 * 4 * 256 * 256), qry_h=sized(1024 * 512), clocks={"sm_mhz": 1900, "sm_max_mhz": 1965, "reasons": []},
              other=[{"kernel": "a", "bound": "hbm", "achieved": 6250.9, "unit": "GB/s", "ms_per_launch": 0.1342, "launches_per_step": 360},
                     {"kernel": "b", "bound": "tensor", "achieved": 1273.4, "unit": "TFLOP/s", "ms_per_launch": 0.0911, "launches_per_step": 48},
                     {"kernel": "c", "bound": "hbm", "achieved": 6050.0, "unit": "GB/s", "ms_per_launch": 0.2, "launches_per_step": 24}],
              api={"value": 900.0, "unit": "layouts/s"}, phases={"search_ms": 2.9, "encode_ms": 84.0, "decode_ms": 62.0})
    exec(block, ns)
    d = json.loads(json.dumps(ns["line"]))  # what emit() prints
    for key in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"]:
        assert key in d, key
    assert d["value"] == round(1024 / 0.186, 2) and d["ms_per_step"] == 186.0 and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] == 1075838976
    r = d["roofline"]  # the dominant kernel: the decode cross-attention stream (HBM bound), with the ncu traffic figure
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["share_of_step"] == round(0.1342 * 360 / 186.0, 4) and r["traffic"] == 593771776 + 6663424
    k = d["roofline_knn"]  # the k-NN half of the metric, timed live
    assert k["bound"] == "hbm" and abs(k["frac"] - k["achieved"] / k["peak"]) < 1e-3 and 0 < k["share_of_step"] < 0.05
    assert "workload" in d["config"] and "model" not in d["config"] and d["gpu_launches"] == 28910
    assert d["phases"]["decode_ms"] == 62.0
    assert [o["share_of_step"] for o in d["roofline_other"]] == [round(0.0911 * 48 / 186.0, 4), round(0.2 * 24 / 186.0, 4)]
    assert [o["kernel"] for o in d["roofline_other"]] == ["b", "c"] and abs(d["roofline_other"][1]["frac"] - 6050.0 / r["peak"]) < 1e-3
