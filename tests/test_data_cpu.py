"""CPU: wire formats of the retrieval tables, host collation (f1), LR schedule / checkpoint files (f4),
and the reference's own property checks for the task preprocessors (tests/train/helpers/test_task_preprocessor.py:28-58)."""
import collections
import os

import pytest
import torch

from tests import helpers


def test_cache_table_roundtrip_and_reference_pickle(tmp_path):
    from ralf_b200 import data as D

    table = {i: [(i * 7 + j) % 100 for j in range(33)] for i in range(20)}
    p = D.cache_table_path("cgl", "val", "dreamsim", 32, root=str(tmp_path))
    assert os.path.basename(p) == "cgl_val_dreamsim_wo_head_table_between_dataset_indexes_top_k32.pt"
    D.save_cache_table(table, p)
    assert D.load_cache_table(p, 16) == {k: v[:16] for k, v in table.items()}
    # what the reference writes: a pickled defaultdict(list) (retriever.py:189,226)
    dd = collections.defaultdict(list)
    dd.update(table)
    torch.save(dd, p)
    assert D.load_cache_table(p, 16) == {k: v[:16] for k, v in table.items()}
    # and what the reference's loader does with ours: torch.load + slicing (retrieval_dataset_wrapper.py:30-32)
    D.save_cache_table(table, p)
    ref = {k: v[:16] for k, v in torch.load(p).items()}
    assert ref == {k: v[:16] for k, v in table.items()}
    with pytest.raises(ValueError):
        D.load_cache_table(str(tmp_path / "missing.pt"), 16)


def test_retrieval_yaml_roundtrip(tmp_path):
    from ralf_b200 import data as D

    db_ids = [str(1000 + i) for i in range(50)]
    table = {str(i): [(i + j) % 50 for j in range(20)] for i in range(5)}
    p = str(tmp_path / "cgl" / "val.yaml")
    D.export_retrieval_yaml(table, db_ids, p, top_k=16)
    raw = D.load_retrieval_yaml(p)
    assert raw["3"] == [db_ids[(3 + j) % 50] for j in range(16)]
    back = D.load_retrieval_yaml(p, id_to_index={i: n for n, i in enumerate(db_ids)})
    assert back == {k: v[:16] for k, v in table.items()}
    text = open(p).read()
    assert text.startswith("'0':\n- '1000'\n")  # same scalar style as data_splits/retrieval/*/*.yaml


def test_reference_yaml_tables_parse_when_present():
    from ralf_b200 import data as D

    p = "/root/reference/data_splits/retrieval/pku/val.yaml"
    if not os.path.exists(p):
        pytest.skip("reference tree not mounted")
    t = D.load_retrieval_yaml(p)
    assert len(t) > 100 and all(len(v) == 16 and all(isinstance(x, str) for x in v) for v in t.values())


def test_collate_main_matches_padding_rules():
    from ralf_b200 import data as D

    col = D.RetrievalCollator(layouts=None, max_seq_length=5, top_k=2, table_idx={})
    ex = [{"id": "7", "label": [1, 2], "center_x": [0.1, 0.2], "center_y": [0.3, 0.4], "width": [0.5, 0.6], "height": [0.7, 0.8]},
          {"id": "9", "label": [], "center_x": [], "center_y": [], "width": [], "height": []}]
    out = col.collate_main(ex)
    assert out["label"].tolist() == [[1, 2, 0, 0, 0], [0, 0, 0, 0, 0]]
    assert out["mask"].tolist() == [[True, True, False, False, False], [True, False, False, False, False]]
    assert out["width"][1].tolist() == pytest.approx([0.05, 0, 0, 0, 0]) and out["id"] == ["7", "9"]
    assert out["center_x"].dtype == torch.float32 and out["label"].dtype == torch.int64


def test_layout_table_from_rows_packs_valid_first():
    from ralf_b200 import data as D

    rows = [{"id": 3, "label": [2, 0, 1], "center_x": [.1, .2, .3], "center_y": [.4, .5, .6], "width": [.7, .8, .9], "height": [.15, .25, .35]},
            {"id": 4, "label": [1], "center_x": [.5], "center_y": [.5], "width": [.2], "height": [.2]}]
    t = D.LayoutTable.from_rows(rows, max_seq_length=2)
    assert tuple(t.packed.shape) == (2, 6, 2) and t.ids == [3, 4]
    assert t.packed[0, 0].tolist() == [2.0, 0.0] and t.packed[0, 1].tolist() == [1.0, 1.0]
    assert t.packed[1, 1].tolist() == [1.0, 0.0] and t.packed[1, 4].tolist() == pytest.approx([0.2, 0.0])


def test_multistep_lr_matches_torch_scheduler():
    from ralf_b200.checkpoint import multistep_lr

    for epochs, ms in ((50, (0.7,)), (10, (0.5, 0.9)), (20, (3, 15))):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=1e-4)
        mil = [int(m * epochs) if isinstance(m, float) else m for m in ms]
        sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=mil, gamma=0.1)
        for epoch in range(1, epochs + 1):
            assert multistep_lr(1e-4, epoch, epochs, ms) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12)
            opt.step()
            sch.step()


def test_model_files_use_reference_names(tmp_path):
    from ralf_b200 import checkpoint as C
    from ralf_b200 import generator as G

    m = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10)
    path = C.save_model(m, str(tmp_path), best_or_final="final", prefix="gen")
    assert os.path.basename(path) == "gen_final_model.pt"
    sd = torch.load(path, map_location="cpu")          # inference.py:319 reads it exactly like this
    assert list(sd.keys()) == list(m.state_dict().keys())
    m2 = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10)
    C.load_model(m2, str(tmp_path), "cpu", best_or_final="final", prefix="gen")
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


@pytest.mark.parametrize("task", ["uncond", "c", "cwh", "partial", "refinement"])
def test_task_preprocessor_reference_properties(task):
    """check_get_condition / check_output of the reference's test_task_preprocessor.py:28-58 over random batches."""
    import copy

    from oracle import synth
    from ralf_b200 import task as T

    tok = helpers.make_tokenizer()
    pre = T.TaskPreprocessor(tok, task)
    for seed in range(12):
        batch = synth.synth_batch(5, 8, 8, 10, 1, 4, seed=seed)
        torch.manual_seed(seed)
        cond, _ = T.get_condition(copy.deepcopy(batch), task, tok)
        if task != "uncond":
            assert -1 not in cond.seq[cond.mask].tolist()
            assert len(set(cond.seq[~cond.mask].tolist())) <= 1
        out = pre(cond)
        seq, pad_mask = out["seq"], out["pad_mask"]
        assert seq.min() >= 0 and seq.max() < pre.N_total
        assert set(seq[pad_mask].tolist()) <= {pre.name_to_id("pad")}
        assert pre.name_to_id("pad") not in seq[~pad_mask].tolist()
        assert (seq[:, 0] == pre.name_to_id("bos")).all() and (seq[:, 1] == pre.name_to_id(pre.TASK)).all()
        assert ((seq == pre.name_to_id("eos")).sum(dim=1) == 1).all()


def test_instance_transforms_reference_kats():
    """Known-answer tests of the reference: tests/train/helpers/test_hfds_instance_wise_transforms.py:7-37."""
    from ralf_b200 import data as D

    assert D.sort_label_transform({"label": [1, 2, 0], "data": ["a", "b", "c"]}) == {"label": [0, 1, 2], "data": ["c", "a", "b"]}
    assert D.sort_label_transform({"label": [1, 0], "data": [[0, 1], [2, 3]]}) == {"label": [0, 1], "data": [[2, 3], [0, 1]]}
    inputs = {"center_x": [0.5, 0.5, 0.4], "center_y": [0.5, 0.3, 0.3], "width": [1.0, 0.5, 0.5], "height": [0.8, 0.6, 0.6]}
    assert D.lexicographic_order(inputs) == [2, 1, 0]
    ex = {"id": "x", "label": [2, 0, 2], "center_x": [0.5, 0.5, 0.2], "center_y": [0.5, 0.5, 0.5], "width": [0.2, 0.2, 0.2],
          "height": [0.2, 0.2, 0.6], "retrieved": [1, 2, 3]}
    out = D.apply_transforms(dict(ex), ("image", "sort_label", "sort_lexicographic"))
    assert out["id"] == "x" and out["retrieved"] == [1, 2, 3]
    assert out["center_x"] == [0.2, 0.5, 0.5] and out["label"] == [2, 0, 2]  # the second sort wins, like in the reference
    assert D.shuffle_transform({"label": []}) == {"label": []}
    col = D.RetrievalCollator(layouts=None, max_seq_length=4, top_k=1, table_idx={}, transforms=("sort_label",))
    got = col.collate_main([{"id": "1", "label": [3, 1], "center_x": [.1, .2], "center_y": [.3, .4], "width": [.5, .6], "height": [.7, .8]}])
    assert got["label"][0].tolist() == [1, 3, 0, 0] and got["center_x"][0].tolist() == pytest.approx([.2, .1, 0, 0])


def test_linear_bucketizer_cycle_property():
    """Reference tests/train/helpers/test_bucketizer.py:26-38: decode(encode(x)) within half a bin, ids idempotent."""
    from ralf_b200.tokenizer import LinearBucketizer

    g = torch.Generator().manual_seed(0)
    for _ in range(100):
        n_data = int(torch.randint(1, 11, (1,), generator=g))
        n_boundaries = 2 ** int(torch.randint(1, 9, (1,), generator=g))
        x = torch.rand((n_data, 4), generator=g)
        b = LinearBucketizer(n_boundaries)
        ids = b.encode(x)
        x_cycle = b.decode(ids)
        assert (torch.abs(x - x_cycle) <= 1 / (2 * n_boundaries) + 1e-7).all()
        assert (ids == b.encode(x_cycle)).all()


@pytest.mark.parametrize("trial", [0, 1, 2, 3])
def test_collate_main_matches_reference_collate_fn(trial):
    """Ragged rows (empty layouts get the reference's dummy element): values AND dtypes of the reference's collate_fn
    (data.py:42-117; fixture tests/golden/collate_cases.npz, rows rebuilt here by the generator's recipe)."""
    import random

    import numpy as np

    from ralf_b200 import data as D

    rnd = random.Random(trial)
    exs = []
    for b in range(rnd.randint(1, 6)):
        n = rnd.choice([0, 1, 3, 10])
        g = torch.Generator().manual_seed(trial * 100 + b)
        exs.append({"id": str(trial * 10 + b), "label": [rnd.randint(0, 3) for _ in range(n)],
                    **{k: [rnd.random() for _ in range(n)] for k in ["center_x", "center_y", "width", "height"]},
                    "image": torch.rand((3, 4, 4), generator=g), "saliency": torch.rand((1, 4, 4), generator=g)})
    z = np.load(helpers.GOLDEN + "/collate_cases.npz")
    out = D.RetrievalCollator(layouts=None, max_seq_length=10, top_k=1, table_idx={}).collate_main(exs)
    for k in ["label", "mask", "center_x", "center_y", "width", "height", "image", "saliency"]:
        ref = z[f"{trial}_{k}"]
        assert out[k].numpy().dtype == ref.dtype, k
        np.testing.assert_array_equal(out[k].numpy(), ref)
    assert out["id"] == z[f"{trial}_id"].tolist()


def test_random_retrieval_draws_like_the_reference_wrapper():
    """RandomRetrievalDatasetWrapper.__getitem__ (helpers/random_retrieval_dataset_wrapper.py:70-72): one
    torch.randint(0, len(split), [top_k]) per sample, in order."""
    from ralf_b200 import data as D

    table = object.__new__(D.LayoutTable)
    table.packed = torch.zeros(500, 6, 10)
    col = D.RetrievalCollator(table, max_seq_length=10, top_k=16, random_retrieval=True, num_query_rows=321)
    torch.manual_seed(7)
    got = col.indices(["a", "b", "c"])
    torch.manual_seed(7)
    want = torch.stack([torch.randint(low=0, high=321, size=[16]) for _ in range(3)])
    assert torch.equal(got, want) and got.dtype == torch.int64 and int(got.max()) < 321
