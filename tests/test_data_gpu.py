"""GPU: the input pipeline / formats / resume rows end to end on the device."""
import copy

import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu


def _model(dev, seed=1):
    from ralf_b200 import generator as G

    m = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10, top_k=16)
    m.load_state_dict(helpers.synth_weights("ralf_cgl", seed), strict=True)
    return m.to(dev)


def test_gpu_collator_reproduces_reference_batch_and_tokens(cuda_device):
    """Gallery = the exemplars of the golden batch laid out as database rows; the collator's device gather must rebuild
    retrieved{...} exactly, and generation from the collated batch must give the reference's golden token ids."""
    from ralf_b200 import data as D
    from ralf_b200 import task as T

    z, meta = helpers.load_golden("ralf_cgl_256")
    batch = helpers.synth_batch(meta)
    B, K, E = meta["B"], meta["K"], meta["E"]
    r = batch["retrieved"]
    rows, table_idx = [], {}
    rng = np.random.default_rng(0)
    order = rng.permutation(B * K)            # scatter the exemplars over the database
    pos = {int(src): n for n, src in enumerate(order)}
    for src in order:
        b, k = divmod(int(src), K)
        n = int(r["mask"][b, k].sum())
        rows.append({"id": f"db{src}", **{key: r[key][b, k, :n].tolist() for key in ["label", "center_x", "center_y", "width", "height"]}})
    for b in range(B):
        table_idx[str(b)] = [pos[b * K + k] for k in range(K)] + [0] * 16   # 32-wide cache rows, cut to top_k
    layouts = D.LayoutTable.from_rows(rows, E, device=cuda_device)
    col = D.RetrievalCollator(layouts, E, top_k=K, table_idx=table_idx)
    examples = []
    for b in range(B):
        n = int(batch["mask"][b].sum())
        examples.append({"id": batch["id"][b], "image": batch["image"][b], "saliency": batch["saliency"][b],
                         **{key: batch[key][b, :n].tolist() for key in ["label", "center_x", "center_y", "width", "height"]}})
    out = col(examples)
    for key in ["label", "mask", "center_x", "center_y", "width", "height"]:
        assert torch.equal(out[key], batch[key]), key
        assert torch.equal(out["retrieved"][key].cpu(), r[key]), key
    model = _model(cuda_device, meta["seed"]).eval()
    cond, _ = T.get_condition(out, "uncond", model.tokenizer)
    res = model.sample(cond=cond.to(cuda_device), sampling_cfg={"name": "deterministic"}, cond_type="uncond", return_seq=True)
    np.testing.assert_array_equal(res["seq"].numpy(), z["gen_seq"])


def test_build_cache_table_drops_self_on_train_split(cuda_device, tmp_path):
    from ralf_b200 import data as D
    from ralf_b200.retrieval import GpuRetriever

    g = torch.Generator().manual_seed(2)
    N, d, k = 6000, 512, 32
    emb = torch.nn.functional.normalize(torch.randn((N, d), generator=g), dim=1)
    ret = GpuRetriever(emb, device=cuda_device)
    qids = list(range(100, 164))
    queries = emb[qids]
    train = D.build_cache_table(ret, queries, qids, "train", top_k=k, batch=40)
    val = D.build_cache_table(ret, queries, qids, "val", top_k=k, batch=40)
    scores = queries.double() @ emb.double().T
    want = torch.topk(scores, k + 1, dim=1).indices
    for n, qid in enumerate(qids):
        assert val[qid] == want[n].tolist() and val[qid][0] == qid      # the query is its own best match
        assert train[qid] == want[n, 1:].tolist() and len(train[qid]) == k
    p = D.cache_table_path("cgl", "train", "dreamsim", k, root=str(tmp_path))
    D.save_cache_table(train, p)
    assert D.load_cache_table(p, 16) == {q: v[:16] for q, v in train.items()}


def test_online_retrieval_in_collator(cuda_device):
    from oracle import synth
    from ralf_b200 import data as D
    from ralf_b200.retrieval import GpuRetriever

    g = torch.Generator().manual_seed(3)
    N, E, K = 3000, 10, 16
    lay = synth.synth_batch(N, 1, 1, E, 1, 4, seed=5)
    layouts = D.LayoutTable.from_tensors({k: lay[k] for k in ["label", "mask", "center_x", "center_y", "width", "height"]}, cuda_device)
    emb = torch.nn.functional.normalize(torch.randn((N, 512), generator=g), dim=1)
    col = D.RetrievalCollator(layouts, E, top_k=K, retriever=GpuRetriever(emb, device=cuda_device))
    q = torch.nn.functional.normalize(torch.randn((4, 512), generator=g), dim=1)
    ex = [{"id": str(i), "label": [1], "center_x": [.5], "center_y": [.5], "width": [.1], "height": [.1]} for i in range(4)]
    out = col(ex, query_embeddings=q.to(cuda_device))
    want = torch.topk(q.double() @ emb.double().T, K, dim=1).indices
    assert torch.equal(out["retrieved"]["index"].cpu(), want)
    for key in ["label", "mask", "center_x"]:
        assert torch.equal(out["retrieved"][key].cpu(), lay[key][want]), key
    assert out["retrieved"]["image"].shape == (4, K, 4, 1, 1)


def test_checkpoint_resume_continues_identically(cuda_device, tmp_path):
    from oracle import synth
    from ralf_b200 import checkpoint as C
    from ralf_b200.train import TrainEngine

    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=16)
    ma = _model(cuda_device, 27).train()
    inputs, targets = ma.preprocess(batch)
    ta = TrainEngine(ma, lr=1e-3, seed=5)
    for _ in range(2):
        ta.train_step(inputs, targets)
    C.save_model(ma, str(tmp_path), "epoch1", prefix="gen")
    C.save_train_state(ta, str(tmp_path / "gen_epoch1_train_state.pt"), epoch=1, best_val_loss=1.5)
    la = float(ta.train_step(inputs, targets))
    mb = _model(cuda_device, 99).train()                     # different weights until the checkpoint is loaded
    C.load_model(mb, str(tmp_path), cuda_device, "epoch1", prefix="gen")
    tb = TrainEngine(mb, lr=1e-3, seed=0)
    info = C.load_train_state(tb, str(tmp_path / "gen_epoch1_train_state.pt"))
    assert info == {"epoch": 1, "best_val_loss": 1.5, "step_count": 2} and tb.seed == ta.seed
    lb = float(tb.train_step(inputs, targets))
    assert abs(la - lb) <= 1e-5 * abs(la), (la, lb)
    assert (ta.ps.flat_p - tb.ps.flat_p).abs().max().item() <= 1e-5 * ta.ps.flat_p.abs().max().item()


def test_evaluate_is_mean_of_batch_losses(cuda_device):
    from oracle import synth
    from ralf_b200 import checkpoint as C

    model = _model(cuda_device, 28)
    batches = [synth.synth_batch(2, 128, 128, 10, 16, 4, seed=40 + i) for i in range(3)]
    model.eval()
    singles = []
    with torch.no_grad():
        for b in batches:
            i, t = model.preprocess(copy.deepcopy(b))
            i = {k: (v.to(cuda_device) if torch.is_tensor(v) else v) for k, v in i.items()}
            singles.append(float(model.train_loss(i, {k: v.to(cuda_device) for k, v in t.items()}, test=True)[1]["nll_loss"]))
    out = C.evaluate(model, [copy.deepcopy(b) for b in batches])
    assert abs(out["nll_loss"] - sum(singles) / 3) < 1e-6 and abs(out["total"] - out["nll_loss"]) < 1e-9
    half = C.evaluate(model, [copy.deepcopy(b) for b in batches], rank=1, world_size=2)   # shard 1 of 2, no group: local mean
    assert abs(half["nll_loss"] - singles[1]) < 1e-6
