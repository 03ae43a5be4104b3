"""Import the UNMODIFIED reference classes from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) validate the restatement in
``oracle/ralf_oracle.py`` and (b) generate the golden fixtures under ``tests/golden/``
(see ``tests/golden/make_golden.py``).  It only works in the build container, where
``/root/reference`` is mounted; nothing on the GPU box may import it.

The reference needs a handful of packages that are not installed here (omegaconf, timm,
hydra, seaborn, faiss ...).  None of them is on the hot path's arithmetic, so they are
replaced by tiny stub modules (recipe: SURVEY.md Appendix B).  ``timm.create_model`` is mapped
to torchvision's resnet50, which has the same graph and state-dict keys.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "image2layout"))


class DictConfig(dict):
    """Attribute-access dict standing in for omegaconf.DictConfig."""

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _install_stubs() -> None:
    import torchvision

    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class OmegaConf:
            set_struct = staticmethod(lambda *a, **k: None)
            to_container = staticmethod(lambda x, **k: dict(x))
            create = staticmethod(lambda x=None: DictConfig(x or {}))

        @contextlib.contextmanager
        def open_dict(x):
            yield x

        oc.DictConfig, oc.OmegaConf, oc.open_dict = DictConfig, OmegaConf, open_dict
        sys.modules["omegaconf"] = oc
    if "timm" not in sys.modules:
        tm = types.ModuleType("timm")
        tm.create_model = lambda n, **k: getattr(torchvision.models, n)(weights=None)
        sys.modules["timm"] = tm
    for n in ["seaborn", "hydra", "prdc", "pytorch_fid", "matplotlib", "matplotlib.pyplot"]:
        try:
            __import__(n)
        except Exception:
            sys.modules[n] = types.ModuleType(n)
    import datasets  # noqa: F401  (must be imported before the faiss stub)

    sys.modules.setdefault("faiss", types.ModuleType("faiss"))


def bootstrap(workdir: str) -> None:
    """Make ``image2layout`` importable and chdir to a scratch dir holding the weight files
    the reference constructors insist on reading (random, seeded; overwritten later by
    ``load_state_dict`` with the synthetic weights the tests use)."""
    assert available(), "reference tree not mounted"
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.makedirs(workdir, exist_ok=True)
    os.chdir(workdir)
    import torchvision

    from image2layout.train.fid.model import FIDNetV3

    p = os.path.join(workdir, "cache/PRECOMPUTED_WEIGHT_DIR/resnet50_a1_0-14fe96d1.pth")
    if not os.path.exists(p):
        os.makedirs(os.path.dirname(p), exist_ok=True)
        torch.manual_seed(0)
        torch.save(torchvision.models.resnet50(weights=None).state_dict(), p)
    for ds_name, ncls in [("cgl", 4), ("pku10", 3)]:
        p = os.path.join(workdir, f"tmp/fidnet/{ds_name}/model_best.pth.tar")
        if not os.path.exists(p):
            os.makedirs(os.path.dirname(p), exist_ok=True)
            torch.manual_seed(0)
            torch.save({"state_dict": FIDNetV3(num_label=ncls, max_bbox=10).state_dict()}, p)


LABELS = {"cgl": ["logo", "text", "underlay", "embellishment"], "pku": ["text", "logo", "underlay"]}


def make_tokenizer(dataset_name: str = "cgl", max_seq_length: int = 10):
    import datasets as ds
    from image2layout.train.helpers.layout_tokenizer import LayoutSequenceTokenizer

    label_feature = ds.ClassLabel(names=LABELS[dataset_name])
    tok = LayoutSequenceTokenizer(
        label_feature=label_feature,
        max_seq_length=max_seq_length,
        num_bin=128,
        var_order=["label", "width", "height", "center_x", "center_y"],
        pad_until_max=False,
        special_tokens=["pad", "bos", "eos"],
        is_loc_vocab_shared=False,
        geo_quantization="linear",
    )
    features = ds.Features({"label": ds.Sequence(label_feature)})
    return tok, features


def make_ralf(dataset_name: str = "cgl", max_seq_length: int = 10):
    """The shipped RALF class (retrieval_augmented_autoreg.py:998-1033)."""
    from image2layout.train.models.retrieval_augmented_autoreg import (
        ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg as RALF,
    )

    tok, features = make_tokenizer(dataset_name, max_seq_length)
    model = RALF(
        features=features,
        tokenizer=tok,
        dataset_name=dataset_name,
        max_seq_length=max_seq_length,
        db_dataset=None,
        retrieval_backbone="dreamsim",
        random_retrieval=False,
        top_k=16,
        saliency_k="None",
        auxilary_task="uncond",
    )
    return model.eval(), tok, features


def make_autoreg(dataset_name: str = "cgl", max_seq_length: int = 10):
    """The Autoreg baseline (autoreg.py:590-622), BASELINE config 1."""
    from image2layout.train.models.autoreg import ConcateAuxilaryTaskAutoreg

    tok, features = make_tokenizer(dataset_name, max_seq_length)
    model = ConcateAuxilaryTaskAutoreg(
        features=features, tokenizer=tok, auxilary_task="uncond"
    )
    return model.eval(), tok, features
