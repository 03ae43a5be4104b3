"""Shared test helpers: golden fixtures, synthetic weights/batches (test infrastructure)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def load_schema(name):
    with open(os.path.join(GOLDEN, f"schema_{name}.json")) as f:
        return json.load(f)


def synth_weights(schema_name, seed):
    from oracle import synth

    return synth.synth_state_dict(load_schema(schema_name), seed=seed)


def synth_batch(meta, num_labels=4):
    from oracle import synth

    return synth.synth_batch(meta["B"], meta["H"], meta["W"], meta["E"], meta["K"], num_labels, seed=meta["seed"])


def make_tokenizer(dataset="cgl", max_seq_length=10):
    from ralf_b200.tokenizer import LayoutSequenceTokenizer

    names = {"cgl": ["logo", "text", "underlay", "embellishment"], "pku": ["text", "logo", "underlay"]}[dataset]
    return LayoutSequenceTokenizer(names, max_seq_length)


def image4(batch):
    return torch.cat([batch["image"], batch["saliency"]], dim=1)
