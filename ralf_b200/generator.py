"""Drop-in model classes: the reference's Hydra ``_target_`` surface for the RALF hot path.

``image2layout.train.models.generator`` re-exports the classes Hydra instantiates
(train/models/generator.py:1-9; config/generator/ralf.yaml:1-7, autoreg.yaml:1-2).  The two classes here
keep that contract -- constructor kwargs, ``state_dict`` key names/shapes (strict ``load_state_dict`` of a
reference checkpoint), and the methods ``train.py`` / ``inference.py`` call -- while every
tensor operation runs in :class:`ralf_b200.engine.Engine` (hand-written sm_100a kernels, no CPU fallback).

Scope: unconstrained generation (``uncond``) forward / loss / greedy sampling.  The
constrained tasks (c, cwh, partial, refinement, relation) and stochastic sampling are row f3 ("next") and
raise NotImplementedError rather than silently taking another path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import torch
import torch.nn as nn
from torch import Tensor

from .engine import Engine
from .tokenizer import LayoutSequenceTokenizer

TASK_TOKENS = ["end_of_task", "label", "label_size", "relationship", "refinement", "completion", "uncondition"]
PREPROCESS_SPECIAL = ["sep", "relation_sep", "canvas"]
N_REL_LOC, N_REL_SIZE = 6, 4  # helpers/relationships.py:11-24


# ------------------------------------------------------------------------------------------------
# conditional inputs (models/common/base_model.py:17-109), uncond subset
# ------------------------------------------------------------------------------------------------
@dataclass
class ConditionalInputs:
    image: Tensor
    id: Any = None
    task: Optional[str] = None
    seq: Optional[Tensor] = None
    mask: Optional[Tensor] = None
    retrieved: dict = field(default_factory=dict)

    def to(self, x: Any) -> "ConditionalInputs":
        self.image = self.image.to(x)
        self.retrieved = {k: (v.to(x) if torch.is_tensor(v) else v) for k, v in self.retrieved.items()}
        return self


def get_condition(batch: dict, cond_type: Optional[str], tokenizer: LayoutSequenceTokenizer):
    """helpers/task.py:45-183 for cond_type in (None, "none", "uncond")."""
    if cond_type not in (None, "none", "uncond"):
        raise NotImplementedError(f"cond_type={cond_type!r}: constrained tasks are SURVEY.md 8(f3), not built yet")
    image = batch["image"] if batch["image"].size(1) == 4 else torch.cat([batch["image"], batch["saliency"]], dim=1)
    try:
        ids = torch.tensor(list(map(int, batch["id"])), dtype=torch.long)
    except Exception:
        ids = batch.get("id")
    retrieved = batch.get("retrieved", {})
    if isinstance(retrieved, list):
        assert len(retrieved) == 1
        retrieved = retrieved[0]
    return ConditionalInputs(image=image, id=ids, task=cond_type, retrieved=retrieved), batch


class UnconditionalPreprocessor:
    """layoutformerpp/task_preprocessor.py:58-140,354-384: constraint sequence [bos, uncondition, end_of_task, eos]."""

    def __init__(self, tokenizer: LayoutSequenceTokenizer) -> None:
        self.tokenizer = tokenizer
        self.tokens = TASK_TOKENS + PREPROCESS_SPECIAL + [f"rel_elem_{i}" for i in range(tokenizer.max_seq_length)] + \
            [f"rel_loc_{i}" for i in range(N_REL_LOC)] + [f"rel_size_{i}" for i in range(N_REL_SIZE)]

    @property
    def N_total(self) -> int:
        return self.tokenizer.N_total + len(self.tokens)

    def name_to_id(self, name: str) -> int:
        if name in self.tokenizer.special_tokens:
            return self.tokenizer.name_to_id(name)
        return self.tokens.index(name) + self.tokenizer.N_total

    def __call__(self, cond) -> dict:
        B = cond.image.size(0)
        ids = [self.name_to_id(n) for n in ("bos", "uncondition", "end_of_task", "eos")]
        seq = torch.tensor(ids, dtype=torch.long, device=cond.image.device)[None].expand(B, -1).contiguous()
        return {"seq": seq, "pad_mask": seq == self.tokenizer.name_to_id("pad")}


# ------------------------------------------------------------------------------------------------
# parameter schema (state-dict contract)
# ------------------------------------------------------------------------------------------------
def _mha(p):
    return [(p + ".in_proj_weight", (768, 256)), (p + ".in_proj_bias", (768,)),
            (p + ".out_proj.weight", (256, 256)), (p + ".out_proj.bias", (256,))]


def _ffn_norms(p, ff, norms):
    out = [(p + ".linear1.weight", (ff, 256)), (p + ".linear1.bias", (ff,)),
           (p + ".linear2.weight", (256, ff)), (p + ".linear2.bias", (256,))]
    for n in norms:
        out += [(f"{p}.{n}.weight", (256,)), (f"{p}.{n}.bias", (256,))]
    return out


def _enc_layers(p, n, ff=1024):
    out = []
    for i in range(n):
        out += _mha(f"{p}.{i}.self_attn") + _ffn_norms(f"{p}.{i}", ff, ["norm1", "norm2"])
    return out


def _bn(p, c):
    return [(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,), "buffer"),
            (p + ".running_var", (c,), "buffer"), (p + ".num_batches_tracked", (), "long")]


def _feed_forward(p, hidden=1024):
    return [(p + ".net.0.weight", (256,)), (p + ".net.0.bias", (256,)), (p + ".net.1.weight", (hidden, 256)),
            (p + ".net.1.bias", (hidden,)), (p + ".net.4.weight", (256, hidden)), (p + ".net.4.bias", (256,))]


def param_schema(num_labels: int, vocab: int, const_vocab: int, is_ralf: bool) -> list[tuple]:
    """(name, shape[, kind]) for every state-dict entry of the reference class, in the reference's order.
    kind: param (default) | frozen | buffer | long | bool."""
    s: list[tuple] = [("flag_img", (1,), "long"), ("flag_user_const", (1,), "long")]
    b = "encoder.extractor.body"
    s += [(b + ".conv1.weight", (64, 4, 7, 7))] + _bn(b + ".bn1", 64)
    inpl = 64
    for li, (nblk, planes) in enumerate([(3, 64), (4, 128), (6, 256), (3, 512)], start=1):
        for bi in range(nblk):
            p = f"{b}.layer{li}.{bi}"
            s += [(p + ".conv1.weight", (planes, inpl, 1, 1))] + _bn(p + ".bn1", planes)
            s += [(p + ".conv2.weight", (planes, planes, 3, 3))] + _bn(p + ".bn2", planes)
            s += [(p + ".conv3.weight", (planes * 4, planes, 1, 1))] + _bn(p + ".bn3", planes * 4)
            if bi == 0:
                s += [(p + ".downsample.0.weight", (planes * 4, inpl, 1, 1))] + _bn(p + ".downsample.1", planes * 4)
            inpl = planes * 4
    e = "encoder.extractor"
    s += [(e + ".fpn_conv11_4.weight", (256, 1024, 1, 1)), (e + ".fpn_conv11_4.bias", (256,)),
          (e + ".fpn_conv11_5.weight", (256, 2048, 1, 1)), (e + ".fpn_conv11_5.bias", (256,)),
          (e + ".fpn_conv33.weight", (256, 256, 3, 3)), (e + ".fpn_conv33.bias", (256,)),
          (e + ".proj.weight", (256, 512, 1, 1)), (e + ".proj.bias", (256,))]
    s += _enc_layers("transformer_encoder.layers", 6)
    for i in range(6):
        p = f"decoder.transformer.layers.{i}"
        s += _mha(p + ".self_attn") + _mha(p + ".multihead_attn") + _ffn_norms(p, 1024, ["norm1", "norm2", "norm3"])
    s += [("decoder.emb.weight", (vocab, 256)), ("decoder.pos_emb.pe", (1, 5000, 256), "pe"),
          ("decoder.head.0.weight", (256,)), ("decoder.head.0.bias", (256,)), ("decoder.head.1.weight", (vocab, 256))]
    uc = [(n, sh) for (n, sh) in _enc_layers("user_const_encoder.encoder.layers", 6)] + \
        [("user_const_encoder.emb.weight", (const_vocab, 256)), ("user_const_encoder.pos_emb.pe", (1, 5000, 256), "pe")]
    if is_ralf:
        f = "layout_encoer"
        fid = [(f + ".emb_label.weight", (num_labels, 256)), (f + ".fc_bbox.weight", (256, 4)),
               (f + ".fc_bbox.bias", (256,)), (f + ".enc_fc_in.weight", (256, 512)), (f + ".enc_fc_in.bias", (256,)),
               (f + ".enc_transformer.token", (1, 1, 256)), (f + ".enc_transformer.token_mask", (1, 1), "bool")]
        fid += _enc_layers(f + ".enc_transformer.core.layers", 4, ff=128)
        fid += [(f + ".dec_fc_in.weight", (256, 512)), (f + ".dec_fc_in.bias", (256,))]
        s += [(t[0], t[1], t[2] if len(t) > 2 else "frozen") for t in fid]  # freeze_layout_encoder (:150-154)
        s += [("pos_emb_1d.pe", (1, 5000, 256), "pe")] + _feed_forward("layout_adapter") + _feed_forward("head")
        s += uc + [("task_emb.weight", (2, 1))]
        s += [("attn.norm.weight", (256,)), ("attn.norm.bias", (256,)), ("attn.to_q.weight", (512, 256)),
              ("attn.to_kv.weight", (1024, 256)), ("attn.to_out.0.weight", (256, 512)), ("attn.to_out.0.bias", (256,))]
    else:
        s += uc + [("task_emb.weight", (2, 1))]
    return s


class _Node(nn.Module):
    """Bare container so dotted reference key names map onto a module tree."""


def _register(root: nn.Module, name: str, tensor: Tensor, kind: str) -> None:
    parts = name.split(".")
    node = root
    for p in parts[:-1]:
        if not hasattr(node, p):
            node.add_module(p, _Node())
        node = getattr(node, p)
    if kind in ("param", "frozen"):
        node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=(kind == "param")))
    else:
        node.register_buffer(parts[-1], tensor)


def _sine_pe(max_len=5000, d=256):
    from .engine import _sine_pe_1d

    return _sine_pe_1d(max_len, d)[None]


# ------------------------------------------------------------------------------------------------
# model classes
# ------------------------------------------------------------------------------------------------
class _B200LayoutModel(nn.Module):
    IS_RALF = True

    def __init__(self, features: Any, tokenizer: Any, dataset_name: str = "cgl", max_seq_length: int = 10,
                 db_dataset: Any = None, d_model: int = 256, top_k: int = 16, retrieval_backbone: str = "saliency",
                 random_retrieval: bool = False, saliency_k: Any = 8, auxilary_task: Optional[str] = "uncond",
                 use_multitask: bool = False, use_flag_embedding: bool = True, precision: str = "bf16x3",
                 **kwargs: Any) -> None:
        super().__init__()
        if d_model != 256:
            raise NotImplementedError("the B200 kernels are specialised for d_model = 256 (reference default)")
        if auxilary_task not in (None, "uncond"):
            raise NotImplementedError(f"auxilary_task={auxilary_task!r}: constrained tasks are SURVEY.md 8(f3)")
        self.features = features
        self.tokenizer = self._host_tokenizer(tokenizer)
        self.dataset_name = dataset_name
        self.max_seq_length = max_seq_length
        self.d_model = d_model
        self.top_k = top_k
        self.retrieval_backbone = retrieval_backbone
        self.random_retrieval = random_retrieval
        self.saliency_k = saliency_k
        self.auxilary_task = auxilary_task or "uncond"
        self.use_multitask = use_multitask
        self.use_flag_embedding = use_flag_embedding
        self.precision = precision
        self.preprocessor = UnconditionalPreprocessor(self.tokenizer)
        g = torch.Generator().manual_seed(0)
        for entry in param_schema(self.tokenizer.N_label, self.tokenizer.N_total, self.preprocessor.N_total, self.IS_RALF):
            name, shape = entry[0], entry[1]
            kind = entry[2] if len(entry) > 2 else "param"
            if kind == "pe":
                t, kind = _sine_pe(), "buffer"
            elif kind == "long":
                t = torch.ones(shape, dtype=torch.long) if name == "flag_user_const" else torch.zeros(shape, dtype=torch.long)
            elif kind == "bool":
                t = torch.zeros(shape, dtype=torch.bool)
            elif name.endswith("running_var") or (len(shape) == 1 and name.endswith(".weight")):
                t = torch.ones(shape)
            elif len(shape) <= 1:
                t = torch.zeros(shape)
            else:  # placeholder init; real runs load a checkpoint (inference.py:319) or train from the reference init
                fan_in = 1
                for d_ in shape[1:]:
                    fan_in *= d_
                t = torch.randn(shape, generator=g) * (0.02 if "emb" in name else fan_in ** -0.5)
            _register(self, name, t, kind)
        self._engine: Optional[Engine] = None

    @staticmethod
    def _host_tokenizer(tok: Any) -> LayoutSequenceTokenizer:
        if isinstance(tok, LayoutSequenceTokenizer):
            return tok
        # a reference LayoutSequenceTokenizer: rebuild the host mirror from its public properties
        assert tok.geo_quantization == "linear" and not tok.is_loc_vocab_shared, "only the linear tokenizer is mirrored"
        return LayoutSequenceTokenizer(list(tok._label_feature.names), tok.max_seq_length, tok.N_bbox_per_var,
                                       list(tok.var_order), list(tok.special_tokens))

    # ---- nn.Module plumbing --------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._engine = None
        return out

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self) -> Engine:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("ralf_b200 runs on a B200 only: move the model to cuda (there is no CPU fallback)")
        if self._engine is None or self.training:
            # training mutates parameters every step; weights are re-prepared on demand
            self._engine = Engine(self.state_dict(), dev, is_ralf=self.IS_RALF, top_k=self.top_k,
                                  npass=3 if self.precision == "bf16x3" else 1)
        return self._engine

    @property
    def special_token_ids(self) -> dict:
        ids = {k: self.tokenizer.name_to_id(k) for k in self.tokenizer.special_tokens}
        ids.setdefault("mask", -1)
        return ids

    def compute_stats(self) -> None:
        n = sum(p.numel() for p in self.parameters()) / 1e6
        print(f"number of parameters: {n:.2f}M")

    def update_per_epoch(self, epoch: int, warmup: int, epochs: int) -> None:
        pass

    def aggregate_sampling_config(self, sampling_cfg, test_cfg=None):
        return sampling_cfg

    # ---- train.py / inference.py surface --------------------------------------------------------
    def preprocess(self, inputs: dict) -> tuple[dict, dict]:
        """retrieval_augmented_autoreg.py:764-785 -> :509-523 (autoreg.py equivalent)."""
        cond, inputs = get_condition(inputs, self.auxilary_task, self.tokenizer)
        const = self.preprocessor(cond)
        data = self.tokenizer.encode(inputs)
        image = torch.cat([inputs["image"], inputs["saliency"]], dim=1) if inputs["image"].size(1) != 4 else inputs["image"]
        out = {"seq": data["seq"][:, :-1], "tgt_key_padding_mask": ~data["mask"][:, :-1], "image": image,
               "seq_layout_const": const["seq"], "seq_layout_const_pad_mask": const["pad_mask"]}
        if self.IS_RALF:
            assert inputs["retrieved"]["image"].size(2) == 4, f"{inputs['retrieved']['image'].shape=}"
            out["retrieved"] = inputs["retrieved"]
        return out, {"seq": data["seq"][:, 1:]}

    def _encode(self, inputs: dict):
        return self.engine().encode(inputs["image"], inputs.get("retrieved"), inputs["seq_layout_const"],
                                    inputs["seq_layout_const_pad_mask"])

    @torch.no_grad()
    def _encode_into_memory(self, inputs: dict) -> dict:
        return {"memory": self._encode(inputs)[0]}

    @torch.no_grad()
    def forward(self, inputs: dict) -> dict:
        """-> {"logits": [B, S, V]} (retrieval_augmented_autoreg.py:190-207)."""
        mem, mem_s = self._encode(inputs)
        B, Mlen = mem.shape[0], mem.shape[1]
        logits = self.engine().decoder_logits(inputs["seq"], inputs["tgt_key_padding_mask"], mem_s, B, Mlen)
        return {"logits": logits}

    def train_loss(self, inputs: dict, targets: dict, test: bool = False):
        """CrossEntropyLoss(label_smoothing=0.1, ignore_index=pad) over b s c -> b c s (:209-216).
        Forward value only: the backward/optimizer kernels are SURVEY.md 8 rows a12/a13 (round 2)."""
        from . import ops

        outputs = self(inputs)
        loss = ops.ce_label_smooth(outputs["logits"], targets["seq"].to(outputs["logits"].device), 0.1,
                                   self.tokenizer.name_to_id("pad"))
        return outputs, {"nll_loss": loss}

    @torch.no_grad()
    def sample(self, cond: Any, batch_size: Optional[int] = None, sampling_cfg: Any = None,
               cond_type: Optional[str] = "uncond", return_violation: bool = False, use_backtrack: bool = True,
               return_decoded_cond: bool = False, return_seq: bool = False, **kwargs: Any):
        """Greedy generation (retrieval_augmented_autoreg.py:218-325) with KV caches on the GPU."""
        if cond_type not in (None, "none", "uncond"):
            raise NotImplementedError(f"cond_type={cond_type!r}: SURVEY.md 8(f3)")
        name = getattr(sampling_cfg, "name", None) if sampling_cfg is not None else "deterministic"
        if name not in (None, "deterministic"):
            raise NotImplementedError(f"sampling {name!r}: only greedy (helpers/sampling.py:24-25) is built")
        image = cond.image
        B = image.size(0)
        if B == 1 and batch_size and batch_size > 1:
            B = batch_size
            image = image.expand(B, -1, -1, -1)
        const = self.preprocessor(cond if cond.image.size(0) == B else ConditionalInputs(image=image))
        eng = self.engine()
        mem, mem_s = eng.encode(image, getattr(cond, "retrieved", None) if self.IS_RALF else None, const["seq"],
                                const["pad_mask"])
        ids = self.special_token_ids
        seq = eng.generate(mem_s, B, mem.shape[1], self.tokenizer.token_mask, ids["bos"], ids["pad"],
                           self.tokenizer.max_token_length)
        seq = seq.cpu()
        out = self.tokenizer.decode(seq)  # BaseModel.postprocess (base_model.py:367-389)
        if return_seq:
            out["seq"] = seq
        if not return_violation:
            return out
        return out, {"total": 1, "viorated": 0}  # violate.py:81-88 (uncond)


class ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg(_B200LayoutModel):
    """The shipped RALF architecture (retrieval_augmented_autoreg.py:998-1033)."""

    IS_RALF = True


class ConcateAuxilaryTaskAutoreg(_B200LayoutModel):
    """The Autoreg baseline (models/autoreg.py:590-622); BASELINE.json config 1."""

    IS_RALF = False


RALF = ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg
