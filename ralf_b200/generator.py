"""Drop-in model classes: the reference's Hydra ``_target_`` surface for the RALF hot path.

``image2layout.train.models.generator`` re-exports the classes Hydra instantiates
(train/models/generator.py:1-9; config/generator/ralf.yaml:1-7, autoreg.yaml:1-2).  The two classes here
keep that contract -- constructor kwargs, ``state_dict`` key names/shapes (strict ``load_state_dict`` of a
reference checkpoint), and the methods ``train.py`` / ``inference.py`` call -- while every
tensor operation runs in :class:`ralf_b200.engine.Engine` (hand-written sm_100a kernels, no CPU fallback).

Scope: forward / loss / sampling for the tasks uncond, c, cwh, partial, refinement (host side in
ralf_b200/task.py, decoding-space restriction + deterministic / random / top_k / top_p / gumbel sampling in the device
kernel ``ralf_sample_next``).  ``relation`` (Gen-R) decodes per canvas through ``engine.DecodeSession`` under the host-side
backtracking sampler of ``ralf_b200/relation.py``.  Anything else raises NotImplementedError rather than silently
taking another path.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import task as T
from .engine import Engine
from .task import ConditionalInputs, TaskPreprocessor, get_condition  # noqa: F401  (re-exported: reference import sites)
from .tokenizer import LayoutSequenceTokenizer

# model.sample()'s plain greedy decode loop is replayed from a per-shape CUDA graph (measured on B200: 2471 vs 1632 layouts/s
# through the drop-in API at batch 128, profiles/r2_ab_call1.md); RALF_SAMPLE_GRAPH=0 restores the eager loop for A/B runs.
_SAMPLE_GRAPH = __import__("os").environ.get("RALF_SAMPLE_GRAPH", "1") != "0"

UnconditionalPreprocessor = TaskPreprocessor  # task=None/"uncond" (task_preprocessor.py:354-384)



def _cfg_get(cfg: Any, key: str, default: Any = None) -> Any:
    """Hydra DictConfig / dataclass / plain dict access."""
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


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


def _initial_value(name: str, shape: tuple, shapes: dict, weight_init: bool = True) -> Tensor:
    """Initial value of one floating-point state-dict entry, with the distribution the reference constructor leaves it in
    (tests/golden/init_stats.json, dumped from freshly built reference models):
      * every matrix of the three transformers (image encoder, constraint encoder, decoder): Xavier-uniform
        (retrieval_augmented_autoreg.py:171-176, common/common.py:69-72,233-236);
      * decoder embedding + output projection, constraint embedding, task embedding: N(0, 0.02) (common/common.py:74-82,
        :225-226; retrieval_augmented_autoreg.py:694);
      * LayerNorm / BatchNorm: ones and zeros, running statistics 0 / 1; attention in/out projection biases: zeros;
      * every other Linear / Conv2d: PyTorch's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias.
    ``weight_init`` is the reference's constructor flag: on by default for the RALF class, off for the Autoreg baseline
    (models/autoreg.py:39,85-93), which therefore never runs ``init_weights``: there only
    the constraint encoder is Xavier / N(0, 0.02) (:478-488), attention input projections keep nn.MultiheadAttention's own
    Xavier default, and the decoder embedding stays nn.Embedding's N(0, 1).
    Drawn from the global torch generator, like the reference (train.py seeds it)."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "running_var":
        return torch.ones(shape)
    if leaf == "running_mean" or leaf == "in_proj_bias" or name.endswith("out_proj.bias"):
        return torch.zeros(shape)
    small_normal = ("user_const_encoder.emb.weight", "task_emb.weight") + \
        (("decoder.emb.weight", "decoder.head.1.weight") if weight_init else ())
    if name in small_normal:
        return torch.randn(shape) * 0.02
    if name == "decoder.emb.weight":
        return torch.randn(shape)
    if len(shape) == 1:
        if leaf == "weight":
            return torch.ones(shape)  # normalisation gains
        sibling = shapes.get(name[:-len("bias")] + "weight")
        if leaf != "bias" or sibling is None or len(sibling) == 1:
            return torch.zeros(shape)  # normalisation shifts
        fan_in = 1
        for d_ in sibling[1:]:
            fan_in *= d_
        return (torch.rand(shape) * 2 - 1) * fan_in ** -0.5
    fan_in = 1
    for d_ in shape[1:]:
        fan_in *= d_
    xavier = ("transformer_encoder.", "decoder.transformer.", "user_const_encoder.encoder.") if weight_init \
        else ("user_const_encoder.encoder.",)
    if len(shape) == 2 and (name.startswith(xavier) or leaf == "in_proj_weight"):
        bound = (6.0 / (fan_in + shape[0])) ** 0.5
    elif leaf == "weight" and name.endswith("emb_label.weight"):
        return torch.randn(shape)  # nn.Embedding default (frozen FIDNet; replaced by its checkpoint)
    else:
        bound = fan_in ** -0.5
    return (torch.rand(shape) * 2 - 1) * bound


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
                 relation_table: Any = None, pretrained: bool = True, **kwargs: Any) -> None:
        super().__init__()
        if d_model != 256:
            raise NotImplementedError("the B200 kernels are specialised for d_model = 256 (reference default)")
        # Constructor options of the reference (retrieval_augmented_autoreg.py:61-80,636-650) that change the architecture:
        # only their shipped values are built; anything else fails here instead of being silently ignored.
        shipped = {"use_reference_image": False, "layout_backbone": "feature_extractor", "freeze_layout_encoder": True,
                   "decoder_d_model": 256, "encoder_pos_emb": "sine", "decoder_pos_emb": "layout",
                   "global_task_embedding": False, "shared_embedding": False, "decoder_num_layers": 6, "RELATION_SIZE": 10}
        if not use_flag_embedding:
            raise NotImplementedError("use_flag_embedding=False: only the shipped configuration (True) is built")
        for key, value in shipped.items():
            if key in kwargs and kwargs[key] != value:
                raise NotImplementedError(f"{key}={kwargs[key]!r}: only the shipped configuration ({value!r}) is built")
        if saliency_k == "dynamic":
            raise NotImplementedError('saliency_k="dynamic" (hybrid retrieval embedding) is not built')
        unknown = set(kwargs) - set(shipped) - {"weight_init"}
        if unknown:
            raise TypeError(f"unexpected constructor arguments: {sorted(unknown)}")
        assert auxilary_task in T.COND_TYPES, f"{auxilary_task=} must be one of {T.COND_TYPES}"
        self.features = features
        self.tokenizer = self._host_tokenizer(tokenizer)
        self.dataset_name = dataset_name
        self.max_seq_length = max_seq_length
        self.d_model = d_model
        self.top_k = top_k
        self.retrieval_backbone = retrieval_backbone
        self.random_retrieval = random_retrieval
        self.saliency_k = saliency_k
        self.auxilary_task = auxilary_task
        self.use_multitask = use_multitask
        self.use_flag_embedding = use_flag_embedding
        self.precision = precision
        self.relation_table = relation_table  # dict / path; None -> the reference's cache locations, read on first use
        self.preprocessor = self._make_preprocessor(auxilary_task)
        schema = param_schema(self.tokenizer.N_label, self.tokenizer.N_total, self.preprocessor.N_total, self.IS_RALF)
        shapes = {entry[0]: entry[1] for entry in schema}
        for entry in schema:
            name, shape = entry[0], entry[1]
            kind = entry[2] if len(entry) > 2 else "param"
            if kind == "pe":
                t, kind = _sine_pe(), "buffer"
            elif kind == "long":
                t = torch.ones(shape, dtype=torch.long) if name == "flag_user_const" else torch.zeros(shape, dtype=torch.long)
            elif kind == "bool":
                t = torch.zeros(shape, dtype=torch.bool)
            else:
                t = _initial_value(name, shape, shapes, bool(kwargs.get("weight_init", self.IS_RALF)))
            _register(self, name, t, kind)
        if pretrained:
            self._load_pretrained_files()
        self._engine: Optional[Engine] = None
        self._train_engine = None

    @staticmethod
    def _host_tokenizer(tok: Any) -> LayoutSequenceTokenizer:
        if isinstance(tok, LayoutSequenceTokenizer):
            return tok
        # a reference LayoutSequenceTokenizer: rebuild the host mirror from its public properties
        assert tok.geo_quantization == "linear" and not tok.is_loc_vocab_shared, "only the linear tokenizer is mirrored"
        return LayoutSequenceTokenizer(list(tok._label_feature.names), tok.max_seq_length, tok.N_bbox_per_var,
                                       list(tok.var_order), list(tok.special_tokens))

    def _load_pretrained_files(self) -> None:
        """What the reference constructors read from disk, when the files are there (they are overwritten by a later
        ``load_state_dict`` of a trained checkpoint; a from-scratch training run starts from them):
        the ImageNet ResNet50 ``resnet50_a1_0-14fe96d1.pth`` (cwd, else ./cache/PRECOMPUTED_WEIGHT_DIR; common/image.py:36-48),
        whose stem gets a 4th input channel = the mean of the three (:70-78), and the frozen layout encoder
        ``tmp/fidnet/<dataset>/model_best.pth.tar`` (else ./cache/PRECOMPUTED_WEIGHT_DIR/fidnet/...; fid/model.py:131-169,
        "pku" -> "pku10").  The reference asserts that they exist; here a missing file leaves the initial values in place."""
        import logging
        import os

        log = logging.getLogger(__name__)
        own = dict(self.named_parameters())
        own.update(dict(self.named_buffers()))

        def first(*paths):
            return next((p for p in paths if os.path.exists(p)), None)

        def copy_in(prefix: str, sd: dict) -> int:
            n = 0
            with torch.no_grad():
                for k, v in sd.items():
                    dst = own.get(prefix + k)
                    if dst is not None and torch.is_tensor(v) and tuple(dst.shape) == tuple(v.shape):
                        dst.copy_(v)
                        n += 1
            return n

        weight_dir = os.path.join(".", "cache", "PRECOMPUTED_WEIGHT_DIR")
        path = first("resnet50_a1_0-14fe96d1.pth", os.path.join(weight_dir, "resnet50_a1_0-14fe96d1.pth"))
        if path is None:
            log.info("resnet50_a1_0-14fe96d1.pth not found: the ResNet50 trunk keeps its initial values")
        else:
            sd = dict(torch.load(path, map_location="cpu", weights_only=True))
            w = sd["conv1.weight"]
            sd["conv1.weight"] = torch.cat([w, w.mean(dim=1, keepdim=True)], dim=1)
            log.info("ResNet50 trunk: %d tensors from %s", copy_in("encoder.extractor.body.", sd), path)
        if self.IS_RALF:
            ds_name = "pku10" if self.dataset_name == "pku" else self.dataset_name
            path = first(os.path.join("tmp", "fidnet", ds_name, "model_best.pth.tar"),
                         os.path.join(weight_dir, "fidnet", ds_name, "model_best.pth.tar"))
            if path is None:
                log.info("FIDNetV3 checkpoint for %r not found: the layout encoder keeps its initial values", ds_name)
            else:
                sd = torch.load(path, map_location="cpu", weights_only=True)["state_dict"]
                log.info("FIDNetV3 layout encoder: %d tensors from %s", copy_in("layout_encoer.", sd), path)

    def _make_preprocessor(self, task: Optional[str]) -> TaskPreprocessor:
        """PREPROCESSOR[task](tokenizer=...) (task_preprocessor.py:593-602).  The relation task reads its relationship
        table from ``relation_table`` (dict or file) or from where the reference keeps it (task_preprocessor.py:498-506)."""
        if task != "relation":
            return TaskPreprocessor(self.tokenizer, task)
        import os

        from . import relation as R

        table = self.relation_table
        if table is None:
            for cand in (os.path.join("cache", R.REFERENCE_TABLE_NAME),
                         os.path.join("cache", "PRECOMPUTED_WEIGHT_DIR", "relationship", R.REFERENCE_TABLE_NAME)):
                if os.path.exists(cand):
                    table = cand
                    break
            else:
                raise FileNotFoundError(f"relation task: pass relation_table= or provide cache/{R.REFERENCE_TABLE_NAME}")
        if isinstance(table, str):  # read once: the multitask mixture rebuilds the preprocessor at every step (:745-750)
            table = self.relation_table = R.load_relation_table(table)
        return R.RelationPreprocessor(self.tokenizer, table)

    @property
    def top_k(self) -> int:
        return self._top_k

    @top_k.setter
    def top_k(self, k: int) -> None:
        """inference.py:346 re-assigns ``model.top_k`` between runs ("dynamic top-k"): the engines that already exist
        follow (the weights do not depend on k; retrieved inputs are cut to the first k exemplars)."""
        self._top_k = int(k)
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.top_k = self._top_k
        te = getattr(self, "_train_engine", None)
        if te is not None:
            te.infer.top_k = self._top_k

    # ---- nn.Module plumbing --------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._engine = None
        self._train_engine = None
        return out

    def _apply(self, fn, *a, **k):
        self._engine = None
        self._train_engine = None  # parameter storage may move: the flat master buffer is rebuilt on next use
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

    def postprocess(self, outputs: dict) -> dict:
        """BaseModel.postprocess (base_model.py:367-389): token ids (``seq``) -- or ``logits`` [B, S, V], arg-maxed under the
        tokenizer's per-position vocabulary mask -- to the layout dict {label, mask, center_x, center_y, width, height}."""
        if "seq" in outputs:
            seq = outputs["seq"]
        else:
            logits = outputs["logits"].detach().to(torch.float32).cpu().clone()
            if logits.size(-1) == self.tokenizer.max_token_length:  # (B, C, S) layout of the diffusion models
                logits = logits.permute(0, 2, 1)
            logits[:, ~self.tokenizer.token_mask.bool()] = -float("inf")
            seq = torch.argmax(logits, dim=-1)
        return self.tokenizer.decode(seq.cpu())

    def compute_stats(self) -> None:
        n = sum(p.numel() for p in self.parameters()) / 1e6
        print(f"number of parameters: {n:.2f}M")

    def update_per_epoch(self, epoch: int, warmup: int, epochs: int) -> None:
        pass

    def aggregate_sampling_config(self, sampling_cfg, test_cfg=None):
        """base_model.py:148-186 for an autoregressive class: for ``cond_type == "refinement"`` the test config's
        ``refine_mode / refine_offset_ratio / refine_lambda`` are copied into the sampling config (the keys the reference's
        logit-adjustment code reads; the autoregressive sampler itself ignores them)."""
        if test_cfg is not None and _cfg_get(test_cfg, "cond_type") == "refinement":
            assert _cfg_get(test_cfg, "refine_lambda") > 0.0
            try:
                from omegaconf import open_dict  # struct-mode DictConfig needs it; plain dicts do not
            except ImportError:
                import contextlib

                open_dict = contextlib.nullcontext
            for name in ("mode", "offset_ratio", "lambda"):
                key = f"refine_{name}"
                with open_dict(sampling_cfg):
                    sampling_cfg[key] = _cfg_get(test_cfg, key)
        return sampling_cfg

    # ---- multitask plumbing (retrieval_augmented_autoreg.py:707-750) --------------------------------
    def set_task_preprocessor(self, task: Optional[str]) -> None:
        assert task in T.COND_TYPES, f"{task=} must be one of {T.COND_TYPES}"
        if not self.use_multitask:
            return
        self.auxilary_task = task
        self.preprocessor = self._make_preprocessor(task)

    def get_random_task(self) -> str:
        """Task mixture of LayoutFormer++ (Tab. 3 of its supplement; retrieval_augmented_autoreg.py:721-735)."""
        import random

        tasks = ["uncond", "c", "cwh", "partial", "refinement", "relation"]
        return random.choices(tasks, weights=[1 / 12, 1 / 3, 1 / 3, 1 / 12, 1 / 3, 1 / 12])[0]

    # ---- train.py / inference.py surface --------------------------------------------------------
    def preprocess(self, inputs: dict) -> tuple[dict, dict]:
        """retrieval_augmented_autoreg.py:764-785 -> :509-523 (autoreg.py equivalent)."""
        if self.use_multitask:
            self.set_task_preprocessor(self.get_random_task())
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

    def optim_groups(self, base_lr: Optional[float] = None, weight_decay: float = 0.0,
                     forced_no_weight_decay: Optional[list] = None, custom_lr: Optional[dict] = None) -> list:
        """models/common/base_model.py:207-347: Linear / MultiheadAttention / Conv weights decay, biases and LayerNorm /
        BatchNorm / Embedding weights do not; ``custom_lr`` {prefix: lr} groups come first (train.py:217-223 passes
        {"encoder.extractor.body": lr * 0.1}); parameter names sorted inside every group, frozen parameters left out."""
        from .train import _is_decay

        named = {n: p for n, p in self.named_parameters() if p.requires_grad}
        forced = set(forced_no_weight_decay or [])
        decay = {n for n, p in named.items() if _is_decay(n, p) and n not in forced}
        no_decay = set(named) - decay
        groups, taken = [], set()
        for prefix, lr in (custom_lr or {}).items():
            for names, wd in ((decay, weight_decay), (no_decay, 0.0)):
                sel = sorted(n for n in names if n.startswith(prefix))
                taken.update(sel)
                if sel:
                    groups.append({"params": [named[n] for n in sel], "weight_decay": wd, "lr": lr})
        for names, wd in ((decay, weight_decay), (no_decay, 0.0)):
            sel = sorted(names - taken)
            if sel or not custom_lr:
                groups.append({"params": [named[n] for n in sel], "weight_decay": wd, "lr": base_lr})
        return groups

    def trainer(self, **kw):
        """The TrainEngine bound to this module (created on first use; its flat fp32 buffer becomes the storage of the
        module's parameters).  ``trainer().train_step`` is the fused path; ``train_loss`` below is the drop-in one."""
        if self._train_engine is None:
            import torch.distributed as dist

            from .train import TrainEngine

            if dist.is_available() and dist.is_initialized() and "world_size" not in kw:
                # launched through the reference's ddp_setup (helpers/distrubuted.py:10-20): do the gradient all-reduce
                # its DDPWrapper was meant to do (train_loss is reached through __getattr__, so DDP's reducer never arms)
                kw = {**kw, "world_size": dist.get_world_size(), "rank": dist.get_rank()}
            self._train_engine = TrainEngine(self, **kw)
        return self._train_engine

    def train_loss(self, inputs: dict, targets: dict, test: bool = False):
        """CrossEntropyLoss(label_smoothing=0.1, ignore_index=pad) over b s c -> b c s (:209-216).
        model.train() with grad enabled: the loss carries a backward that runs our kernels and fills ``p.grad``
        (TrainEngine.loss_with_grad), so train.py:440-454 works as is.  Otherwise (evaluate(): eval + no_grad): value only."""
        from . import ops

        if self.training and torch.is_grad_enabled() and not test:
            loss, logits = self.trainer().loss_with_grad(inputs, targets)
            self._engine = None  # inference operands are re-prepared from the updated parameters on demand
            return {"logits": logits}, {"nll_loss": loss}
        outputs = self(inputs)
        loss = ops.ce_label_smooth(outputs["logits"], targets["seq"].to(outputs["logits"].device), 0.1,
                                   self.tokenizer.name_to_id("pad"))
        return outputs, {"nll_loss": loss}

    @torch.no_grad()
    def sample(self, cond: Any, batch_size: Optional[int] = None, sampling_cfg: Any = None,
               cond_type: Optional[str] = "uncond", return_violation: bool = False, use_backtrack: bool = True,
               return_decoded_cond: bool = False, return_seq: bool = False, generator: Optional[torch.Generator] = None,
               **kwargs: Any):
        """Generation (retrieval_augmented_autoreg.py:218-325) with KV caches on the GPU: tasks uncond / c / cwh /
        partial / refinement / relation; sampling deterministic / random / top_k / top_p / gumbel
        (helpers/sampling.py:18-68).  ``generator``: optional torch CUDA generator for the uniforms of the stochastic
        samplers.  ``relation`` with ``use_backtrack`` goes through ``sample_relation`` (:335-507)."""
        assert cond_type in T.COND_TYPES, f"{cond_type=}"
        if cond_type == "relation" and use_backtrack:
            return self.sample_relation(cond, batch_size=batch_size, sampling_cfg=sampling_cfg,
                                        return_violation=return_violation, return_decoded_cond=return_decoded_cond,
                                        return_seq=return_seq, **kwargs)
        if self.use_multitask:
            self.set_task_preprocessor(getattr(cond, "task", cond_type))
        name = _cfg_get(sampling_cfg, "name", "deterministic") or "deterministic"
        sampling = {"name": name, "temperature": _cfg_get(sampling_cfg, "temperature", 1.0),
                    "top_k": _cfg_get(sampling_cfg, "top_k", 5), "top_p": _cfg_get(sampling_cfg, "top_p", 0.9)}
        image = cond.image
        B = image.size(0)
        if B == 1 and batch_size and batch_size > 1:
            assert cond_type in T.UNCOND, "batch_size expansion is only defined for unconstrained generation"
            B = batch_size
            image = image.expand(B, -1, -1, -1)
            cond = ConditionalInputs(image=image, task=cond_type, retrieved=getattr(cond, "retrieved", {}))
        const = self.preprocessor(cond)  # rewrites <eos> -> <pad> inside cond.seq like the reference (task.py)
        ids = self.special_token_ids
        steps = self.tokenizer.max_token_length
        forced = T.forced_token_table(cond_type, getattr(cond, "seq", None), ids["pad"], ids["eos"], steps,
                                      self.tokenizer.N_var_per_element)
        eng = self.engine()
        mem, mem_s = eng.encode(image, getattr(cond, "retrieved", None) if self.IS_RALF else None, const["seq"],
                                const["pad_mask"])
        if _SAMPLE_GRAPH and forced is None and name == "deterministic":
            seq = eng.generate_graphed(mem_s, B, mem.shape[1], self.tokenizer.token_mask, ids["bos"], ids["pad"], steps)
        else:
            seq = eng.generate(mem_s, B, mem.shape[1], self.tokenizer.token_mask, ids["bos"], ids["pad"], steps,
                               forced=forced, sampling=sampling, rng=generator)
        seq = seq.cpu()
        out = self.tokenizer.decode(seq)  # BaseModel.postprocess (base_model.py:367-389)
        if return_seq:
            out["seq"] = seq
        if return_decoded_cond:
            out["decoded_tokens"] = self.preprocessor.decode_tokens(const["seq"])
        if not return_violation:
            return out
        prepared = None
        if cond_type == "relation":  # only the label slots were restricted (:97-105); the relations are scored afterwards
            from .relation import RelationConstraint

            fn = RelationConstraint(self.preprocessor)
            prepared = [fn.prepare(const["seq"][b]) for b in range(B)]
        vio = T.calculate_violation(cond_type, cond, seq, self.tokenizer, output=out,
                                    prepared_rel_constraints=prepared)  # violate.py:24-139
        if cond_type in ("none", "uncond", "c", "cwh", "refinement"):
            assert vio["viorated"] == 0, f"{vio=}"
        return out, vio

    @torch.no_grad()
    def sample_relation(self, cond: Any, batch_size: Optional[int] = None, sampling_cfg: Any = None,
                        return_violation: bool = False, prob_gate: float = 0.3, RELATION_SIZE: int = 10,
                        return_decoded_cond: bool = False, return_seq: bool = False, **kwargs: Any):
        """Gen-R with backtracking (retrieval_augmented_autoreg.py:335-507).  The batch is encoded once on the GPU; every
        canvas is then decoded on its own through a KV-cached ``DecodeSession`` (a rewind keeps the cache rows of the
        surviving prefix) under ``relation.sample_with_backtracking``, which owns the host RNG draws of the reference
        (``random.randint`` for the rewind position, ``torch.multinomial`` on the host row for the token)."""
        from . import relation as R
        from .engine import KV24, DecodeSession

        if self.use_multitask:
            self.set_task_preprocessor("relation")
        pre = self.preprocessor
        assert isinstance(pre, R.RelationPreprocessor), "the model was not built / switched to the relation task"
        pre.set_relation_size(RELATION_SIZE)
        image = cond.image
        B = image.size(0)
        if B == 1 and batch_size and batch_size > 1:
            raise NotImplementedError("batch_size expansion is only defined for unconstrained generation")
        const = pre(cond)
        ids = self.special_token_ids
        steps = self.tokenizer.max_token_length
        forced = T.forced_token_table("relation", cond.seq, ids["pad"], ids["eos"], steps,
                                      self.tokenizer.N_var_per_element)
        eng = self.engine()
        mem, mem_s = eng.encode(image, getattr(cond, "retrieved", None) if self.IS_RALF else None, const["seq"],
                                const["pad_mask"])
        Mlen = mem.shape[1]
        kvm = eng.cross_kv(mem_s, kv24=KV24 and eng.npass == 3)
        fn = R.RelationConstraint(pre)
        rows, prepared = [], []
        for b in range(B):
            cons = fn.prepare(const["seq"][b])
            session = DecodeSession(eng, [k[b * Mlen:(b + 1) * Mlen] for k in kvm], Mlen, steps, ids["pad"])
            rows.append(R.sample_with_backtracking(session.logits_of, fn, cons, forced[b], bos_id=ids["bos"],
                                                   eos_id=ids["eos"], max_token_length=steps,
                                                   sampling_cfg=sampling_cfg, prob_gate=prob_gate))
            prepared.append(cons)
        seq = R.pad_like_reference(rows, steps)
        out = self.tokenizer.decode(seq)
        if return_seq:
            out["seq"] = seq
        if return_decoded_cond:
            out["decoded_tokens"] = pre.decode_tokens(const["seq"])
        if not return_violation:
            return out
        return out, R.violation_count(out, prepared)


class ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg(_B200LayoutModel):
    """The shipped RALF architecture (retrieval_augmented_autoreg.py:998-1033)."""

    IS_RALF = True


class ConcateAuxilaryTaskAutoreg(_B200LayoutModel):
    """The Autoreg baseline (models/autoreg.py:590-622); BASELINE.json config 1."""

    IS_RALF = False


RALF = ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg
