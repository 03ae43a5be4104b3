"""Checkpoint / resume and the evaluation loop around the training step.

* ``save_model`` / ``load_model``: the reference's files (helpers/io.py:42-74): ``<ckpt_dir>/<prefix>_<tag>_model.pt``
  holding the plain ``state_dict`` (DDP ``.module`` unwrapped) -- interchangeable with the reference in both directions.
* ``save_train_state`` / ``load_train_state``: what the reference does NOT save and a resumable run needs: AdamW moments
  (per parameter name), step count, dropout seed, epoch, best validation loss.
* ``multistep_lr``: ``MultiStepLRScheduler`` (schedulers/multi_step_lr.py:9-46; config/scheduler/multi_step_lr.yaml:
  milestones as fractions of the epoch count, gamma 0.1) as a pure function of the epoch.
* ``evaluate``: train.py:492-527, with the validation set SHARDED over ranks (the reference evaluates the full set on
  every rank) and the mean reduced over the process group.
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence

import torch


def _model_path(ckpt_dir: str, best_or_final: str, prefix: Optional[str]) -> str:
    name = f"{prefix}_{best_or_final}_model.pt" if prefix else f"{best_or_final}_model.pt"
    return os.path.join(ckpt_dir, name)


def save_model(model: torch.nn.Module, ckpt_dir: str, best_or_final: str = "best", prefix: Optional[str] = None) -> str:
    path = _model_path(ckpt_dir, best_or_final, prefix)
    os.makedirs(ckpt_dir, exist_ok=True)
    sd = model.module.state_dict() if hasattr(model, "module") else model.state_dict()
    torch.save({k: v.detach().cpu() for k, v in sd.items()}, path)
    return path


def load_model(model: torch.nn.Module, ckpt_dir: str, device, best_or_final: str = "best",
               prefix: Optional[str] = None) -> torch.nn.Module:
    with open(_model_path(ckpt_dir, best_or_final, prefix), "rb") as f:
        model.load_state_dict(torch.load(f, map_location=device))
    return model


def save_train_state(engine, path: str, *, epoch: int = 0, best_val_loss: float = float("inf")) -> None:
    """AdamW state of a :class:`ralf_b200.train.TrainEngine` keyed by parameter name (layout-independent)."""
    ps = engine.ps
    state = {"step_count": engine.step_count, "base_seed": engine.base_seed, "dropout": engine.dropout, "epoch": epoch,
             "best_val_loss": best_val_loss, "exp_avg": {}, "exp_avg_sq": {}}
    for name, (off, cnt, shape) in ps.offsets.items():
        state["exp_avg"][name] = ps.flat_m[off:off + cnt].view(shape).cpu().clone()
        state["exp_avg_sq"][name] = ps.flat_v[off:off + cnt].view(shape).cpu().clone()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(state, path)


def load_train_state(engine, path: str) -> dict:
    """Restore moments / counters saved by :func:`save_train_state` into ``engine`` (the model weights travel separately
    through ``load_model``; call this after the engine has been built on the loaded model)."""
    state = torch.load(path, map_location="cpu", weights_only=False)
    ps = engine.ps
    missing = set(ps.offsets) - set(state["exp_avg"])
    if missing:
        raise KeyError(f"optimizer state lacks {sorted(missing)[:3]}... ({len(missing)} parameters)")
    for name, (off, cnt, shape) in ps.offsets.items():
        ps.flat_m[off:off + cnt].copy_(state["exp_avg"][name].reshape(-1))
        ps.flat_v[off:off + cnt].copy_(state["exp_avg_sq"][name].reshape(-1))
    engine.step_count = int(state["step_count"])
    if "base_seed" in state:  # the per-rank dropout stream is re-derived from the base seed and THIS engine's rank
        engine.base_seed = int(state["base_seed"])
    engine._weights_changed()
    return {k: state[k] for k in ("epoch", "best_val_loss", "step_count")}


def multistep_lr(base_lr: float, epoch: int, epochs: int, milestones: Sequence = (0.7,), gamma: float = 0.1) -> float:
    """Learning rate DURING 1-based ``epoch`` under MultiStepLRScheduler stepped once per epoch (train.py:283-287):
    float milestones are fractions of ``epochs`` (int(m * epochs)), ints are absolute epochs."""
    ms = [int(m * epochs) if isinstance(m, float) else int(m) for m in milestones]
    done = epoch - 1  # scheduler.step() calls so far
    return base_lr * gamma ** sum(1 for m in ms if done >= m)


@torch.no_grad()
def evaluate(model, batches: Iterable[dict], *, rank: int = 0, world_size: int = 1, process_group=None) -> dict:
    """Mean validation losses (train.py:492-527).  ``batches`` are collated batches; rank r takes batches r, r+W, ...;
    sums and counts are all-reduced, so every rank returns the global means."""
    was_training = model.training
    model.eval()
    dev = model.device
    total, count = None, 0
    for i, batch in enumerate(batches):
        if i % world_size != rank:
            continue
        inputs, targets = model.preprocess(batch)
        inputs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inputs.items()}
        targets = {k: v.to(dev) for k, v in targets.items()}
        _, losses = model.train_loss(inputs, targets, test=True)
        vals = torch.stack([losses[k].detach().float().reshape(()) for k in sorted(losses)])
        total = vals if total is None else total + vals
        names = sorted(losses)
        count += 1
    if total is None:
        total, names = torch.zeros(1, device=dev), ["nll_loss"]
    packed = torch.cat([total, torch.tensor([float(count)], device=total.device)])
    if world_size > 1:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():  # otherwise: this shard's local mean
            dist.all_reduce(packed, group=process_group)
    out = {k: float(packed[i] / packed[-1].clamp(min=1)) for i, k in enumerate(names)}
    out["total"] = sum(v for k, v in out.items() if "loss" in k)
    if was_training:
        model.train()
    return out
