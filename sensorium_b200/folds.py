"""Fold-parallel training driver.

The reference trains the 7 cross-validation folds one after the other on one GPU (scripts/train.py:173-189, 12 h per
fold on an A6000); every fold runs the same two stages — a linear warm-up (``LambdaLR(lambda x: x / num_iterations)``,
scripts/train.py:122-126) and a cosine-annealed training stage with an ``EmaCheckpoint`` that keeps the best
``model-{epoch:03d}-{val_corr:.6f}.pth`` (scripts/train.py:127-136).  The folds are independent, so on an 8-GPU box
they run side by side: one process per GPU (torchrun), fold ``k`` on rank ``k % world``, no collective on the training
path (a barrier at the end only).  The on-disk result is what the reference produces — ``experiment_dir/fold_k/
model-*.pth`` in the argus checkpoint layout — so ``get_best_model_path`` / ``Predictor`` / ``EnsemblePredictor`` pick the
models up unchanged.

Only the loop around ``MouseModel.train_step`` / ``val_step`` is here; datasets and loaders are the caller's (the
reference's CPU data pipeline is out of scope): ``make_loaders(train_splits, val_splits) -> (train_loader, val_loader)``.
"""
from __future__ import annotations

import copy
import math
import re
from pathlib import Path
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import constants
from .argus_models import MouseModel
from .ema import ModelEma, save_ema_model
from .metrics import CorrelationMetric
from .utils import get_lr, init_weights

folds_splits = [f"fold_{fold}" for fold in constants.folds]


def get_best_model_path(dir_path, return_score: bool = False, more_better: bool = True):
    """utils.py:22-43: the checkpoint whose file name carries the best score (``...-<score>.pth``)."""
    dir_path = Path(dir_path)
    model_scores = []
    for model_path in dir_path.glob("*.pth"):
        score = re.search(r"-(\d+(?:\.\d+)?).pth", str(model_path))
        if score is not None:
            model_scores.append((model_path, float(score.group(0)[1:-4])))
    if not model_scores:
        if return_score:
            return None, -np.inf if more_better else np.inf
        return None
    best = sorted(model_scores, key=lambda x: x[1], reverse=more_better)[0]
    return (best[0], best[1]) if return_score else best[0]


def shard_folds(folds: Sequence[str], rank: int, world: int) -> List[str]:
    """Folds trained by ``rank``: round-robin (7 folds on 8 GPUs: one fold per GPU, one GPU idle)."""
    return [f for i, f in enumerate(folds) if i % world == rank]


def stage_lr(stage: str, base_lr: float, min_lr: float, iteration: int, num_iterations: int) -> float:
    """Learning rate of iteration ``iteration`` (0-based, number of scheduler steps taken so far) of a stage:
    warm-up = torch LambdaLR(lambda x: x / num_iterations); train = torch CosineAnnealingLR(T_max, eta_min)
    stepped every iteration (scripts/train.py:122-136)."""
    if stage == "warmup":
        return base_lr * iteration / num_iterations
    if stage == "train":
        return min_lr + (base_lr - min_lr) * (1.0 + math.cos(math.pi * iteration / num_iterations)) / 2.0
    raise ValueError(f"unknown stage '{stage}'")


def _validate(model: MouseModel, val_loader) -> Dict[str, float]:
    metric = CorrelationMetric()
    metric.reset()
    losses = []
    for batch in val_loader:
        out = model.val_step(batch, None)
        metric.update(out)
        losses.append(out["loss"])
    state = SimpleNamespace(phase="val", metrics={})
    metric.epoch_complete(state)
    state.metrics["val_loss"] = float(np.mean(losses)) if losses else float("nan")
    return state.metrics


def train_fold(config: dict, save_dir, train_loader, val_loader, log: Optional[Callable[[str], None]] = None,
               distill_model_path=None) -> Optional[Path]:
    """One fold: scripts/train.py:43-170 without the data pipeline.  Returns the path of the best checkpoint."""
    from .predictors import _load_model
    config = copy.deepcopy(config)
    save_dir = Path(save_dir)
    save_dir.mkdir(parents=True, exist_ok=True)
    log = log or (lambda s: None)
    model = MouseModel(config["argus_params"])
    if config.get("init_weights"):
        init_weights(model.nn_module)
    if config.get("ema_decay"):
        model.model_ema = ModelEma(model.nn_module, decay=config["ema_decay"])
    if distill_model_path is not None:
        teacher = _load_model(distill_model_path, device=config["argus_params"]["device"], optimizer=None, loss=None)
        teacher.eval()
        model.distill_model = teacher.nn_module
        model.distill_ratio = config["distill"]["ratio"]
    base_lr = model.get_lr()[0]
    min_lr = get_lr(config.get("min_base_lr", 0.0), config["batch_size"])
    best_path, best_score, epoch = None, -np.inf, 0
    for num_epochs, stage in zip(config["num_epochs"], config["stages"]):
        num_iterations = len(train_loader) * num_epochs
        it = 0
        for _ in range(num_epochs):
            epoch += 1
            losses = []
            for batch in train_loader:
                model.set_lr(stage_lr(stage, base_lr, min_lr, it, num_iterations))
                losses.append(model.train_step(batch, None)["loss"])
                it += 1
            metrics = _validate(model, val_loader) if val_loader is not None else {}
            log(f"{save_dir.name} {stage} epoch {epoch}: train_loss {np.mean(losses):.5f} "
                + " ".join(f"{k} {float(v):.5f}" for k, v in metrics.items() if k in ("val_loss", "val_corr")))
            score = float(metrics.get("val_corr", -np.inf))
            if stage == "train" and score > best_score:
                new_path = save_dir / f"model-{epoch:03d}-{score:.6f}.pth"
                if model.model_ema is not None:
                    save_ema_model(model, new_path)            # EmaCheckpoint.save_model, ema.py:61-73
                else:
                    model.save(new_path)
                if best_path is not None and best_path != new_path and best_path.exists():
                    best_path.unlink()                         # max_saves=1
                best_path, best_score = new_path, score
    return best_path


def run_folds(config: dict, experiment_dir, make_loaders: Callable, folds: str = "all", group=None,
              log: Optional[Callable[[str], None]] = print) -> Dict[str, Optional[Path]]:
    """The fold loop of scripts/train.py:173-189 spread over the ranks of ``group`` (or run serially in one process).
    Every rank trains ``shard_folds(...)``; the result — {fold: best checkpoint path} of ALL folds — is gathered on every
    rank at the end."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    experiment_dir = Path(experiment_dir)
    all_folds = folds_splits if folds == "all" else [f"fold_{f}" for f in str(folds).split(",")]
    mine: Dict[str, Optional[str]] = {}
    for fold_split in shard_folds(all_folds, rank, world):
        val_splits = [fold_split]
        train_splits = sorted(set(folds_splits) - set(val_splits))
        cfg = copy.deepcopy(config)
        if torch.cuda.is_available() and str(cfg["argus_params"].get("device", "cuda:0")).startswith("cuda"):
            cfg["argus_params"]["device"] = f"cuda:{torch.cuda.current_device()}"
        train_loader, val_loader = make_loaders(train_splits, val_splits)
        distill = None
        if "distill" in cfg:  # scripts/train.py:57-65: the teacher is the best model of the same fold
            distill = get_best_model_path(experiment_dir.parent / cfg["distill"]["experiment"] / fold_split)
        path = train_fold(cfg, experiment_dir / fold_split, train_loader, val_loader, log=log, distill_model_path=distill)
        mine[fold_split] = str(path) if path is not None else None
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    result = dict(mine)
    if distributed:
        parts: List = [None] * world
        dist.all_gather_object(parts, mine, group=group)
        result = {}
        for p in parts:
            result.update(p)
    return {k: (Path(v) if v is not None else None) for k, v in sorted(result.items())}
