"""B200 drop-ins for the reference's classifier-head INFERENCE loops.

The reference runs its fine-tuned ``resnet18 + nn.Linear(512, C)`` (create_model, src/training/common.py:299-304)
in eval mode in three places, each a ``for batch in loader: softmax(model(inputs))`` loop:

* ``generate_pseudo_labels``  src/training/semi_supervised.py:44-72
* ``evaluate_model``          src/training/common.py:439-506
* ``compute_probs``           src/threshold_sweep.py:21-38

The functions here keep those names, arguments, return values and bookkeeping; ``outputs = model(inputs)`` and the
softmax run on the CUDA engine (``fx_stage_nchw_f32`` + ``fx_classify``: the same tcgen05 trunk as the embedding
path, then a fp32 head + softmax kernel).  ``model`` is the reference's own ``nn.Module``: its ``state_dict`` is
folded and uploaded once per (module, parameter version).  Training (forward+backward, optimiser) is out of scope.

``classify_arrays`` is the fused form for raw images: uint8 HWC arrays -> the evaluation transform
``Resize((224, 224))`` (src/training/common.py:111-117; FX_TRANSFORM_SQUARE224, bit-exact) -> trunk -> head, with
``convert("RGB")`` semantics for gray inputs (src/training/common.py:171,191).  No CPU fallback.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from PIL import Image

from . import _native as N
from .engine import Engine, pack_images
from .feature_extraction import IMAGENET_MEAN, IMAGENET_STD, PRECISION_ENV, _cuda_index

LOGGER = logging.getLogger(__name__)

_CLASSIFIERS: Dict[Tuple[int, str], "Classifier"] = {}


def _model_version(model: torch.nn.Module) -> Tuple[int, Tuple[int, ...]]:
    """Identity of the module + the in-place version counters of its tensors (bumped by optimiser steps / loads)."""
    return id(model), tuple(int(t._version) for t in model.state_dict().values())


class Classifier:
    """resnet18 trunk + fc head of one device's engine, loaded from the reference's nn.Module."""

    def __init__(self, device: torch.device, max_batch: int = 256, precision: str = "bf16"):
        self.engine = Engine(_cuda_index(device), max_batch, precision)
        self.engine.set_transform(N.TRANSFORM_SQUARE224)
        self._loaded = None

    def load(self, model: torch.nn.Module) -> None:
        version = _model_version(model)
        if version == self._loaded:
            return
        state = model.state_dict()
        if "fc.weight" not in state or "fc.bias" not in state:
            raise ValueError("expected the reference's create_model() network: a resnet18 whose fc is nn.Linear(512, C)")
        self.engine.load_state_dict(state)
        self.engine.load_head(state["fc.weight"], state["fc.bias"])
        self._loaded = version

    def logits_probs(self, inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """`outputs = model(inputs)`; `torch.softmax(outputs, dim=1)` for a transformed fp32 [n,3,224,224] batch."""
        eng = self.engine
        x = inputs.to(eng.device, torch.float32).contiguous()
        logits, probs = [], []
        with torch.cuda.device(eng.device):
            for s in range(0, x.shape[0], eng.max_batch):
                _, lg, pr = eng.classify_nchw(x[s : s + eng.max_batch])
                logits.append(lg)
                probs.append(pr)
        return torch.cat(logits), torch.cat(probs)

    def classify_arrays(self, arrays: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Decoded HWC uint8 arrays (3 channels, or 2-D gray == convert("RGB")) -> (embeddings, logits, probs)."""
        eng = self.engine
        outs = ([], [], [])
        with torch.cuda.device(eng.device):
            for s in range(0, len(arrays), eng.max_batch):
                part = arrays[s : s + eng.max_batch]
                buf, descs, total = pack_images(part)
                dev = torch.from_numpy(buf[: max(total, 1)]).to(eng.device)
                for dst, t in zip(outs, eng.classify_device(dev, descs, len(part))):
                    dst.append(t.cpu().numpy())
        return tuple(np.concatenate(o) if o else np.empty((0, 0), np.float32) for o in outs)


def get_classifier(model: torch.nn.Module, device: torch.device, min_batch: int = 64, precision: Optional[str] = None) -> Classifier:
    """Engine + head for `device`, (re)loaded from `model` when its parameters changed.  Precision: bf16 unless the
    SSIP_B200_PRECISION environment variable (or the argument) says fp32, as for the embedding path."""
    precision = precision or os.environ.get(PRECISION_ENV, "bf16")
    key = (_cuda_index(device), precision)
    clf = _CLASSIFIERS.get(key)
    if clf is not None and clf.engine.max_batch < min_batch:
        clf.engine.close()
        clf = None
    if clf is None:
        clf = _CLASSIFIERS[key] = Classifier(device, max(min_batch, 1), precision)
    clf.load(model)
    return clf


class _EvalTransform:
    """build_transforms()["eval"] (src/training/common.py:111-117) as the fused CUDA kernel: PIL image -> fp32
    [3,224,224] CPU tensor, bit-identical to the torchvision Compose.  Expects what the reference's datasets hand it,
    i.e. the image after ``.convert("RGB")``."""

    def __init__(self, device: Optional[torch.device] = None):
        self._device = torch.device(device) if device is not None else torch.device("cuda")
        self._engine: Optional[Engine] = None

    def __call__(self, img: Image.Image) -> torch.Tensor:
        if self._engine is None:
            self._engine = Engine(_cuda_index(self._device), 16, "fp32")
            self._engine.set_transform(N.TRANSFORM_SQUARE224)
        arr = np.asarray(img.convert("RGB"))
        buf, descs, total = pack_images([arr])
        with torch.cuda.device(self._engine.device):
            dev = torch.from_numpy(buf[:total]).to(self._engine.device)
            return self._engine.preprocess_nchw(dev, descs, 1)[0].cpu()

    def __repr__(self) -> str:
        return f"FusedCudaTransform(Resize(({N.CROP}, {N.CROP})), ToTensor(), Normalize(mean={IMAGENET_MEAN}, std={IMAGENET_STD}))"


def build_transforms(image_size: int = 224) -> Dict[str, Callable[[Image.Image], torch.Tensor]]:
    """Only the deterministic ``eval`` transform is an inference path; ``train`` (random flip / rotation,
    src/training/common.py:101-109) belongs to training and is not provided."""
    if image_size != N.CROP:
        raise ValueError("the trunk kernels are built for 224x224 inputs")
    return {"eval": _EvalTransform()}


def generate_pseudo_labels(model: torch.nn.Module, data_loader, device: torch.device, threshold: float = 0.7) -> List[Tuple[str, int, float]]:
    """src/training/semi_supervised.py:44-72: (path, predicted_label, confidence) for confidence >= threshold."""
    clf = get_classifier(model, device)
    pseudo_samples: List[Tuple[str, int, float]] = []
    for images, paths in data_loader:
        _, probabilities = clf.logits_probs(images)
        confidences, predictions = torch.max(probabilities, dim=1)
        for path, prediction, confidence in zip(paths, predictions.cpu().numpy(), confidences.cpu().numpy()):
            if confidence >= threshold:
                pseudo_samples.append((path, int(prediction), float(confidence)))
    LOGGER.info("Generated %d pseudo-labelled samples with threshold %.2f", len(pseudo_samples), threshold)
    return pseudo_samples


def compute_probs(model: torch.nn.Module, loader, device: torch.device, pos_index: int) -> Tuple[np.ndarray, np.ndarray]:
    """src/threshold_sweep.py:21-38: (y_true, P(class pos_index)) over the loader."""
    clf = get_classifier(model, device)
    y_true: List[int] = []
    y_prob: List[float] = []
    for batch in loader:
        inputs, labels = batch[:2]
        _, probs = clf.logits_probs(inputs)
        y_true.extend(torch.as_tensor(labels).cpu().numpy().tolist())
        y_prob.extend(probs[:, pos_index].cpu().numpy().tolist())
    return np.array(y_true), np.array(y_prob)


def evaluate_model(model: torch.nn.Module, data_loader, device: torch.device, pos_index: Optional[int] = None,
                   threshold: Optional[float] = None) -> Tuple[Dict[str, Any], np.ndarray, np.ndarray, np.ndarray, List[str]]:
    """src/training/common.py:439-506: metrics + y_true, y_pred, y_prob, sample paths."""
    from sklearn.metrics import accuracy_score, precision_recall_fscore_support

    clf = get_classifier(model, device)
    y_true: List[int] = []
    y_pred: List[int] = []
    y_prob: List[float] = []
    sample_paths: List[str] = []
    for batch in data_loader:
        inputs, labels = batch[:2]
        extras = batch[2:] if len(batch) > 2 else []
        paths = extras[0] if extras else ["" for _ in range(len(labels))]
        outputs, probs_full = clf.logits_probs(inputs)
        pos_col = (1 if probs_full.shape[1] > 1 else 0) if pos_index is None else pos_index
        probabilities = probs_full[:, pos_col]
        if threshold is None or probs_full.shape[1] != 2:
            predictions = outputs.argmax(dim=1)
        else:
            neg_col = 1 - pos_col
            predictions = torch.where(probabilities >= threshold, torch.tensor(pos_col, device=outputs.device),
                                      torch.tensor(neg_col, device=outputs.device))
        y_true.extend(torch.as_tensor(labels).cpu().numpy().tolist())
        y_pred.extend(predictions.cpu().numpy().tolist())
        y_prob.extend(probabilities.cpu().numpy().tolist())
        sample_paths.extend([str(p) for p in list(paths)])
    if pos_index is not None:
        y_true_bin = (np.array(y_true) == pos_index).astype(int)
        y_pred_bin = (np.array(y_pred) == pos_index).astype(int)
        accuracy = accuracy_score(y_true_bin, y_pred_bin)
        precision, recall, f1, _ = precision_recall_fscore_support(y_true_bin, y_pred_bin, average="binary", zero_division=0)
    else:
        accuracy = accuracy_score(y_true, y_pred)
        precision, recall, f1, _ = precision_recall_fscore_support(y_true, y_pred, average="binary", zero_division=0)
    metrics = {"accuracy": float(accuracy), "precision": float(precision), "recall": float(recall), "f1": float(f1)}
    return metrics, np.array(y_true), np.array(y_pred), np.array(y_prob), sample_paths
