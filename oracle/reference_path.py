"""CPU port of the reference hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package must never do so.

What is restated (reference = ``/root/reference``; it does not exist on the GPU box, so the
reference module itself cannot travel -- this port, which calls the very same third-party
functions the reference calls, does):

* ``src/feature_extraction.py:184-207``  ``build_transform`` -> ``port_transform``
* ``src/feature_extraction.py:210-227``  ``load_model``      -> ``port_model`` (weight download
  neutralised: ``weights=None`` under a fixed seed)
* ``src/feature_extraction.py:233-240``  ``preprocess_image``
* ``src/feature_extraction.py:243-313``  ``batched`` / ``extract_embeddings`` loop, incl. the
  decode-failure bookkeeping at ``:281-284`` and the ``per_file_times`` rule at ``:297-300``

The arithmetic lives in torchvision 0.26.0 / Pillow 12.2.0 / torch 2.11 (installed in the image on
both boxes; ``pyproject.toml:11-15`` pins lower bounds only).

Pinned by ``tests/test_oracle_cpu.py``: same outputs as the real ``src.feature_extraction`` module
(when ``/root/reference`` is mounted) and as the committed golden vectors in ``tests/golden/``.
"""
from __future__ import annotations

import ctypes
import logging
import os
import subprocess
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
from PIL import Image, UnidentifiedImageError
from torch import nn
from torchvision import models, transforms

MEAN = [0.485, 0.456, 0.406]  # src/feature_extraction.py:64
STD = [0.229, 0.224, 0.225]  # src/feature_extraction.py:65
RESIZE, CROP = 256, 224  # src/feature_extraction.py:66-67
WEIGHT_SEED = 1234
BN_SEED = 1235


@dataclass(frozen=True)
class PortRecord:
    """Same four fields as the reference's ImageRecord (src/feature_extraction.py:85-92)."""

    absolute_path: Path
    relative_path: Path
    bucket: str
    label: Optional[str]


@dataclass
class PortResults:
    """Same four fields as ExtractionResults (src/feature_extraction.py:95-102)."""

    embeddings: np.ndarray
    records: list
    failures: list
    per_file_times: list


def make_backbone(seed: int = WEIGHT_SEED, randomize_bn: bool = False) -> nn.Module:
    """torchvision resnet18 with seeded random init instead of the IMAGENET1K_V1 download.

    ``randomize_bn`` replaces the identity BatchNorm statistics torchvision initialises
    (tv:models/resnet.py:208-213) with seeded non-trivial ones so BN folding is exercised.
    """
    torch.manual_seed(seed)
    net = models.resnet18(weights=None)
    if randomize_bn:
        gen = torch.Generator().manual_seed(BN_SEED)
        for mod in net.modules():
            if isinstance(mod, nn.BatchNorm2d):
                c = mod.num_features
                mod.weight.data = 0.8 + 0.4 * torch.rand(c, generator=gen)
                mod.bias.data = 0.1 * torch.randn(c, generator=gen)
                mod.running_mean.data = 0.1 * torch.randn(c, generator=gen)
                mod.running_var.data = 0.6 + 0.8 * torch.rand(c, generator=gen)
    return net


def port_model(device: torch.device, seed: int = WEIGHT_SEED, randomize_bn: bool = False) -> nn.Module:
    """load_model (src/feature_extraction.py:210-227) minus the download."""
    net = make_backbone(seed, randomize_bn)
    net.eval()
    for p in net.parameters():
        p.requires_grad_(False)
    trunk = nn.Sequential(*list(net.children())[:-1])
    trunk.eval()
    trunk.to(device)
    return trunk


def port_transform() -> Callable[[Image.Image], torch.Tensor]:
    """build_transform (src/feature_extraction.py:200-207)."""
    return transforms.Compose(
        [
            transforms.Resize(RESIZE),
            transforms.CenterCrop(CROP),
            transforms.ToTensor(),
            transforms.Normalize(mean=MEAN, std=STD),
        ]
    )


def port_preprocess_image(path: Path, transform) -> torch.Tensor:
    """preprocess_image (src/feature_extraction.py:233-240): no mode conversion."""
    with Image.open(path) as img:
        return transform(img)


def port_preprocess_array(arr: np.ndarray, transform=None) -> torch.Tensor:
    """Same transform on an already-decoded HWC uint8 array (what Image.open would hand over)."""
    transform = transform or port_transform()
    return transform(Image.fromarray(arr))


def port_extract_embeddings(
    records: Sequence, device: torch.device, batch_size: int = 32, seed: int = WEIGHT_SEED, randomize_bn: bool = False
) -> PortResults:
    """extract_embeddings (src/feature_extraction.py:251-313)."""
    transform = port_transform()
    model = port_model(device, seed, randomize_bn)
    chunks: List[np.ndarray] = []
    kept, failures, times = [], [], []
    for start in range(0, len(records), batch_size):
        batch = records[start : start + batch_size]
        tensors, ok = [], []
        t0 = time.perf_counter()
        for rec in batch:
            try:
                tensors.append(port_preprocess_image(rec.absolute_path, transform))
                ok.append(rec)
            except (UnidentifiedImageError, OSError) as exc:
                logging.error("Failed to decode %s: %s", rec.absolute_path, exc)
                failures.append(rec.absolute_path)
        if not tensors:
            continue
        x = torch.stack(tensors).to(device)
        with torch.no_grad():
            feats = torch.flatten(model(x), 1)
        chunks.append(feats.cpu().numpy())
        kept.extend(ok)
        dt = time.perf_counter() - t0
        times.extend([dt / len(ok)] * len(ok))
    if not chunks:
        raise RuntimeError("No embeddings were generated; all images failed to decode?")
    return PortResults(np.concatenate(chunks, axis=0), kept, failures, times)


def port_embed_arrays(
    images: Sequence[np.ndarray], batch_size: int = 32, seed: int = WEIGHT_SEED, randomize_bn: bool = False
) -> np.ndarray:
    """Embeddings of decoded HWC uint8 arrays through the same transform + trunk, CPU fp32."""
    transform = port_transform()
    model = port_model(torch.device("cpu"), seed, randomize_bn)
    out = []
    for start in range(0, len(images), batch_size):
        x = torch.stack([port_preprocess_array(a, transform) for a in images[start : start + batch_size]])
        with torch.no_grad():
            out.append(torch.flatten(model(x), 1).numpy())
    return np.concatenate(out, axis=0)


# --------------------------------------------------------------------------------------------
# Classifier-head inference: restatement of the three reference loops that run the
# fine-tuned resnet18 in eval mode.  Same third-party calls; no training, no plots.
# --------------------------------------------------------------------------------------------

HEAD_SEED = 4321


def port_eval_transform() -> Callable[[Image.Image], torch.Tensor]:
    """build_transforms()["eval"] (src/training/common.py:111-117)."""
    return transforms.Compose(
        [transforms.Resize((CROP, CROP)), transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)]
    )


_HEAD_CACHE: dict = {}


def make_classifier(num_classes: int = 2, seed: int = WEIGHT_SEED, randomize_bn: bool = True) -> nn.Module:
    """create_model(num_classes, pretrained=False) (src/training/common.py:299-304) with the seeded backbone and a
    seeded fc that is CALIBRATED on 16 seeded images so that each logit has mean 0 / std 2 over them: the post-ReLU
    embeddings are all positive with large per-dimension means, so an uncalibrated random head
    predicts one class with probability 1.0 for every input and would test nothing."""
    net = make_backbone(seed, randomize_bn)
    net.fc = nn.Linear(net.fc.in_features, num_classes)
    net.eval()
    key = (num_classes, seed, randomize_bn)
    if key not in _HEAD_CACHE:
        from ssip_b200 import synthetic  # seeded generators only

        gen = torch.Generator().manual_seed(HEAD_SEED)
        w = torch.randn(net.fc.weight.shape, generator=gen, dtype=torch.float64)
        calib = list(synthetic.mri_like_images(8, 512, seed=HEAD_SEED)) + list(synthetic.noise_images(8, 300, 260, seed=HEAD_SEED + 1))
        t = port_eval_transform()
        x = torch.stack([t(Image.fromarray(a)) for a in calib])
        trunk = nn.Sequential(*list(net.children())[:-1])
        with torch.no_grad():
            emb = torch.flatten(trunk(x), 1).double()
        z = emb @ w.T
        w = w * (2.0 / z.std(dim=0))[:, None]
        bias = -(emb @ w.T).mean(dim=0)
        _HEAD_CACHE[key] = (w.float(), bias.float())
    w, bias = _HEAD_CACHE[key]
    with torch.no_grad():
        net.fc.weight.copy_(w)
        net.fc.bias.copy_(bias)
    return net


def port_generate_pseudo_labels(model, data_loader, device, threshold: float = 0.7):
    """generate_pseudo_labels (src/training/semi_supervised.py:44-72): (path, label, confidence) above threshold."""
    model.eval()
    out = []
    with torch.no_grad():
        for images, paths in data_loader:
            probs = torch.softmax(model(images.to(device)), dim=1)
            conf, pred = torch.max(probs, dim=1)
            for path, p, c in zip(paths, pred.cpu().numpy(), conf.cpu().numpy()):
                if c >= threshold:
                    out.append((path, int(p), float(c)))
    return out


def port_compute_probs(model, loader, device, pos_index: int):
    """compute_probs (src/threshold_sweep.py:21-38)."""
    model.eval()
    y_true, y_prob = [], []
    with torch.no_grad():
        for batch in loader:
            inputs, labels = batch[:2]
            probs = torch.softmax(model(inputs.to(device)), dim=1)[:, pos_index]
            y_true.extend(labels.cpu().numpy().tolist())
            y_prob.extend(probs.cpu().numpy().tolist())
    return np.array(y_true), np.array(y_prob)


# --------------------------------------------------------------------------------------------
# ctypes view of the plain-C restatement (oracle/preprocess_oracle.c)
# --------------------------------------------------------------------------------------------

_HERE = Path(__file__).resolve().parent
_LIB: Optional[ctypes.CDLL] = None


def build_c_oracle() -> Path:
    """Compile oracle/preprocess_oracle.c with gcc (idempotent)."""
    so = _HERE / "libfx_oracle.so"
    src = _HERE / "preprocess_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return so


def c_oracle() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(str(build_c_oracle()))
        lib.fxo_ksize.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.fxo_coeffs.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.fxo_resized_size.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)] * 2
        lib.fxo_crop_offset.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.fxo_resize_bilinear_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] + [ctypes.c_int] * 2
        lib.fxo_build_lut.argtypes = [ctypes.c_void_p]
        lib.fxo_resize_crop_u8.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p]
        lib.fxo_preprocess_rgb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.fxo_preprocess_square224_rgb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def c_preprocess_rgb(arr: np.ndarray) -> np.ndarray:
    """C restatement of the whole transform: HWC uint8 (3 ch) -> fp32 [3,224,224]."""
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    assert arr.ndim == 3 and arr.shape[2] == 3
    out = np.empty((3, CROP, CROP), np.float32)
    rc = c_oracle().fxo_preprocess_rgb(arr.ctypes.data, arr.shape[0], arr.shape[1], out.ctypes.data)
    if rc != 0:
        raise ValueError("resized image smaller than the crop")
    return out


def c_preprocess_square224(arr: np.ndarray) -> np.ndarray:
    """C restatement of the evaluation transform (Resize((224,224)) + ToTensor + Normalize): fp32 [3,224,224]."""
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    assert arr.ndim == 3 and arr.shape[2] == 3
    out = np.empty((3, CROP, CROP), np.float32)
    c_oracle().fxo_preprocess_square224_rgb(arr.ctypes.data, arr.shape[0], arr.shape[1], out.ctypes.data)
    return out


def c_resize_crop(arr: np.ndarray) -> np.ndarray:
    """C restatement of Resize(256)+CenterCrop(224): HWC uint8 -> uint8 [224,224,C]."""
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    out = np.empty((CROP, CROP, arr.shape[2]), np.uint8)
    rc = c_oracle().fxo_resize_crop_u8(arr.ctypes.data, arr.shape[0], arr.shape[1], arr.shape[2], out.ctypes.data)
    if rc != 0:
        raise ValueError("resized image smaller than the crop")
    return out


def c_coeffs(in_size: int, out_size: int):
    lib = c_oracle()
    ks = lib.fxo_ksize(in_size, out_size)
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ks), np.int32)
    lib.fxo_coeffs(in_size, out_size, bounds.ctypes.data, kk.ctypes.data)
    return bounds, kk


def c_lut() -> np.ndarray:
    lut = np.empty((3, 256), np.float32)
    c_oracle().fxo_build_lut(lut.ctypes.data)
    return lut


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
