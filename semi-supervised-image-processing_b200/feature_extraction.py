"""Drop-in for ``src.feature_extraction`` of Septimus4/semi-supervised-image-processing, running
the hot path on hand-written sm_100a CUDA (libfx_b200.so) instead of stock PyTorch.

Same CLI (``--data-dir --device --batch-size --verbose``, src/feature_extraction.py:510-535), same
importable names and the same five artifacts under ``outputs/``:

    python -m ssip_b200.feature_extraction --data-dir mri_dataset_brain_cancer_oc --device cuda    # every visible GPU
    python -m ssip_b200.feature_extraction --device cuda:3                                         # that GPU only
    torchrun --nproc-per-node 8 -m ssip_b200.feature_extraction --device cuda      # the same sharding under torchrun

What changes underneath: files are still decoded by Pillow on the host (a thread pool instead of
a serial loop), but Resize(256)/CenterCrop(224)/ToTensor/Normalize is one fused CUDA kernel that
reproduces the reference transform bit for bit, and the frozen ResNet-18 trunk runs as BN-folded
implicit-GEMM convolutions on the tcgen05 tensor cores.  ``--device cpu`` is refused: this
package has no CPU path (run the reference for that).

Weights.  The reference hard-codes the IMAGENET1K_V1 download (src/feature_extraction.py:217-218)
and so does this module by default.  Offline, point ``SSIP_B200_WEIGHTS`` at a torchvision
resnet18 ``state_dict`` file, or set it to ``random:<seed>`` (``random-bn:<seed>`` additionally
randomises the BatchNorm statistics) for a seeded random initialisation -- the parity tests and
bench.py use that.  ``SSIP_B200_PRECISION`` = ``bf16`` (default) or ``fp32`` (tight tolerance).
"""
from __future__ import annotations

import argparse
import atexit
import json
import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from datetime import datetime, timezone
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from PIL import Image, UnidentifiedImageError

from . import _artifacts
from . import _native as N
from ._decode_pool import HOST_RESIZE_SHORT_SIDE, DecodePool, host_resize_if_oversized, rebuild_exception
from . import dist as fxdist
from .engine import Engine, pack_images

# --- constants: same names and values as src/feature_extraction.py:53-77 -------------------------
DEFAULT_DATA_DIR = Path("mri_dataset_brain_cancer_oc")
DEFAULT_OUTPUT_ROOT = Path("outputs")
FEATURE_OUTPUT_DIR = DEFAULT_OUTPUT_ROOT / "features"
LOG_OUTPUT_DIR = DEFAULT_OUTPUT_ROOT / "logs"
NOTE_OUTPUT_DIR = DEFAULT_OUTPUT_ROOT / "notes"
LOG_PATH = LOG_OUTPUT_DIR / "feature_extraction.log"
EMBEDDING_ARRAY_PATH = FEATURE_OUTPUT_DIR / "embeddings.npy"
EMBEDDING_CSV_PATH = FEATURE_OUTPUT_DIR / "embeddings.csv"
METADATA_PATH = FEATURE_OUTPUT_DIR / "metadata.json"
SUMMARY_NOTE_PATH = NOTE_OUTPUT_DIR / "feature_summary.md"

IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]
TARGET_RESIZE = 256
TARGET_CROP = 224
BATCH_SIZE = 32
NEIGHBOR_SAMPLE = 8
RNG_SEED = 42

LABELED_BUCKET = "avec_labels"
UNLABELED_BUCKET = "sans_label"

BACKBONE_NAME = "torchvision.resnet18"
BACKBONE_WEIGHTS = "ResNet18_Weights.IMAGENET1K_V1"
BACKBONE_LAYER = "global_avg_pool"

WEIGHTS_ENV = "SSIP_B200_WEIGHTS"
PRECISION_ENV = "SSIP_B200_PRECISION"
GRAY_CARRIAGE_ENV = "SSIP_B200_GRAY_CARRIAGE"  # "1": ship R==G==B files as one plane
DECODE_THREADS_ENV = "SSIP_B200_DECODE_THREADS"
SINGLE_GPU_ENV = "SSIP_B200_SINGLE_GPU"  # "1": `--device cuda` stays on the current GPU even if several are visible
ALL_RANKS_HOST_ENV = "SSIP_B200_ALL_RANKS_HOST"  # "1": under torchrun every rank (not only rank 0) gets the host matrix
DECODE_MODE_ENV = "SSIP_B200_DECODE"  # "process" | "thread" | "nvjpeg" | "auto" (default: worker processes from 512 files up)
JPEG_BACKEND_ENV = "SSIP_B200_NVJPEG_BACKEND"  # "auto" | "hardware" | "gpu": which nvJPEG decoder SSIP_B200_DECODE=nvjpeg uses


@dataclass(frozen=True)
class ImageRecord:
    """One dataset file (fields as src/feature_extraction.py:85-92)."""

    absolute_path: Path
    relative_path: Path
    bucket: str
    label: Optional[str]


@dataclass
class ExtractionResults:
    """What extract_embeddings returns (fields as src/feature_extraction.py:95-102)."""

    embeddings: np.ndarray  # host fp32 [N_ok,512]; under torchrun on rank 0 only (None elsewhere, see ALL_RANKS_HOST_ENV)
    records: List[ImageRecord]
    failures: List[Path]
    per_file_times: List[float]
    # extra (defaulted, so the reference's four-field construction still works): the same matrix still on the GPU
    # when it was assembled there (multi-GPU all-gather), for the on-device post-processing below
    device_embeddings: Optional[torch.Tensor] = None


# --- logging / discovery -------------------------------------------------------------------------


def configure_logging(verbose: bool = False) -> None:
    """File (rewritten every run) + stderr, reference format (src/feature_extraction.py:110-122)."""
    LOG_OUTPUT_DIR.mkdir(parents=True, exist_ok=True)
    logging.basicConfig(
        level=logging.DEBUG if verbose else logging.INFO,
        format="%(asctime)s [%(levelname)s] %(message)s",
        handlers=[logging.FileHandler(LOG_PATH, mode="w", encoding="utf-8"), logging.StreamHandler()],
    )


def _files_under(root: Path) -> List[Path]:
    return [p for p in sorted(root.rglob("*")) if p.is_file()]


def discover_image_records(data_dir: Path) -> List[ImageRecord]:
    """Deterministic record list: labeled (sorted label dirs, sorted files) then unlabeled.

    Ordering, bucket names and error behaviour follow src/feature_extraction.py:125-181: every
    regular file is a record (non-images become decode failures later), a missing bucket is a
    warning, a missing directory FileNotFoundError, an empty dataset RuntimeError.
    """
    data_dir = Path(data_dir)
    if not data_dir.exists():
        raise FileNotFoundError(f"Data directory not found: {data_dir}")
    found: List[ImageRecord] = []
    labeled_root = data_dir / LABELED_BUCKET
    if labeled_root.exists():
        for label_dir in sorted(d for d in labeled_root.iterdir() if d.is_dir()):
            found += [ImageRecord(f, f.relative_to(data_dir), "labeled", label_dir.name) for f in _files_under(label_dir)]
    else:
        logging.warning("Labeled bucket missing at %s", labeled_root)
    n_labeled = len(found)
    unlabeled_root = data_dir / UNLABELED_BUCKET
    if unlabeled_root.exists():
        found += [ImageRecord(f, f.relative_to(data_dir), "unlabeled", None) for f in _files_under(unlabeled_root)]
    else:
        logging.warning("Unlabeled bucket missing at %s", unlabeled_root)
    if not found:
        raise RuntimeError(f"No image files discovered under {data_dir}")
    logging.info("Discovered %d images (labeled=%d, unlabeled=%d)", len(found), n_labeled, len(found) - n_labeled)
    return found


# --- engine management ---------------------------------------------------------------------------

_ENGINES: Dict[Tuple[int, str, str], Engine] = {}


def _cuda_index(device: torch.device) -> int:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(
            f"device '{device}' requested, but ssip_b200 only runs on a B200 GPU (no CPU fallback); "
            "use the reference's src.feature_extraction for CPU runs"
        )
    if device.index is not None:
        return device.index
    _, size, local_rank = fxdist.env_world()
    return local_rank if size > 1 else torch.cuda.current_device()


def _seeded_backbone(seed: int, randomize_bn: bool):
    """torchvision resnet18 with a seeded random init (the offline stand-in for the download)."""
    from torchvision import models

    torch.manual_seed(seed)
    net = models.resnet18(weights=None)
    if randomize_bn:
        gen = torch.Generator().manual_seed(seed + 1)
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                c = mod.num_features
                mod.weight.data = 0.8 + 0.4 * torch.rand(c, generator=gen)
                mod.bias.data = 0.1 * torch.randn(c, generator=gen)
                mod.running_mean.data = 0.1 * torch.randn(c, generator=gen)
                mod.running_var.data = 0.6 + 0.8 * torch.rand(c, generator=gen)
    return net


def resolve_state_dict(spec: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """State dict of the backbone: IMAGENET1K_V1 unless SSIP_B200_WEIGHTS says otherwise."""
    spec = spec if spec is not None else os.environ.get(WEIGHTS_ENV, "")
    if spec.startswith("random"):
        kind, _, seed = spec.partition(":")
        return _seeded_backbone(int(seed or "1234"), kind == "random-bn").state_dict()
    if spec:
        state = torch.load(spec, map_location="cpu", weights_only=True)
        return state.get("state_dict", state) if isinstance(state, dict) else state
    from torchvision import models

    return models.resnet18(weights=models.ResNet18_Weights.IMAGENET1K_V1).state_dict()


_NO_WEIGHTS: Dict[str, torch.Tensor] = {}  # sentinel for get_engine: the caller needs no trunk (post-processing kernels only)


def get_engine(device: torch.device, min_batch: int = BATCH_SIZE, state_dict: Optional[Dict[str, torch.Tensor]] = None,
               precision: Optional[str] = None) -> Engine:
    """Engine for `device`, created (and its weights loaded) on first use."""
    index = _cuda_index(device)
    if state_dict is _NO_WEIGHTS:  # post-processing only: any engine of this GPU will do, a bare one otherwise
        for key, eng in _ENGINES.items():
            if key[0] == index:
                return eng
        eng = _ENGINES[(index, "bare", "")] = Engine(index, max(min_batch, 1), "bf16")
        return eng
    precision = precision or os.environ.get(PRECISION_ENV, "bf16")
    key = (index, precision, os.environ.get(WEIGHTS_ENV, "") if state_dict is None else f"id{id(state_dict)}")
    eng = _ENGINES.get(key)
    if eng is not None and eng.max_batch < min_batch:
        eng.close()
        eng = None
    if eng is None:
        eng = Engine(index, max(min_batch, 1), precision)
        eng.load_state_dict(state_dict if state_dict is not None else resolve_state_dict())
        _ENGINES[key] = eng
    return eng


# --- transform / model: same call shapes as the reference ----------------------------------------


def _check_mode(mode: str, bands: int) -> None:
    """The reference never converts modes (src/feature_extraction.py:233-240): anything that is not three 8-bit
    bands makes ToTensor / Normalize raise.  Same exception types here."""
    if mode in ("I", "I;16", "I;16L", "I;16B", "F"):
        raise TypeError(f"Input tensor should be a float tensor after ToTensor; mode {mode!r} images are not 8-bit")
    if bands != 3:
        raise RuntimeError(
            f"output with shape [{bands}, {TARGET_CROP}, {TARGET_CROP}] doesn't match the broadcast shape "
            f"[3, {TARGET_CROP}, {TARGET_CROP}] (image mode {mode!r}; the pipeline does no RGB conversion)"
        )


def _decoded_array(img: Image.Image) -> np.ndarray:
    """The HWC uint8 array ToTensor would see, with the reference's failure modes."""
    _check_mode(img.mode, len(img.getbands()))
    arr = np.asarray(img)
    if arr.dtype != np.uint8:
        raise TypeError(f"unsupported pixel type {arr.dtype} for mode {img.mode!r}")
    return arr


class _CudaTransform:
    """Callable returned by build_transform(): PIL image -> fp32 [3,224,224] (CPU tensor), computed
    by the fused CUDA kernel, bit-identical to the reference's torchvision Compose."""

    def __init__(self, device: Optional[torch.device] = None):
        self._device = torch.device(device) if device is not None else torch.device("cuda")

    def batch(self, arrays: Sequence[np.ndarray]) -> torch.Tensor:
        eng = get_engine(self._device, min_batch=max(len(arrays), 1))
        buf, descs, _ = pack_images(arrays)
        with torch.cuda.device(eng.device):
            dev = torch.from_numpy(buf).to(eng.device)
            return eng.preprocess_nchw(dev, descs, len(arrays))

    def __call__(self, img: Image.Image) -> torch.Tensor:
        return self.batch([_decoded_array(host_resize_if_oversized(img))])[0].cpu()

    def __repr__(self) -> str:
        return (f"FusedCudaTransform(Resize({TARGET_RESIZE}), CenterCrop({TARGET_CROP}), ToTensor(), "
                f"Normalize(mean={IMAGENET_MEAN}, std={IMAGENET_STD}))")


def build_transform() -> Callable[[Image.Image], torch.Tensor]:
    """Deterministic preprocessing (src/feature_extraction.py:184-207) as one CUDA kernel."""
    return _CudaTransform()


class _CudaTrunk(torch.nn.Module):
    """Callable returned by load_model(): [B,3,224,224] -> [B,512,1,1], frozen, eval-only."""

    def __init__(self, device: torch.device):
        super().__init__()
        self._device = torch.device(device)
        self.eval()

    def forward(self, batch: torch.Tensor) -> torch.Tensor:
        eng = get_engine(self._device, min_batch=max(int(batch.shape[0]), 1))
        x = batch.detach().to(eng.device, torch.float32).contiguous()
        out = torch.empty((x.shape[0], 512), dtype=torch.float32, device=eng.device)
        with torch.cuda.device(eng.device):
            for s in range(0, x.shape[0], eng.max_batch):
                out[s : s + eng.max_batch] = eng.forward_nchw(x[s : s + eng.max_batch])
        return out.view(-1, 512, 1, 1)


def load_model(device: torch.device) -> torch.nn.Module:
    """Frozen ResNet-18 minus fc (src/feature_extraction.py:210-227) on the CUDA engine."""
    get_engine(device)  # fails here, loudly, if there is no B200 / no weights
    return _CudaTrunk(device)


def preprocess_image(path: Path, transform: Callable[[Image.Image], torch.Tensor]) -> torch.Tensor:
    """Open one file and apply the transform; no mode conversion (src/feature_extraction.py:233-240)."""
    with Image.open(path) as img:
        return transform(img)


def batched(iterable: Sequence, batch_size: int) -> Iterable[Sequence]:
    """Consecutive slices of at most batch_size items (src/feature_extraction.py:243-248)."""
    for lo in range(0, len(iterable), batch_size):
        yield iterable[lo : lo + batch_size]


# --- the hot loop --------------------------------------------------------------------------------


def _load_file(path: Path):
    """Decode one file on a pool thread -> HWC uint8 array, or the exception to report."""
    try:
        with Image.open(path) as img:
            arr = _decoded_array(host_resize_if_oversized(img))
            if os.environ.get(GRAY_CARRIAGE_ENV) == "1" and (arr[..., 0] == arr[..., 1]).all() and (arr[..., 1] == arr[..., 2]).all():
                arr = np.ascontiguousarray(arr[..., 0])
            return arr
    except (UnidentifiedImageError, OSError) as exc:  # the two the reference tolerates (:281)
        return exc


_DECODE_POOL: Optional[DecodePool] = None


def _process_pool(workers: int, slots: int) -> DecodePool:
    """The worker processes are started once per interpreter (about 2 s) and reused by later calls."""
    global _DECODE_POOL
    if _DECODE_POOL is None or _DECODE_POOL.workers != workers:
        if _DECODE_POOL is not None:
            _DECODE_POOL.close()
        _DECODE_POOL = DecodePool(workers, slots)
        atexit.register(_DECODE_POOL.close)
    return _DECODE_POOL


def _pin(t: torch.Tensor) -> torch.Tensor:
    """Page-lock a shared-memory staging tensor so that its H2D copy is an asynchronous DMA (best effort)."""
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel(), 0)
    t._fx_pinned = int(rc) == 0  # type: ignore[attr-defined]
    if not t._fx_pinned:  # type: ignore[attr-defined]
        logging.warning("cudaHostRegister failed (%s): staging buffer stays pageable, copies will be slower", rc)
    return t


def _unpin(t: Optional[torch.Tensor]) -> None:
    if t is not None and getattr(t, "_fx_pinned", False):
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
        t._fx_pinned = False  # type: ignore[attr-defined]


def _extract_local(records: Sequence[ImageRecord], eng: Engine, batch_size: int, sink: Optional[torch.Tensor] = None):
    """Single-GPU pass over `records`: embeddings + bookkeeping.  With `sink` (a contiguous fp32 CUDA view with room for
    [len(records), 512]: this rank's slot of the all-gather buffer) the rows stay on the device -- the trunk's last
    kernel writes row i of the kept records at sink[i] -- and the returned matrix is None.

    Pipeline slots (fx_embed_host_async, N.HOST_SLOTS of them): while the GPU works on earlier batches (two at a
    time, one per engine lane), the host decodes the next one into a free pinned staging buffer and its H2D copy
    overlaps their kernels.
    """
    kept: List[int] = []
    failures: List[Path] = []
    times: List[float] = []
    blocks: List[np.ndarray] = []
    threads = int(os.environ.get(DECODE_THREADS_ENV, "0")) or min(32, (os.cpu_count() or 8))
    nslots = N.HOST_SLOTS
    staging: List[Optional[torch.Tensor]] = [None] * nslots
    outs = [torch.empty((batch_size, 512), dtype=torch.float32).pin_memory() for _ in range(nslots)] if sink is None else []
    written = 0  # rows already placed in `sink`
    pending: List[Optional[Tuple[List[int], int]]] = [None] * nslots  # per slot: (record indices, n)
    t_last = time.perf_counter()

    def finish(slot: int) -> None:
        nonlocal t_last
        if pending[slot] is None:
            return
        ok, n = pending[slot]
        eng.embed_host_wait(slot)
        if sink is None:
            blocks.append(outs[slot][:n].numpy().copy())
        kept.extend(ok)
        now = time.perf_counter()
        times.extend([(now - t_last) / n] * n)  # batch wall time / successes, as src/feature_extraction.py:297-300
        t_last = now
        pending[slot] = None

    mode = os.environ.get(DECODE_MODE_ENV, "auto")
    use_nvjpeg = mode == "nvjpeg"
    if use_nvjpeg:
        backend = eng.jpeg_init(os.environ.get(JPEG_BACKEND_ENV, "auto"))  # FxError if nvJPEG cannot be had: no silent fallback
        logging.info("JPEG files are decoded on the GPU (nvJPEG, %s backend); everything else by Pillow on the host", backend)
    use_procs = mode == "process" or (mode == "auto" and len(records) >= 512)
    gray = os.environ.get(GRAY_CARRIAGE_ENV) == "1"

    def fail(rec: ImageRecord, text: str) -> None:
        logging.error("Failed to decode %s: %s", rec.absolute_path, text)
        failures.append(rec.absolute_path)

    def stage_with_threads(pool, chunk, lo, slot):
        """Decode on threads, then pack into the slot's pinned buffer -> (descs, n, total bytes, record indices)."""
        decoded = list(pool.map(_load_file, [r.absolute_path for r in chunk]))
        arrays, ok = [], []
        for off, (rec, item) in enumerate(zip(chunk, decoded)):
            if isinstance(item, BaseException):
                fail(rec, str(item))
            else:
                arrays.append(item)
                ok.append(lo + off)
        if not arrays:
            return None
        finish(slot)  # the slot's buffers are free again once its previous batch is out
        need = sum((a.size + 255) // 256 * 256 for a in arrays)
        if staging[slot] is None or staging[slot].numel() < need:
            staging[slot] = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8).pin_memory()
        _, descs, total = pack_images(arrays, out=staging[slot].numpy())
        return descs, len(arrays), total, ok

    def stage_with_nvjpeg(pool, chunk, lo, slot):
        """SURVEY.md 8f rank 1, GPU option: the library reads the batch's files into the slot's page-locked bitstream
        buffer (native threads) and nvJPEG decodes the baseline RGB JPEGs among them straight into the device image buffer;
        every other file (and every JPEG the frame header or the missing end-of-image marker makes doubtful) takes the
        reference's own Pillow call on a pool thread, with the reference's failure handling."""
        finish(slot)  # the slot's bitstream buffer and decoder state are free once its previous batch is out
        info = eng.jpeg_read_files(slot, [str(r.absolute_path) for r in chunk])
        on_gpu = [info[k].status == N.FILE_GPU_JPEG and min(info[k].height, info[k].width) <= HOST_RESIZE_SHORT_SIDE for k in range(len(chunk))]
        host_idx = [k for k, g in enumerate(on_gpu) if not g]
        host = dict(zip(host_idx, pool.map(_load_file, [chunk[k].absolute_path for k in host_idx]))) if host_idx else {}
        take, pixels, ok, shapes = [], [], [], []
        for k, rec in enumerate(chunk):
            if on_gpu[k]:
                take.append(k)
                pixels.append(None)
                shapes.append((info[k].height, info[k].width, 3))
            else:
                item = host[k]
                if isinstance(item, BaseException):
                    fail(rec, str(item))
                    continue
                take.append(k)
                pixels.append(np.ascontiguousarray(item))
                shapes.append((item.shape[0], item.shape[1], 1 if item.ndim == 2 else item.shape[2]))
            ok.append(lo + k)
        if not take:
            return None
        n = len(take)
        sel = (N.FileInfo * n)()
        descs = (N.ImageDesc * n)()
        off = 0
        for j, (k, (h, w, c)) in enumerate(zip(take, shapes)):
            sel[j] = info[k]
            if pixels[j] is not None:
                sel[j].status = N.FILE_HOST_DECODE
            descs[j].offset, descs[j].height, descs[j].width, descs[j].channels = off, h, w, c
            off += (h * w * c + 255) // 256 * 256
        return (sel, pixels), descs, n, off, ok

    def stage_with_processes(pool: DecodePool, chunk, lo, slot):
        """Header pass -> layout -> worker processes decode straight into the slot's shared, page-locked buffer."""
        metas = pool.probe([str(r.absolute_path) for r in chunk])
        jobs, ok, off = [], [], 0
        for k, (rec, m) in enumerate(zip(chunk, metas)):
            if isinstance(m[0], str):  # failure triple (kind, exception name, text)
                if m[0] != "decode":  # not one of the two the reference tolerates: same exception class as thread mode
                    raise rebuild_exception(m[1], m[2])
                fail(rec, m[2])
                continue
            h, w, bands, pil_mode = m
            _check_mode(pil_mode, bands)
            jobs.append((str(rec.absolute_path), off, h, w, bands, gray))
            ok.append(lo + k)
            off += (h * w * bands + 255) // 256 * 256
        if not jobs:
            return None
        finish(slot)
        if staging[slot] is None or staging[slot].numel() < off:
            _unpin(staging[slot])
            staging[slot] = None
            shm, _ = pool.buffer(slot, off)
            staging[slot] = _pin(torch.frombuffer(shm.buf, dtype=torch.uint8))
        done = pool.decode(slot, jobs)
        descs = (N.ImageDesc * len(jobs))()
        n, kept_idx = 0, []
        for job, idx, res in zip(jobs, ok, done):
            if not isinstance(res, int):  # pixel data broken although the header parsed
                if res[0] != "decode":
                    raise rebuild_exception(res[1], res[2])
                fail(records[idx], res[2])
                continue
            descs[n].offset, descs[n].height, descs[n].width, descs[n].channels = job[1], job[2], job[3], res
            kept_idx.append(idx)
            n += 1
        return (descs, n, off, kept_idx) if n else None

    use_procs = use_procs and not use_nvjpeg
    pool_cm = _process_pool(threads, nslots) if use_procs else ThreadPoolExecutor(max_workers=threads)
    stage = stage_with_nvjpeg if use_nvjpeg else (stage_with_processes if use_procs else stage_with_threads)

    def submit(slot, descs, n_ok, total, files=None):
        nonlocal written
        out = outs[slot] if sink is None else sink[written : written + n_ok]
        if files is not None:
            eng.embed_files_async(slot, files[0], files[1], descs, n_ok, total, out)
        elif sink is None:
            eng.embed_host_async(slot, staging[slot], descs, n_ok, total, out)
        else:
            eng.embed_host_async_dev(slot, staging[slot], descs, n_ok, total, out)
        written += n_ok

    try:
        with torch.cuda.device(eng.device):
            slot = 0
            for lo in range(0, len(records), batch_size):
                chunk = records[lo : lo + batch_size]
                mark = len(failures)
                staged = stage(pool_cm, chunk, lo, slot)
                if staged is None:
                    continue
                if use_nvjpeg:
                    files, descs, n_ok, total, ok = staged
                    try:
                        submit(slot, descs, n_ok, total, files)
                    except N.FxError as exc:
                        if exc.status != N.FX_ERR_UNSUPPORTED:
                            raise
                        # nvJPEG turned the batch down (a bitstream the header walk could not fault; nothing was queued):
                        # redo the whole batch with Pillow on the host, bookkeeping included
                        logging.warning("nvJPEG rejected a batch (%s): decoding it on the host", exc.text)
                        del failures[mark:]
                        staged = stage_with_threads(pool_cm, chunk, lo, slot)
                        if staged is None:
                            continue
                        descs, n_ok, total, ok = staged
                        submit(slot, descs, n_ok, total)
                else:
                    descs, n_ok, total, ok = staged
                    submit(slot, descs, n_ok, total)
                pending[slot] = (ok, n_ok)
                slot = (slot + 1) % nslots
            for k in range(nslots):  # oldest first
                finish((slot + k) % nslots)
    finally:
        # an exception above (channel-policy error, worker error, unsupported geometry) can leave up to HOST_SLOTS
        # batches in flight: their DMA reads staging[slot] and writes outs[slot], so drain before releasing either
        for k in range(nslots):
            try:
                eng.embed_host_wait(k)
            except Exception:  # noqa: BLE001 - the original exception is the one to report
                pass
        if use_procs:
            for k in range(nslots):
                _unpin(staging[k])
                staging[k] = None
                pool_cm.release(k)
        else:
            pool_cm.shutdown(wait=True)
    if sink is not None:  # batches were submitted, and their rows placed, in record order
        assert kept == sorted(kept) and len(kept) == written
        return None, kept, failures, times
    order = np.argsort(np.asarray(kept, dtype=np.int64), kind="stable") if kept else np.zeros(0, np.int64)
    local = np.concatenate(blocks, axis=0)[order] if blocks else np.empty((0, 512), np.float32)
    kept = [kept[i] for i in order]
    times = [times[i] for i in order]
    return torch.from_numpy(local), kept, failures, times


def extract_embeddings(records: List[ImageRecord], device: torch.device, batch_size: int = BATCH_SIZE) -> ExtractionResults:
    """Feature extraction over `records` (signature of src/feature_extraction.py:251-255).

    With WORLD_SIZE > 1 (the workers `main` starts for ``--device cuda`` on a multi-GPU box, or torchrun) the records
    are sharded contiguously over the ranks, each rank runs its shard on its own GPU writing its rows straight into its
    slot of the gather buffer, and ONE in-place NCCL all-gather assembles [N,512] on every GPU (`device_embeddings`);
    the host matrix `embeddings` is materialised on rank 0.  Row i belongs to `records[i]` of the returned (kept) list.
    """
    eng = get_engine(device, min_batch=batch_size)
    logging.info("Beginning feature extraction over %d records", len(records))
    distributed = fxdist.ensure_process_group("nccl")
    if distributed:
        fxdist.bind_to_gpu_numa_node(eng.device_index)  # one rank per GPU: stay on that GPU's socket
    rank, size = fxdist.world()
    lo, hi = fxdist.shard_bounds(len(records), rank, size) if distributed else (0, len(records))
    error: Optional[BaseException] = None
    gather_buf = sink = None
    cap = 0
    if distributed:
        # one [R * ceil(N/R), 512] buffer per GPU; this rank's rows are written into its slot by the trunk
        # itself, one in-place NCCL all-gather assembles the matrix, and only rank 0 copies it to the host
        cap = fxdist.shard_bounds(len(records), 0, size)[1]
        with torch.cuda.device(eng.device):
            gather_buf = torch.empty((size * cap, 512), dtype=torch.float32, device=eng.device)
        sink = gather_buf[rank * cap : (rank + 1) * cap]
    try:
        local, kept, failures, times = _extract_local(records[lo:hi], eng, batch_size, sink=sink)
    except Exception as exc:  # noqa: BLE001 - re-raised below, after the other ranks have been told
        if not distributed:
            raise
        error = exc
    if distributed:
        # a rank that failed (e.g. a channel-policy error in its shard) must not leave its peers waiting in the
        # all-gather: exchange the outcome first, then every rank raises
        outcomes = fxdist.allgather_objects(None if error is None else f"{type(error).__name__}: {error}")
        if error is not None:
            raise error
        bad = [(r, o) for r, o in enumerate(outcomes) if o is not None]
        if bad:
            raise RuntimeError(f"feature extraction failed on rank {bad[0][0]}: {bad[0][1]}")
    kept = [lo + k for k in kept]
    if distributed:
        with torch.cuda.device(eng.device):
            full, _ = fxdist.allgather_inplace(gather_buf, cap, len(kept))
        meta = fxdist.allgather_objects((kept, failures, times))
        kept = fxdist.concat_in_rank_order([m[0] for m in meta])
        failures = fxdist.concat_in_rank_order([m[1] for m in meta])
        times = fxdist.concat_in_rank_order([m[2] for m in meta])
    else:
        full = local
    if full.shape[0] == 0:
        raise RuntimeError("No embeddings were generated; all images failed to decode?")
    # the host copy of the gathered matrix is rank 0's job (it writes the artifacts); the other ranks keep the device
    # tensor only, unless SSIP_B200_ALL_RANKS_HOST=1 asks for the reference's "numpy on every caller" everywhere
    want_host = not distributed or rank == 0 or os.environ.get(ALL_RANKS_HOST_ENV) == "1"
    matrix = full.cpu().numpy() if want_host else None
    logging.info("Computed embeddings with shape %s", tuple(full.shape))
    return ExtractionResults(embeddings=matrix, records=[records[i] for i in kept], failures=failures, per_file_times=times,
                             device_embeddings=full if full.is_cuda else None)


# --- post-processing and artifacts ---------------------------------------------------------------


def compute_dataset_digest(records: Sequence[ImageRecord]) -> str:
    """sha256 over (relative path, size, int(mtime)) in path order (src/feature_extraction.py:316-331)."""
    return _artifacts.dataset_digest(records)


def run_sanity_checks(embeddings) -> Dict[str, float]:
    """NaN/Inf guard + summary statistics (src/feature_extraction.py:334-356).

    A numpy matrix is checked with the reference's own numpy expressions; a CUDA tensor is checked where it lives
    (fx_column_stats: one pass for the counts and the column sums, one for the deviations, fp64 accumulation)."""
    if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda:
        eng = get_engine(embeddings.device, min_batch=1, state_dict=_NO_WEIGHTS)
        with torch.cuda.device(embeddings.device):
            st, _, _, _ = eng.column_stats(embeddings.contiguous())
        if st["nan_count"]:
            raise ValueError("Embedding matrix contains NaN values")
        if st["inf_count"]:
            raise ValueError("Embedding matrix contains inf values")
        stats = {"num_vectors": int(embeddings.shape[0]), "dimension": int(embeddings.shape[1]),
                 "mean_abs_mean": st["mean_abs_mean"], "mean_std": st["mean_std"]}
    else:
        if np.isnan(embeddings).any():
            raise ValueError("Embedding matrix contains NaN values")
        if np.isinf(embeddings).any():
            raise ValueError("Embedding matrix contains inf values")
        stats = {
            "num_vectors": int(embeddings.shape[0]),
            "dimension": int(embeddings.shape[1]),
            "mean_abs_mean": float(np.abs(embeddings.mean(axis=0)).mean()),
            "mean_std": float(embeddings.std(axis=0).mean()),
        }
    logging.info("Embedding stats — vectors: %d, dim: %d, mean(|mean|): %.5f, mean(std): %.5f",
                 stats["num_vectors"], stats["dimension"], stats["mean_abs_mean"], stats["mean_std"])
    return stats


def nearest_neighbor_probe(embeddings, records: Sequence[ImageRecord], sample_size: int = NEIGHBOR_SAMPLE,
                           seed: int = RNG_SEED) -> List[Dict[str, object]]:
    """Cosine nearest neighbour of a few seeded queries (src/feature_extraction.py:359-398); numpy matrix -> numpy as
    the reference, CUDA tensor -> fx_neighbor_probe (same sampled rows, same tie rule)."""
    n = embeddings.shape[0]
    k = min(sample_size, n - 1)
    if n < 2 or k <= 0:
        return []
    queries = np.random.default_rng(seed).choice(n, size=k, replace=False)
    probe = []
    if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda:
        eng = get_engine(embeddings.device, min_batch=1, state_dict=_NO_WEIGHTS)
        with torch.cuda.device(embeddings.device):
            rows, sims = eng.neighbor_probe(embeddings.contiguous(), queries)
        for q, best, sim in zip(queries, rows, sims):
            probe.append({"query": str(records[q].relative_path), "neighbor": str(records[int(best)].relative_path),
                          "similarity": float(sim)})
    else:
        unit = embeddings / np.clip(np.linalg.norm(embeddings, axis=1, keepdims=True), a_min=1e-12, a_max=None)
        for q in queries:
            sims = unit[q] @ unit.T
            sims[q] = -np.inf
            best = int(np.argmax(sims))
            probe.append({"query": str(records[q].relative_path), "neighbor": str(records[best].relative_path),
                          "similarity": float(sims[best])})
    logging.info("Nearest-neighbor probe completed for %d samples", len(probe))
    return probe


def _summary_markdown(results: ExtractionResults, stats: Dict[str, float], probe: List[Dict[str, object]], device) -> str:
    lat = results.per_file_times
    mean_lat = float(np.mean(lat)) if lat else float("nan")
    med_lat = float(np.median(lat)) if lat else float("nan")
    if probe:
        table = ["| Query | Neighbor | Cosine |", "| --- | --- | --- |"]
        table += [f"| {p['query']} | {p['neighbor']} | {p['similarity']:.4f} |" for p in probe]
        neighbors = "\n".join(table)
    else:
        neighbors = "No neighbors computed (insufficient samples)."
    failed = "\n".join(f"- {p}" for p in results.failures) if results.failures else "None"
    dim = results.embeddings.shape[1]
    lines = [
        "# Feature Extraction Summary",
        "",
        f"- Backbone: {BACKBONE_NAME} ({BACKBONE_WEIGHTS})",
        f"- Layer: global average pooled features ({dim}-D)",
        f"- Input spec: resize {TARGET_RESIZE} → center crop {TARGET_CROP}, ImageNet normalization",
        f"- Batch size: {BATCH_SIZE}",  # the reference prints the constant, not the CLI value (:481)
        f"- Device: {device}",
        f"- Total images processed: {results.embeddings.shape[0]}",
        f"- Failed decodes: {len(results.failures)}",
        f"- Mean per-image latency (s): {mean_lat:.4f}",
        f"- Median per-image latency (s): {med_lat:.4f}",
        "",
        "## Sanity Check Statistics",
        "",
        f"- Mean of |dimension means|: {stats['mean_abs_mean']:.6f}",
        f"- Mean of dimension standard deviations: {stats['mean_std']:.6f}",
        "",
        "## Nearest Neighbor Spot Check",
        "",
        neighbors,
        "",
        "## Decode Failures",
        "",
        failed,
        "",
    ]
    return "\n".join(lines)


def save_artifacts(results: ExtractionResults, stats: Dict[str, float], neighbor_probe: List[Dict[str, object]], data_dir: Path,
                   device: torch.device) -> None:
    """embeddings.npy / embeddings.csv / metadata.json / feature_summary.md (src/feature_extraction.py:401-502)."""
    FEATURE_OUTPUT_DIR.mkdir(parents=True, exist_ok=True)
    NOTE_OUTPUT_DIR.mkdir(parents=True, exist_ok=True)
    # same bytes as np.save(astype(float32)) / DataFrame.to_csv(index=False), without the second 2 GB buffer and the
    # DataFrame of the reference; the .npy stream, the CSV and the dataset digest run side by side
    digest = _artifacts.save_parallel(EMBEDDING_ARRAY_PATH, EMBEDDING_CSV_PATH, results.embeddings, results.records)
    metadata = {
        "backbone": BACKBONE_NAME,
        "weights": BACKBONE_WEIGHTS,
        "layer": BACKBONE_LAYER,
        "embedding_dimension": int(results.embeddings.shape[1]),
        "input_resize": TARGET_RESIZE,
        "input_crop": TARGET_CROP,
        "normalization_mean": IMAGENET_MEAN,
        "normalization_std": IMAGENET_STD,
        "channel_policy": "No conversion (assumes RGB inputs)",
        "date_utc": datetime.now(timezone.utc).isoformat(),
        "num_images": int(results.embeddings.shape[0]),
        "failed_images": len(results.failures),
        "device": str(device),
        "dataset_dir": str(data_dir),
        "dataset_digest": digest,
        "sanity_checks": stats,
        "neighbor_probe": neighbor_probe,
    }
    with METADATA_PATH.open("w", encoding="utf-8") as fh:
        json.dump(metadata, fh, indent=2)
    SUMMARY_NOTE_PATH.write_text(_summary_markdown(results, stats, neighbor_probe, device), encoding="utf-8")


# --- CLI -----------------------------------------------------------------------------------------


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description="Extract CNN embeddings for the MRI dataset")
    parser.add_argument("--data-dir", type=Path, default=DEFAULT_DATA_DIR, help="Root directory containing 'avec_labels' and 'sans_label'")
    parser.add_argument("--device", type=str, default="cuda" if torch.cuda.is_available() else "cpu",
                        help="Torch device to use (default: cuda if available else cpu)")
    parser.add_argument("--batch-size", type=int, default=BATCH_SIZE, help="Mini-batch size for inference")
    parser.add_argument("--verbose", action="store_true", help="Enable verbose logging")
    return parser.parse_args(argv)


def _fan_out(argv: Optional[Sequence[str]], n_gpus: int) -> int:
    """``--device cuda`` on a box with several visible GPUs = all of them (the reference's flag set is
    unchanged, CUDA_VISIBLE_DEVICES restricts the set): start one worker process per GPU with the torchrun environment
    (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*), each running this module's CLI with the same arguments; rank 0 writes the
    artifacts.  The parent never touches CUDA.  Returns the first non-zero worker exit code (0 if all succeeded)."""
    import socket
    import subprocess
    import sys

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:  # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    args = list(sys.argv[1:] if argv is None else argv)
    root = str(Path(__file__).resolve().parent.parent)  # where the `ssip_b200` alias lives
    pythonpath = os.pathsep.join(x for x in (root, os.environ.get("PYTHONPATH", "")) if x)
    procs = []
    for r in range(n_gpus):
        env = dict(os.environ, PYTHONPATH=pythonpath, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(n_gpus), LOCAL_WORLD_SIZE=str(n_gpus),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-m", "ssip_b200.feature_extraction", *map(str, args)], env=env))
    # wait for all of them; if one fails, stop the others (they would otherwise sit in the collective until NCCL times out)
    failed = 0
    while True:
        codes = [p.poll() for p in procs]
        failed = next((c for c in codes if c), 0)
        if failed or all(c is not None for c in codes):
            break
        time.sleep(0.2)
    if failed:
        for p in procs:
            if p.poll() is None:
                p.terminate()
        for p in procs:
            try:
                p.wait(timeout=10)
            except Exception:  # noqa: BLE001
                p.kill()
    return failed


def main(argv: Optional[Sequence[str]] = None) -> None:
    args = parse_args(argv)
    rank, size, _ = fxdist.env_world()
    if (size == 1 and args.device == "cuda" and os.environ.get(SINGLE_GPU_ENV) != "1" and torch.cuda.is_available()
            and torch.cuda.device_count() > 1):
        rc = _fan_out(argv, torch.cuda.device_count())
        if rc:
            raise SystemExit(rc)
        return
    if rank == 0:
        configure_logging(verbose=args.verbose)
    else:  # only rank 0 owns the log file and the artifacts
        logging.basicConfig(level=logging.WARNING)
    device = torch.device(args.device)
    logging.info("Starting feature extraction on device %s", device)
    records = discover_image_records(args.data_dir)
    t0 = time.perf_counter()
    results = extract_embeddings(records, device=device, batch_size=args.batch_size)
    logging.info("Completed embedding extraction in %.2f seconds", time.perf_counter() - t0)
    if rank == 0:
        # large gathered matrices are post-processed where they already are (SSIP_B200_DEVICE_POSTPROCESS = 0 | 1 | auto)
        mode = os.environ.get("SSIP_B200_DEVICE_POSTPROCESS", "auto")
        on_device = results.device_embeddings is not None and (mode == "1" or (mode == "auto" and results.embeddings.shape[0] >= 100_000))
        matrix = results.device_embeddings if on_device else results.embeddings
        stats = run_sanity_checks(matrix)
        probe = nearest_neighbor_probe(matrix, results.records)
        save_artifacts(results, stats, probe, args.data_dir, device)
        logging.info("Artifacts saved to %s", FEATURE_OUTPUT_DIR)
    if size > 1 and torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
