"""Real-MRI parity fixture minted from the REAL reference module.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_mri.py

Copies the first N_LABELED files of each label directory and the first N_UNLABELED unlabeled files of the reference's
own dataset (``/root/reference/mri_dataset_brain_cancer_oc``, 512x512 RGB JPEGs, AFL-3.0) into
``tests/golden/mri_real/`` with the dataset's directory layout, runs the unmodified
``src.feature_extraction.extract_embeddings`` on that directory (CPU fp32; only the weight download is shimmed, exactly
as make_golden.py does) and writes ``tests/golden/mri_real_golden.npz``:

  paths            relative paths in discover_image_records order (= row order)
  pixels_sha256    sha256 of the decoded HWC uint8 array of every file (pins the JPEG decode on the other box)
  pre_sha256       sha256 of the reference transform's fp32 [3,224,224] tensor of every file
  emb_default      [N,512] embeddings, seeded default-BN weights
  emb_randbn       [N,512] embeddings, seeded randomised-BN weights
"""
from __future__ import annotations

import hashlib
import shutil
import sys
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import src.feature_extraction as fe  # noqa: E402  (the real reference)
from make_golden import _Shim  # noqa: E402

HERE = Path(__file__).resolve().parent
SOURCE = Path("/root/reference/mri_dataset_brain_cancer_oc")
TARGET = HERE / "mri_real"
N_LABELED, N_UNLABELED = 4, 8  # 4 cancer + 4 normal + 8 unlabeled = 16 files


def main():
    torch.set_num_threads(8)
    if TARGET.exists():
        shutil.rmtree(TARGET)
    for label_dir in sorted(d for d in (SOURCE / "avec_labels").iterdir() if d.is_dir()):
        files = sorted(p for p in label_dir.rglob("*") if p.is_file())[:N_LABELED]
        for f in files:
            dst = TARGET / f.relative_to(SOURCE)
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(f, dst)
    for f in sorted(p for p in (SOURCE / "sans_label").rglob("*") if p.is_file())[:N_UNLABELED]:
        dst = TARGET / f.relative_to(SOURCE)
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, dst)

    records = fe.discover_image_records(TARGET)
    transform = fe.build_transform()
    pix, pre = [], []
    for r in records:
        with Image.open(r.absolute_path) as img:
            pix.append(hashlib.sha256(np.ascontiguousarray(np.asarray(img)).tobytes()).hexdigest())
        pre.append(hashlib.sha256(np.ascontiguousarray(fe.preprocess_image(r.absolute_path, transform).numpy()).tobytes()).hexdigest())
    out = {}
    real_models = fe.models
    for name, randbn in (("emb_default", False), ("emb_randbn", True)):
        fe.models = _Shim(randbn)
        try:
            res = fe.extract_embeddings(records, torch.device("cpu"), batch_size=5)
        finally:
            fe.models = real_models
        assert not res.failures and res.embeddings.shape == (len(records), 512)
        out[name] = res.embeddings
    np.savez_compressed(HERE / "mri_real_golden.npz", paths=np.array([str(r.relative_path) for r in records]),
                        pixels_sha256=np.array(pix), pre_sha256=np.array(pre), **out)
    print(f"{len(records)} real MRI files -> {TARGET}, goldens -> mri_real_golden.npz")


if __name__ == "__main__":
    main()
