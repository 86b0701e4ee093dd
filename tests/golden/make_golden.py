"""Mint the golden vectors in tests/golden/ from the REAL reference module.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``/root/reference/src/feature_extraction.py`` unmodified and patches exactly one thing,
the pretrained-weight download inside ``load_model`` (src/feature_extraction.py:217-218), by
swapping ``fe.models.resnet18`` for a seeded ``weights=None`` constructor.
Inputs are regenerated from seeds by ``ssip_b200.synthetic`` so only the outputs are committed.
"""
from __future__ import annotations

import hashlib
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import src.feature_extraction as fe  # noqa: E402  (the real reference)
from oracle.reference_path import make_backbone  # noqa: E402
from ssip_b200 import synthetic  # noqa: E402

PREPROCESS_SHAPES = [(224, 224), (512, 512), (300, 500), (500, 300), (514, 512), (777, 333), (256, 256), (100, 130), (1000, 700)]
HERE = Path(__file__).resolve().parent


class _Shim:
    """Stands in for ``torchvision.models`` inside the reference module: same attribute names,
    but resnet18() ignores ``weights`` (no network) and is seeded."""

    def __init__(self, randomize_bn):
        self.randomize_bn = randomize_bn
        self.ResNet18_Weights = fe.models.ResNet18_Weights

    def resnet18(self, weights=None):
        return make_backbone(randomize_bn=self.randomize_bn)


def reference_embeddings(images, randomize_bn, batch_size=32):
    real_models = fe.models
    fe.models = _Shim(randomize_bn)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            synthetic.write_png_dataset(Path(tmp), images, n_labeled=min(4, len(images)))
            records = fe.discover_image_records(Path(tmp))
            res = fe.extract_embeddings(records, torch.device("cpu"), batch_size=batch_size)
        order = synthetic.dataset_order(len(images), n_labeled=min(4, len(images)))
        emb = np.empty_like(res.embeddings)
        emb[order] = res.embeddings  # back to generator index order
        return emb
    finally:
        fe.models = real_models


def main():
    from PIL import Image

    torch.set_num_threads(8)
    transform = fe.build_transform()
    pre = {}
    for i, (h, w) in enumerate(PREPROCESS_SHAPES):
        arr = synthetic.ragged_images([(h, w)], seed=100 + i)[0]
        t = transform(Image.fromarray(arr)).numpy()
        pre[f"{h}x{w}"] = {
            "seed": 100 + i,
            "sha256": hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest(),
            "sample": t[:, ::37, ::41].astype(np.float64).round(9).tolist(),
        }
    (HERE / "preprocess_golden.json").write_text(json.dumps(pre, indent=1))

    noise = list(synthetic.noise_images(16, 224, 224, seed=0))
    mri = list(synthetic.mri_like_images(8, 512, seed=7))
    ragged = synthetic.ragged_images([(300, 500), (500, 300), (514, 512), (777, 333), (256, 256), (640, 480)], seed=11)
    np.savez_compressed(
        HERE / "embeddings_golden.npz",
        noise_default=reference_embeddings(noise, False),
        noise_randbn=reference_embeddings(noise, True),
        mri_randbn=reference_embeddings(mri, True),
        ragged_randbn=reference_embeddings(ragged, True),
    )
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
