"""Mint tests/golden/head_golden.npz from the REAL reference's classifier-head inference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_head.py

Imports, unmodified, ``src/training/common.py`` (create_model :299-304, build_transforms :96-119, evaluate_model
:439-506), ``src/training/semi_supervised.py`` (generate_pseudo_labels :44-72) and ``src/threshold_sweep.py``
(compute_probs :21-38).  Two shims only: ``matplotlib`` is not installed here (the plotting helpers are never
called), and the ImageNet download is avoided with ``pretrained=False`` + the oracle's seeded parameters.
Inputs are regenerated from seeds by ``ssip_b200.synthetic``; only outputs are committed.
"""
from __future__ import annotations

import hashlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/src")  # threshold_sweep.py imports `training.common`

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))

import src.training.common as ref_common  # noqa: E402
import src.training.semi_supervised as ref_semi  # noqa: E402
import threshold_sweep as ref_sweep  # noqa: E402

from oracle.reference_path import make_classifier  # noqa: E402
from ssip_b200 import synthetic  # noqa: E402

HERE = Path(__file__).resolve().parent


def head_inputs():
    """The fixture's images: MRI-like 512x512 plus ragged RGB noise (seeded)."""
    imgs = list(synthetic.mri_like_images(6, 512, seed=21))
    imgs += synthetic.ragged_images([(300, 500), (224, 224), (640, 480), (100, 130)], seed=22)
    labels = [0, 1, 1, 0, 1, 0, 0, 1, 1, 0]
    paths = [f"img_{i:02d}.png" for i in range(len(imgs))]
    return imgs, labels, paths


def main():
    from PIL import Image

    torch.set_num_threads(8)
    imgs, labels, paths = head_inputs()
    transform = ref_common.build_transforms()["eval"]
    tensors = [transform(Image.fromarray(a).convert("RGB")) for a in imgs]  # as the reference datasets do (:171,:191)
    sha = [hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest() for t in tensors]

    model = ref_common.create_model(2, pretrained=False)
    seeded = make_classifier(2)
    assert list(model.state_dict().keys()) == list(seeded.state_dict().keys())
    model.load_state_dict(seeded.state_dict())
    model.eval()

    x = torch.stack(tensors)
    bs = 4
    loader = [(x[i : i + bs], torch.tensor(labels[i : i + bs]), paths[i : i + bs]) for i in range(0, len(imgs), bs)]
    unl_loader = [(x[i : i + bs], paths[i : i + bs]) for i in range(0, len(imgs), bs)]
    with torch.no_grad():
        logits = model(x)
        probs = torch.softmax(logits, dim=1)
    dev = torch.device("cpu")
    m0, yt0, yp0, ypr0, sp0 = ref_common.evaluate_model(model, loader, dev)
    m1, yt1, yp1, ypr1, _ = ref_common.evaluate_model(model, loader, dev, pos_index=0, threshold=0.5)
    thr = float(np.median(probs.max(dim=1).values.numpy()))  # keeps about half of the samples
    pseudo = ref_semi.generate_pseudo_labels(model, unl_loader, dev, threshold=thr)
    ct, cp = ref_sweep.compute_probs(model, [(a, b) for a, b, _ in loader], dev, pos_index=1)
    np.savez_compressed(
        HERE / "head_golden.npz",
        transform_sha256=np.array(sha),
        logits=logits.numpy(), probs=probs.numpy(),
        eval_default_metrics=np.array([m0["accuracy"], m0["precision"], m0["recall"], m0["f1"]]),
        eval_default_y_true=yt0, eval_default_y_pred=yp0, eval_default_y_prob=ypr0, eval_default_paths=np.array(sp0),
        eval_thr_metrics=np.array([m1["accuracy"], m1["precision"], m1["recall"], m1["f1"]]),
        eval_thr_y_pred=yp1, eval_thr_y_prob=ypr1,
        pseudo_threshold=np.array(thr),
        pseudo_paths=np.array([p for p, _, _ in pseudo]), pseudo_labels=np.array([l for _, l, _ in pseudo]),
        pseudo_conf=np.array([c for _, _, c in pseudo]),
        sweep_y_true=ct, sweep_y_prob=cp,
    )
    print("logits", logits.numpy().round(4).tolist())
    print("probs", probs.numpy().round(4).tolist())
    print("pseudo", len(pseudo), "of", len(imgs), "at threshold", thr)
    print("metrics", m0, m1)


if __name__ == "__main__":
    main()
