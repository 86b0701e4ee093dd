"""GPU parity of the classifier-head inference drop-ins against golden vectors minted from the
REAL reference (tests/golden/make_golden_head.py: src/training/common.py evaluate_model, src/training/semi_supervised.py
generate_pseudo_labels, src/threshold_sweep.py compute_probs) and against the oracle port.

Tolerances.  The evaluation transform is integer work: bit-exact.  Logits are w.emb + b with the embedding within the
north-star's relative L2 (1e-2 in bf16, 1e-5 in the fp32 mode), so |dlogit_c| <= tol * ||w_c|| * ||emb|| (Cauchy-Schwarz)
-- asserted in that form; decisions (argmax, thresholds) must agree wherever the reference's margin exceeds that bound."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import reference_path as rp
from ssip_b200 import _native as N
from ssip_b200 import inference, synthetic
from ssip_b200.engine import Engine, pack_images

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def head_inputs():
    imgs = list(synthetic.mri_like_images(6, 512, seed=21))
    imgs += synthetic.ragged_images([(300, 500), (224, 224), (640, 480), (100, 130)], seed=22)
    labels = [0, 1, 1, 0, 1, 0, 0, 1, 1, 0]
    paths = [f"img_{i:02d}.png" for i in range(len(imgs))]
    return imgs, labels, paths


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(golden_dir / "head_golden.npz")


@pytest.fixture(scope="module")
def model():
    return rp.make_classifier(2)


@pytest.fixture(scope="module")
def batch_tensor():
    imgs, _, _ = head_inputs()
    return torch.stack([torch.from_numpy(rp.c_preprocess_square224(a)) for a in imgs])


def _bound(model, x, tol):
    trunk = torch.nn.Sequential(*list(model.children())[:-1])
    with torch.no_grad():
        emb = torch.flatten(trunk(x), 1)
    return tol * emb.norm(dim=1)[:, None] * model.fc.weight.detach().norm(dim=1)[None, :]  # [n, C]


def test_eval_transform_bit_exact_vs_reference_golden(golden):
    imgs, _, _ = head_inputs()
    eng = Engine(0, max_batch=16, precision="fp32")
    eng.set_transform(N.TRANSFORM_SQUARE224)
    buf, descs, total = pack_images(imgs)
    out = eng.preprocess_nchw(torch.from_numpy(buf[:total]).cuda(), descs, len(imgs)).cpu().numpy()
    for i in range(len(imgs)):
        assert hashlib.sha256(np.ascontiguousarray(out[i]).tobytes()).hexdigest() == str(golden["transform_sha256"][i]), i
    # more geometries against the C restatement, incl. one-axis identity, > 5 taps, up- and down-scaling mixed
    shapes = [(224, 500), (777, 224), (1000, 700), (61, 67), (448, 448), (2048, 1536), (225, 223)]
    more = synthetic.ragged_images(shapes, seed=9)
    buf, descs, total = pack_images(more)
    out = eng.preprocess_nchw(torch.from_numpy(buf[:total]).cuda(), descs, len(more)).cpu().numpy()
    for i, a in enumerate(more):
        assert np.array_equal(out[i], rp.c_preprocess_square224(a)), shapes[i]
    # the transform object handed to datasets
    from PIL import Image

    t = inference.build_transforms()["eval"]
    assert torch.equal(t(Image.fromarray(imgs[6])), torch.from_numpy(rp.c_preprocess_square224(imgs[6])))
    eng.set_transform(N.TRANSFORM_EXTRACT)  # and back: the default geometry is unaffected
    assert np.array_equal(eng.preprocess_nchw(torch.from_numpy(buf[:total]).cuda(), descs, 1).cpu().numpy()[0], rp.c_preprocess_rgb(more[0]))
    eng.close()


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("fp32", 1e-5)])
def test_logits_and_probs_match_reference_golden(golden, model, batch_tensor, precision, tol):
    clf = inference.get_classifier(model, DEV, min_batch=8, precision=precision)  # 10 images -> two engine batches
    logits, probs = clf.logits_probs(batch_tensor)
    logits, probs = logits.cpu().numpy(), probs.cpu().numpy()
    bound = _bound(model, batch_tensor, tol).numpy()
    err = np.abs(logits - golden["logits"])
    assert (err <= bound + 1e-6).all(), f"max logit error {err.max():.4f}, bound {bound.min():.4f}..{bound.max():.4f}"
    # softmax: compare against the exact softmax of OUR logits (kernel arithmetic), then against the golden
    mine = torch.softmax(torch.from_numpy(logits), dim=1).numpy()
    assert np.abs(probs - mine).max() <= 2e-6
    gap_err = np.abs((logits[:, 1] - logits[:, 0]) - (golden["logits"][:, 1] - golden["logits"][:, 0]))
    assert (np.abs(probs - golden["probs"]).max(axis=1) <= gap_err / 4 + 2e-6).all()  # d sigmoid / dx <= 1/4
    assert np.allclose(probs.sum(1), 1.0, atol=1e-6)


def test_fused_raw_image_path_equals_tensor_path(model, batch_tensor):
    """uint8 images -> FX_TRANSFORM_SQUARE224 -> trunk -> head == the same engine fed the reference's tensors."""
    imgs, _, _ = head_inputs()
    clf = inference.get_classifier(model, DEV, min_batch=16)
    emb, logits, probs = clf.classify_arrays(imgs)
    lg2, pr2 = clf.logits_probs(batch_tensor)
    assert np.array_equal(logits, lg2.cpu().numpy()) and np.array_equal(probs, pr2.cpu().numpy())
    gray = clf.classify_arrays([np.ascontiguousarray(a[..., 0]) for a in imgs[:6]])  # convert("RGB") of a gray file
    assert np.array_equal(gray[1], logits[:6])
    assert emb.shape == (10, 512) and np.isfinite(emb).all()


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_three_reference_loops_match_golden(golden, model, batch_tensor, precision, monkeypatch):
    monkeypatch.setenv("SSIP_B200_PRECISION", precision)
    _, labels, paths = head_inputs()
    x, bs = batch_tensor, 4
    loader = [(x[i : i + bs], torch.tensor(labels[i : i + bs]), paths[i : i + bs]) for i in range(0, len(labels), bs)]
    ptol = 0.1 if precision == "bf16" else 1e-4
    # evaluate_model, default arguments
    m, yt, yp, ypr, sp = inference.evaluate_model(model, loader, DEV)
    assert np.array_equal(yt, golden["eval_default_y_true"]) and list(sp) == list(golden["eval_default_paths"])
    assert np.array_equal(yp, golden["eval_default_y_pred"])  # the fixture's smallest margin is ~1.07 logits
    assert np.abs(ypr - golden["eval_default_y_prob"]).max() <= ptol
    assert np.allclose([m["accuracy"], m["precision"], m["recall"], m["f1"]], golden["eval_default_metrics"])
    # evaluate_model with pos_index / threshold
    m, _, yp, ypr, _ = inference.evaluate_model(model, loader, DEV, pos_index=0, threshold=0.5)
    assert np.array_equal(yp, golden["eval_thr_y_pred"])
    assert np.abs(ypr - golden["eval_thr_y_prob"]).max() <= ptol
    assert np.allclose([m["accuracy"], m["precision"], m["recall"], m["f1"]], golden["eval_thr_metrics"])
    # compute_probs
    ct, cp = inference.compute_probs(model, [(a, b) for a, b, _ in loader], DEV, pos_index=1)
    assert np.array_equal(ct, golden["sweep_y_true"]) and np.abs(cp - golden["sweep_y_prob"]).max() <= ptol
    # generate_pseudo_labels: same selection wherever the reference's confidence is not within ptol of the threshold
    thr = float(golden["pseudo_threshold"])
    got = inference.generate_pseudo_labels(model, [(a, p) for a, _, p in loader], DEV, threshold=thr)
    ref_conf = golden["probs"].max(axis=1)
    sure_in = {paths[i] for i in range(len(paths)) if ref_conf[i] >= thr + ptol}
    sure_out = {paths[i] for i in range(len(paths)) if ref_conf[i] < thr - ptol}
    got_paths = {p for p, _, _ in got}
    assert sure_in <= got_paths and not (got_paths & sure_out)
    ref_label = dict(zip(golden["pseudo_paths"].tolist(), golden["pseudo_labels"].tolist()))
    for p, lab, conf in got:
        if p in ref_label:
            assert lab == ref_label[p]
    if precision == "fp32":
        assert sorted(got_paths) == sorted(golden["pseudo_paths"].tolist())


def test_errors_are_loud(model):
    eng = Engine(0, max_batch=4)
    with pytest.raises(N.FxError) as err:
        eng._check(eng._lib.fx_classify(eng._h, 1, None, None, None, None))
    assert err.value.status == N.FX_ERR_STATE
    with pytest.raises(N.FxError):
        eng.set_transform(7)
    with pytest.raises(ValueError):
        eng.load_head(torch.zeros(2, 100), torch.zeros(2))
    with pytest.raises(N.FxError):
        eng.load_head(torch.zeros(N.MAX_CLASSES + 1, 512), torch.zeros(N.MAX_CLASSES + 1))
    with pytest.raises(RuntimeError):
        inference.get_classifier(model, torch.device("cpu"))
    eng.close()
