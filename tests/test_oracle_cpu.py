"""The oracle (oracle/) pinned against (a) the golden vectors minted from the REAL reference module
(tests/golden/make_golden.py), (b) the installed torchvision/Pillow transform the reference calls,
(c) the reference module itself when /root/reference is mounted (build container only)."""
import hashlib
import json
import sys

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import reference_path as rp
from ssip_b200 import synthetic

SHAPES = [(224, 224), (512, 512), (300, 500), (500, 300), (514, 512), (777, 333), (256, 256), (100, 130), (1000, 700)]


def test_c_oracle_matches_reference_golden_preprocess(golden_dir):
    golden = json.loads((golden_dir / "preprocess_golden.json").read_text())
    assert set(golden) == {f"{h}x{w}" for h, w in SHAPES}
    for key, g in golden.items():
        h, w = map(int, key.split("x"))
        arr = synthetic.ragged_images([(h, w)], seed=g["seed"])[0]
        out = rp.c_preprocess_rgb(arr)
        assert hashlib.sha256(out.tobytes()).hexdigest() == g["sha256"], key
        np.testing.assert_array_equal(out[:, ::37, ::41].astype(np.float64).round(9), np.array(g["sample"]))


@pytest.mark.parametrize("shape", SHAPES + [(225, 224), (224, 1000), (2048, 1536), (61, 67)])
def test_c_oracle_matches_installed_torchvision(shape):
    arr = synthetic.ragged_images([shape], seed=sum(shape))[0]
    want = rp.port_transform()(Image.fromarray(arr))
    got = torch.from_numpy(rp.c_preprocess_rgb(arr))
    assert torch.equal(want, got)


def test_c_oracle_matches_installed_torchvision_on_seeded_random_geometries():
    """48 seeded (H, W) pairs: extreme aspect ratios, sides of 1..3 pixels, up- and down-scales up to 13x, odd crops.
    The reference path is Image -> Resize(256) -> CenterCrop(224) -> ToTensor -> Normalize (src/feature_extraction.py:200-205)."""
    rng = np.random.default_rng(20261018)
    shapes = [(1, 1), (1, 700), (700, 1), (2, 3), (3, 2), (223, 225), (255, 257), (257, 255), (256, 3000), (3000, 256)]
    while len(shapes) < 48:
        h, w = int(rng.integers(4, 1400)), int(rng.integers(4, 1400))
        if h * w <= 1_200_000:
            shapes.append((h, w))
    transform = rp.port_transform()
    for h, w in shapes:
        arr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = transform(Image.fromarray(arr))
        got = torch.from_numpy(rp.c_preprocess_rgb(arr))
        assert torch.equal(want, got), (h, w)


def test_c_oracle_resize_is_pillow_exact_on_gray_and_rgb():
    rng = np.random.default_rng(3)
    for (h, w, oh, ow) in [(512, 512, 256, 256), (224, 224, 256, 256), (300, 500, 256, 426), (100, 100, 256, 256), (512, 512, 224, 224)]:
        for c in (1, 3):
            a = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
            img = Image.fromarray(a[:, :, 0] if c == 1 else a)
            want = np.asarray(img.resize((ow, oh), Image.BILINEAR)).reshape(oh, ow, c)
            got = np.empty((oh, ow, c), np.uint8)
            rp.c_oracle().fxo_resize_bilinear_u8(a.ctypes.data, h, w, c, got.ctypes.data, oh, ow)
            np.testing.assert_array_equal(got, want)


def test_known_answer_coefficients():
    # SURVEY.md Appendix A worked constants
    bounds, kk = rp.c_coeffs(512, 256)
    assert kk.shape[1] == 5
    one = 1 << 22
    np.testing.assert_array_equal(kk[100, :4], [one // 8, 3 * one // 8, 3 * one // 8, one // 8])
    assert tuple(bounds[100]) == (199, 4)
    assert tuple(bounds[0]) == (0, 3) and tuple(bounds[255]) == (509, 3)
    np.testing.assert_array_equal(kk[0, :3], np.round(np.array([3 / 7, 3 / 7, 1 / 7]) * one).astype(np.int32))
    bounds, kk = rp.c_coeffs(224, 256)
    assert kk.shape[1] == 3
    assert tuple(bounds[0]) == (0, 1) and kk[0, 0] == one
    np.testing.assert_array_equal(kk[1, :2], [int(0.1875 * one), int(0.8125 * one)])
    assert tuple(bounds[128]) == (111, 2)


def test_lut_matches_torch_ops():
    lut = rp.c_lut()
    v = torch.arange(256, dtype=torch.uint8)
    x = v.to(torch.float32).div(255)
    for c in range(3):
        want = (x - torch.tensor(rp.MEAN[c], dtype=torch.float32)) / torch.tensor(rp.STD[c], dtype=torch.float32)
        assert torch.equal(torch.from_numpy(lut[c]), want)


def test_crop_offsets_round_half_even():
    lib = rp.c_oracle()
    for size in range(224, 700):
        assert lib.fxo_crop_offset(size, 224) == int(round((size - 224) / 2.0))


def test_port_embeddings_match_reference_golden(golden_dir):
    z = np.load(golden_dir / "embeddings_golden.npz")
    noise = list(synthetic.noise_images(16, 224, 224, seed=0))
    np.testing.assert_array_equal(rp.port_embed_arrays(noise, randomize_bn=False), z["noise_default"])
    np.testing.assert_array_equal(rp.port_embed_arrays(noise, randomize_bn=True), z["noise_randbn"])
    mri = list(synthetic.mri_like_images(8, 512, seed=7))
    np.testing.assert_array_equal(rp.port_embed_arrays(mri, randomize_bn=True), z["mri_randbn"])


def test_golden_embeddings_are_discriminating(golden_dir):
    # cosine between DIFFERENT images is already ~0.999 with random weights, so the
    # parity tests lean on relative L2; make sure the golden rows are not degenerate duplicates.
    z = np.load(golden_dir / "embeddings_golden.npz")
    e = z["mri_randbn"]
    rel = np.linalg.norm(e[0] - e[1]) / np.linalg.norm(e[0])
    assert rel > 2e-2


def test_port_matches_real_reference_module(have_reference, tmp_path):
    if not have_reference:
        pytest.skip("/root/reference not mounted (GPU box)")
    sys.path.insert(0, "/root/reference")
    import src.feature_extraction as fe

    imgs = synthetic.ragged_images([(224, 224), (300, 500), (512, 512)], seed=5)
    synthetic.write_png_dataset(tmp_path, imgs, n_labeled=2)
    records = fe.discover_image_records(tmp_path)
    transform = fe.build_transform()
    order = synthetic.dataset_order(len(imgs), n_labeled=2)
    for rec, idx in zip(records, order):
        want = fe.preprocess_image(rec.absolute_path, transform)
        assert torch.equal(want, torch.from_numpy(rp.c_preprocess_rgb(imgs[idx])))
    port_records = [rp.PortRecord(r.absolute_path, r.relative_path, r.bucket, r.label) for r in records]
    real_models = fe.models

    class Shim:
        ResNet18_Weights = real_models.ResNet18_Weights

        @staticmethod
        def resnet18(weights=None):
            return rp.make_backbone(randomize_bn=True)

    fe.models = Shim
    try:
        ref = fe.extract_embeddings(records, torch.device("cpu"), batch_size=2)
    finally:
        fe.models = real_models
    port = rp.port_extract_embeddings(port_records, torch.device("cpu"), batch_size=2, randomize_bn=True)
    np.testing.assert_array_equal(ref.embeddings, port.embeddings)
    assert [r.relative_path for r in ref.records] == [r.relative_path for r in port.records]


# ---- classifier-head inference -------------------------------------------


def _head_inputs():
    imgs = list(synthetic.mri_like_images(6, 512, seed=21))
    imgs += synthetic.ragged_images([(300, 500), (224, 224), (640, 480), (100, 130)], seed=22)
    return imgs, [0, 1, 1, 0, 1, 0, 0, 1, 1, 0], [f"img_{i:02d}.png" for i in range(10)]


def test_eval_transform_restatement_matches_reference_golden(golden_dir):
    import hashlib

    g = np.load(golden_dir / "head_golden.npz")
    imgs, _, _ = _head_inputs()
    for i, a in enumerate(imgs):
        out = rp.c_preprocess_square224(a)
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g["transform_sha256"][i])


def test_head_port_matches_reference_golden(golden_dir):
    """The oracle's restatement of generate_pseudo_labels / compute_probs (and the seeded classifier) reproduce what
    the real reference functions returned in the build container (tests/golden/make_golden_head.py)."""
    g = np.load(golden_dir / "head_golden.npz")
    imgs, labels, paths = _head_inputs()
    torch.set_num_threads(8)
    model = rp.make_classifier(2)
    x = torch.stack([torch.from_numpy(rp.c_preprocess_square224(a)) for a in imgs])
    with torch.no_grad():
        logits = model(x)
    assert np.allclose(logits.numpy(), g["logits"], rtol=0, atol=2e-4)
    loader = [(x[i : i + 4], torch.tensor(labels[i : i + 4]), paths[i : i + 4]) for i in range(0, 10, 4)]
    ct, cp = rp.port_compute_probs(model, [(a, b) for a, b, _ in loader], torch.device("cpu"), pos_index=1)
    assert np.array_equal(ct, g["sweep_y_true"]) and np.allclose(cp, g["sweep_y_prob"], atol=1e-5)
    got = rp.port_generate_pseudo_labels(model, [(a, p) for a, _, p in loader], torch.device("cpu"), float(g["pseudo_threshold"]))
    assert [p for p, _, _ in got] == g["pseudo_paths"].tolist()
    assert [l for _, l, _ in got] == g["pseudo_labels"].tolist()


# ---- real MRI files: fixture minted by tests/golden/make_golden_mri.py from the real reference ----


def mri_real_records(golden_dir):
    """The 16 committed JPEGs of the reference's own dataset in discover_image_records order (= the fixture's row order)."""
    from ssip_b200.feature_extraction import discover_image_records

    g = np.load(golden_dir / "mri_real_golden.npz")
    records = discover_image_records(golden_dir / "mri_real")
    assert [str(r.relative_path) for r in records] == g["paths"].tolist()
    return g, records


def test_real_mri_fixture_pins_decode_transform_and_port(golden_dir):
    g, records = mri_real_records(golden_dir)
    assert len(records) == 16 and {r.label for r in records} == {"cancer", "normal", None}
    for r, pix, pre in zip(records, g["pixels_sha256"], g["pre_sha256"]):
        with Image.open(r.absolute_path) as img:
            arr = np.ascontiguousarray(np.asarray(img))
        assert arr.shape == (512, 512, 3)
        assert hashlib.sha256(arr.tobytes()).hexdigest() == str(pix)  # same libjpeg decode as in the build container
        assert hashlib.sha256(rp.c_preprocess_rgb(arr).tobytes()).hexdigest() == str(pre)  # C restatement == reference transform
    port = [rp.PortRecord(r.absolute_path, r.relative_path, r.bucket, r.label) for r in records]
    for key, randbn in (("emb_default", False), ("emb_randbn", True)):
        got = rp.port_extract_embeddings(port, torch.device("cpu"), batch_size=5, randomize_bn=randbn).embeddings
        rel = np.linalg.norm(got - g[key], axis=1) / np.linalg.norm(g[key], axis=1)
        assert rel.max() <= 1e-6, (key, rel.max())  # same torch build; oneDNN may pick another kernel on another CPU
