"""GPU parity of the fused preprocess kernel (fx_preprocess_nchw_f32, through the C ABI) against the
oracle: bit-exact, as the work is integer + table lookup.  Reference behaviour being matched:
src/feature_extraction.py:200-207,233-240 (torchvision Compose on a PIL image)."""
import hashlib
import json

import numpy as np
import pytest
import torch

from oracle import reference_path as rp
from ssip_b200 import _native as N
from ssip_b200 import synthetic
from ssip_b200.engine import Engine, pack_images, uniform_descs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine(0, max_batch=64, precision="fp32")
    yield e
    e.close()


def _gpu_pre(eng, images):
    buf, descs, total = pack_images(images)
    dev = torch.from_numpy(buf[: max(total, 1)]).cuda()
    return eng.preprocess_nchw(dev, descs, len(images)).cpu()


def test_golden_vectors_from_the_real_reference(eng, golden_dir):
    golden = json.loads((golden_dir / "preprocess_golden.json").read_text())
    for key, g in golden.items():
        h, w = map(int, key.split("x"))
        arr = synthetic.ragged_images([(h, w)], seed=g["seed"])[0]
        out = _gpu_pre(eng, [arr])[0].numpy()
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == g["sha256"], key


SHAPES = [(224, 224), (512, 512), (300, 500), (500, 300), (514, 512), (777, 333), (256, 256), (100, 130), (1000, 700),
          (225, 224), (224, 1000), (61, 67), (2048, 1536), (640, 480)]


def test_ragged_batch_bit_exact_vs_oracle(eng):
    images = synthetic.ragged_images(SHAPES, seed=21)
    got = _gpu_pre(eng, images)
    for i, img in enumerate(images):
        want = torch.from_numpy(rp.c_preprocess_rgb(img))
        assert torch.equal(got[i], want), f"{SHAPES[i]}: {(got[i] != want).sum().item()} values differ"


def test_matches_installed_torchvision_transform(eng):
    from PIL import Image

    images = synthetic.ragged_images([(512, 512), (224, 224), (333, 777)], seed=5)
    got = _gpu_pre(eng, images)
    t = rp.port_transform()
    for i, img in enumerate(images):
        assert torch.equal(got[i], t(Image.fromarray(img)))


def test_mri_like_and_gray_carriage(eng):
    imgs = list(synthetic.mri_like_images(4, 512, seed=2))
    rgb = _gpu_pre(eng, imgs)
    gray = _gpu_pre(eng, [np.ascontiguousarray(a[..., 0]) for a in imgs])
    for i, a in enumerate(imgs):
        want = torch.from_numpy(rp.c_preprocess_rgb(a))
        assert torch.equal(rgb[i], want)
        assert torch.equal(gray[i], want)  # one plane replicated == three identical stored planes


def test_extreme_values_and_constant_images(eng):
    imgs = [np.zeros((512, 512, 3), np.uint8), np.full((224, 224, 3), 255, np.uint8), np.full((300, 400, 3), 128, np.uint8)]
    chk = np.indices((512, 512)).sum(0) % 2 * 255
    imgs.append(np.repeat(chk[:, :, None], 3, 2).astype(np.uint8))
    got = _gpu_pre(eng, imgs)
    for i, a in enumerate(imgs):
        assert torch.equal(got[i], torch.from_numpy(rp.c_preprocess_rgb(a)))


def test_full_batch_of_uniform_images(eng):
    x = synthetic.noise_images(64, 224, 224, seed=9)
    dev = torch.from_numpy(x.reshape(-1)).cuda()
    got = eng.preprocess_nchw(dev, uniform_descs(64, 224, 224), 64).cpu()
    for i in (0, 17, 63):
        assert torch.equal(got[i], torch.from_numpy(rp.c_preprocess_rgb(x[i])))
    # size-independent property: every image of the batch gives the same result as when sent alone
    alone = eng.preprocess_nchw(dev[41 * 150528 : 42 * 150528].clone(), uniform_descs(1, 224, 224), 1).cpu()
    assert torch.equal(alone[0], got[41])


def test_errors_are_loud(eng):
    rgba = np.zeros((64, 64, 4), np.uint8)
    buf, descs, total = pack_images([rgba])
    with pytest.raises(N.FxError) as err:
        eng.preprocess_nchw(torch.from_numpy(buf).cuda(), descs, 1)
    assert err.value.status == N.FX_ERR_UNSUPPORTED
    assert eng.preprocess_nchw(torch.zeros(16, dtype=torch.uint8).cuda(), uniform_descs(0, 1, 1), 0).shape[0] == 0


# ---- bf16 conv1 staging tensor (what the trunk actually reads) ------------------------------------

@pytest.fixture(scope="module")
def eng_b():
    e = Engine(0, max_batch=64, precision="bf16")
    yield e
    e.close()


def _check_staging(eng_b, images):
    buf, descs, total = pack_images(images)
    dev = torch.from_numpy(buf[: max(total, 1)]).cuda()
    eng_b.preprocess(dev, descs, len(images))
    raw, padded = eng_b.staged_crop(len(images))
    raw, padded = raw.cpu(), padded.cpu()
    for i, img in enumerate(images):
        want = torch.from_numpy(rp.c_preprocess_rgb(img if img.ndim == 3 else np.repeat(img[:, :, None], 3, 2))).to(torch.bfloat16)
        assert torch.equal(padded[i, :, 3:227, 3:227], want), f"image {i} {img.shape}"
    # conv zero padding of the normalised tensor + unused channels / column: exactly zero
    z = padded.clone()
    z[:, :, 3:227, 3:227] = 0
    assert not z.any()
    assert not raw[..., 12:].any()


@pytest.mark.parametrize("shape", [(224, 224), (512, 512), (256, 256), (300, 500), (500, 300), (448, 448), (225, 224), (640, 480), (100, 130)])
def test_bf16_staging_column_walk_kernel_bit_exact(eng_b, shape):
    """All-RGB, few-tap batches take preprocess_s2d_kernel: staging == bf16(oracle), padding == 0."""
    _check_staging(eng_b, synthetic.ragged_images([shape] * 3, seed=shape[0] + shape[1]))


@pytest.mark.parametrize("shape", [(512, 512), (224, 224), (300, 500), (448, 333), (256, 256), (640, 480)])
def test_bf16_staging_gray_carriage_column_walk_kernel_bit_exact(eng_b, shape):
    """All-gray-carriage, few-tap batches take the one-plane variant of preprocess_s2d_kernel (the C4 workload as one
    plane): the single resampled byte feeds the three channels, each with its own mean / std -- bit-exact against the
    reference transform of the R==G==B image (src/feature_extraction.py:200-207 on a 3-channel file with equal planes)."""
    h, w = shape
    imgs = [np.ascontiguousarray(synthetic.mri_like_images(1, max(h, w), seed=h + w + k)[0][:h, :w, 0]) for k in range(2)]
    imgs.append(np.random.default_rng(h * w).integers(0, 256, (h, w), dtype=np.uint8))  # noise: every tap matters
    _check_staging(eng_b, imgs)


def test_bf16_staging_mixed_batch_and_gray(eng_b):
    """Mixed geometry classes (incl. >5 taps and gray carriage) fall back to the banded kernel; same contract."""
    imgs = synthetic.ragged_images([(224, 224), (1000, 700), (512, 512), (2048, 1536), (61, 67)], seed=3)
    imgs.append(np.ascontiguousarray(synthetic.mri_like_images(1, 512, seed=4)[0][..., 0]))
    _check_staging(eng_b, imgs)
    # a fast batch right after a mixed one reuses the same staging buffer: padding must still be zero
    _check_staging(eng_b, list(synthetic.noise_images(64, 224, 224, seed=1)))


def test_bf16_staging_equals_the_fp32_transform_rounded(eng_b, eng):
    x = list(synthetic.mri_like_images(5, 512, seed=9))
    buf, descs, total = pack_images(x)
    dev = torch.from_numpy(buf[:total]).cuda()
    f32 = eng.preprocess_nchw(dev, descs, len(x))
    eng_b.preprocess(dev, descs, len(x))
    _, padded = eng_b.staged_crop(len(x))
    assert torch.equal(padded[:, :, 3:227, 3:227], f32.to(torch.bfloat16))


def test_repeated_descriptor_table_reuses_the_plan_with_new_pixels(eng_b):
    """Same geometry, different images, alternating lanes/streams: the cached descriptor upload must not cache pixels."""
    s1 = torch.cuda.Stream()
    for rep in range(4):
        imgs = synthetic.ragged_images([(224, 224), (512, 512), (300, 500)], seed=100 + rep)
        eng_b.select_lane(rep & 1)
        with torch.cuda.stream(s1 if rep & 1 else torch.cuda.current_stream()):
            _check_staging(eng_b, imgs)
    eng_b.select_lane(0)
