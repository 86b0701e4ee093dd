"""Parity at the batch sizes the headline numbers run on (BASELINE.json configs[1], [2] and the top of [4]): batch 256,
512 and 1024 on bf16, batch 256 on the fp32 tight mode -- the kernels choose other tilings there (split tail wave in
layer 4, other M-tile boxes, more work tiles per CTA) than at the batch <= 64 sizes of test_gpu_trunk.py.

Two checks per size: (a) every row equals, bit for bit, the row the same image gets in a batch of 64 (those are the rows
test_gpu_trunk.py compares with the reference goldens), (b) a seeded subsample of rows is compared DIRECTLY with the
oracle port of src/feature_extraction.py:272-300 on the same images (BASELINE.json tolerance: relative L2 <= 1e-2 and
cosine >= 0.999 in bf16; relative L2 <= 1e-5 in the tight mode)."""
import numpy as np
import pytest
import torch

from oracle import reference_path as rp
from ssip_b200 import synthetic
from ssip_b200.engine import Engine, uniform_descs

pytestmark = pytest.mark.gpu


def _run(eng, x_dev, n, batch, h=224, w=224):
    out = torch.empty((n, 512), dtype=torch.float32, device="cuda")
    descs = uniform_descs(batch, h, w)
    per = h * w * 3
    for s in range(0, n, batch):
        eng.embed_device(x_dev[s * per : (s + batch) * per], descs, batch, out=out[s : s + batch])
    torch.cuda.synchronize()
    return out


def _oracle_rows(images, rows, randbn):
    return rp.port_embed_arrays([images[i] for i in rows], randomize_bn=randbn)


def _check(got, want, rel_tol, cos_tol):
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
    assert rel.max() <= rel_tol and cos.min() >= cos_tol, (rel.max(), cos.min())


def test_bf16_batch_256_512_1024_rows_equal_batch_64_and_the_oracle():
    n = 2048
    x = synthetic.noise_images(n, 224, 224, seed=77)
    x[5] = synthetic.mri_like_images(1, 224, seed=2)[0]  # a structured image among the noise
    dev = torch.from_numpy(x.reshape(-1)).cuda()
    eng = Engine(0, max_batch=1024, precision="bf16")
    eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
    base = _run(eng, dev, n, 64)
    for batch in (256, 512, 1024):
        got = _run(eng, dev, n, batch)
        assert torch.equal(got, base), f"batch {batch}: rows differ from batch 64 by up to {(got - base).abs().max().item():.3e}"
    # both lanes at batch 512 (what bench.py alternates between at N > 1)
    eng.select_lane(1)
    assert torch.equal(_run(eng, dev, n, 512), base)
    eng.select_lane(0)
    rows = sorted(np.random.default_rng(9).choice(n, size=63, replace=False).tolist() + [5])
    _check(base[rows].cpu().numpy(), _oracle_rows(x, rows, True), 1e-2, 0.999)
    assert bool(torch.isfinite(base).all())
    eng.close()


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_odd_batch_sizes_give_the_rows_of_batch_8(precision):
    """Ragged last batches (and any rank's shard under torchrun) run at whatever size is left: odd leftover image of a CTA pair,
    partial M-tiles, other split-tail decisions.  Row i must not depend on how many images travel with it."""
    n = 1000 if precision == "bf16" else 200
    x = synthetic.noise_images(n, 224, 224, seed=123)
    dev = torch.from_numpy(x.reshape(-1)).cuda()
    eng = Engine(0, max_batch=1024 if precision == "bf16" else 256, precision=precision)
    eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
    base = _run(eng, dev[: (n // 8) * 8 * 150528], (n // 8) * 8, 8)  # n is a multiple of 8
    sizes = (1, 2, 3, 5, 37, 63, 65, 100, 129, 255, 257, 511, 777, 1000) if precision == "bf16" else (1, 3, 37, 65, 129, 200)
    for b in sizes:
        out = eng.embed_device(dev[: b * 150528], uniform_descs(b, 224, 224), b)
        torch.cuda.synchronize()
        assert torch.equal(out, base[:b]), f"{precision} batch {b}: differs from the batch-8 rows by up to {(out - base[:b]).abs().max().item():.3e}"
    eng.close()


def test_bf16_batch_512_ragged_sources_equal_batch_32():
    # 512x512 sources (the C4 geometry) at batch 512: the preprocess takes the 4-tap path, the trunk the batch-512 tilings
    n = 512
    x = np.stack(list(synthetic.mri_like_images(n, 512, seed=13)))
    dev = torch.from_numpy(x.reshape(-1)).cuda()
    eng = Engine(0, max_batch=512, precision="bf16")
    eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
    base = _run(eng, dev, n, 32, 512, 512)
    assert torch.equal(_run(eng, dev, n, 512, 512, 512), base)
    rows = sorted(np.random.default_rng(10).choice(n, size=16, replace=False).tolist())
    _check(base[rows].cpu().numpy(), _oracle_rows(x, rows, True), 1e-2, 0.999)
    eng.close()


def test_fp32_tight_mode_batch_256_rows_equal_batch_32_and_the_oracle():
    n = 512
    x = synthetic.noise_images(n, 224, 224, seed=78)
    dev = torch.from_numpy(x.reshape(-1)).cuda()
    eng = Engine(0, max_batch=256, precision="fp32")
    eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
    base = _run(eng, dev, n, 32)
    got = _run(eng, dev, n, 256)
    assert torch.equal(got, base)
    rows = sorted(np.random.default_rng(11).choice(n, size=32, replace=False).tolist())
    _check(got[rows].cpu().numpy(), _oracle_rows(x, rows, True), 1e-5, 0.999999)
    eng.close()
