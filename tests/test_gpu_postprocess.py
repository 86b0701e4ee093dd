"""GPU parity of the on-device post-processing: fx_column_stats / fx_standardize /
fx_neighbor_probe against golden vectors minted from the REAL reference (tests/golden/make_golden_post.py:
run_sanity_checks + nearest_neighbor_probe of src/feature_extraction.py:334-398, standardize_embeddings of
src/standardize_features.py:12-61) and against numpy / scikit-learn run here on the same matrix.

Tolerances: the reference's own statistics are fp32 numpy reductions (row-by-row accumulation), ours accumulate in
fp64 -> agreement to 2e-6 relative at this size, and to 1e-12 against an fp64 numpy evaluation.  The scaler's mean /
scale are fp64 in scikit-learn as well: 1e-12, and equal after rounding to fp32; standardized features: bit-exact.  Neighbour rows must be identical
(first maximum on ties), similarities within 2e-6."""
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch

from ssip_b200 import _native as N
from ssip_b200 import feature_extraction as fx
from ssip_b200 import standardize_features as sf
from ssip_b200.engine import Engine

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def post_matrix(n: int = 700, d: int = 512, seed: int = 77) -> np.ndarray:  # == tests/golden/make_golden_post.py
    rng = np.random.default_rng(seed)
    base = rng.gamma(2.0, 0.5, size=(1, d)).astype(np.float32)
    x = (base * (1.0 + 0.15 * rng.standard_normal((n, d)))).astype(np.float32)
    x = np.abs(x) + np.float32(0.01)
    if n > 650:
        x[5] = x[400]
        x[650] = x[17]
    x[:, 3] = np.float32(0.75)
    x[:, 9] = np.float32(1.5) + np.float32(1e-7) * (np.arange(n) % 2)
    return np.ascontiguousarray(x)


def _records(n):
    return [fx.ImageRecord(Path(f"/d/img_{i:04d}.png"), Path(f"img_{i:04d}.png"), "labeled" if i < 40 else "unlabeled",
                           ("cancer" if i % 2 else "normal") if i < 40 else None) for i in range(n)]


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(golden_dir / "post_golden.npz")


@pytest.fixture(scope="module")
def eng():
    e = Engine(0, max_batch=1)
    yield e
    e.close()


def test_sanity_statistics_match_reference_golden_and_fp64(golden, eng):
    x = post_matrix()
    dev = torch.from_numpy(x).cuda()
    stats = fx.run_sanity_checks(dev)
    assert stats["num_vectors"] == 700 and stats["dimension"] == 512
    assert abs(stats["mean_abs_mean"] - float(golden["mean_abs_mean"])) <= 2e-6 * float(golden["mean_abs_mean"])
    assert abs(stats["mean_std"] - float(golden["mean_std"])) <= 2e-6 * float(golden["mean_std"])
    host = fx.run_sanity_checks(x)  # the numpy path of the same function == the reference's expressions
    assert host["mean_abs_mean"] == float(golden["mean_abs_mean"]) and host["mean_std"] == float(golden["mean_std"])
    x64 = x.astype(np.float64)
    st, mean, std, var = eng.column_stats(dev)
    assert np.allclose(mean.cpu().numpy(), x64.mean(0), rtol=1e-13, atol=0)
    assert np.allclose(var.cpu().numpy(), x64.var(0), rtol=1e-10, atol=1e-30)
    assert abs(st["mean_abs_mean"] - np.abs(x64.mean(0)).mean()) <= 1e-12
    assert abs(st["mean_std"] - x64.std(0).mean()) <= 1e-12
    assert st["nan_count"] == 0 and st["inf_count"] == 0


def test_nan_and_inf_raise_like_the_reference(eng):
    x = post_matrix(300)
    bad = x.copy()
    bad[17, 100] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        fx.run_sanity_checks(torch.from_numpy(bad).cuda())
    bad = x.copy()
    bad[299, 511] = np.inf
    bad[0, 0] = -np.inf
    with pytest.raises(ValueError, match="inf"):
        fx.run_sanity_checks(torch.from_numpy(bad).cuda())
    st, _, _, _ = eng.column_stats(torch.from_numpy(bad).cuda())
    assert st["inf_count"] == 2 and st["nan_count"] == 0


def test_neighbor_probe_matches_reference_golden_and_tie_rule(golden, eng):
    x = post_matrix()
    recs = _records(700)
    dev = torch.from_numpy(x).cuda()
    probe = fx.nearest_neighbor_probe(dev, recs)
    assert [p["query"] for p in probe] == golden["probe_query"].tolist()
    assert [p["neighbor"] for p in probe] == golden["probe_neighbor"].tolist()
    assert np.abs(np.array([p["similarity"] for p in probe]) - golden["probe_similarity"]).max() <= 2e-6
    # duplicates: rows 5 == 400 and 650 == 17 -> the duplicate is the neighbour; explicit queries incl. first / last row
    q = [5, 400, 17, 650, 0, 699]
    rows, sims = eng.neighbor_probe(dev, q)
    unit = x / np.clip(np.linalg.norm(x, axis=1, keepdims=True), 1e-12, None)
    for j, qi in enumerate(q):
        s = unit[qi] @ unit.T
        s[qi] = -np.inf
        assert rows[j] == int(np.argmax(s)), (qi, rows[j], int(np.argmax(s)))
        assert abs(sims[j] - s[rows[j]]) <= 2e-6
    assert rows[0] == 400 and rows[1] == 5 and rows[2] == 650 and rows[3] == 17
    # three identical rows: the first maximum wins, as np.argmax
    y = x[:300].copy()
    y[100] = y[200] = y[7]
    rows, _ = eng.neighbor_probe(torch.from_numpy(y).cuda(), [7, 100, 200])
    assert rows.tolist() == [100, 7, 7]
    assert fx.nearest_neighbor_probe(dev[:1], recs[:1]) == []
    with pytest.raises(N.FxError):
        eng.neighbor_probe(dev, [700])


def test_standard_scaler_matches_reference_golden_and_sklearn(golden, eng):
    from sklearn.preprocessing import StandardScaler

    x = post_matrix()
    z, mean, scale = sf.fit_transform_device(torch.from_numpy(x).cuda())
    z = z.cpu().numpy()
    assert np.array_equal(mean.astype(np.float32), golden["scaler_mean"]) and np.array_equal(scale.astype(np.float32), golden["scaler_scale"])
    assert np.array_equal(z[::7, ::5], golden["features_sample"])  # fp32 IEEE arithmetic on equal fp32 mean / scale
    sk = StandardScaler()
    want = sk.fit_transform(x.astype(np.float32))
    assert np.allclose(mean, sk.mean_, rtol=1e-13, atol=0) and np.allclose(scale, sk.scale_, rtol=1e-10, atol=0)
    assert scale[3] == 1.0  # constant column
    assert np.array_equal(z, want)


def test_standardize_features_cli_writes_the_reference_bundle(golden, tmp_path):
    x = post_matrix()
    recs = _records(700)
    np.save(tmp_path / "embeddings.npy", x)
    pd.DataFrame({"index": range(700), "path": [str(r.relative_path) for r in recs], "bucket": [r.bucket for r in recs],
                  "label": [r.label or "" for r in recs]}).sample(frac=1.0, random_state=3).to_csv(tmp_path / "embeddings.csv", index=False)
    sf.main(["--embeddings-npy", str(tmp_path / "embeddings.npy"), "--embeddings-csv", str(tmp_path / "embeddings.csv"),
             "--output-npz", str(tmp_path / "out" / "std.npz"), "--log-level", "WARNING"])
    z = np.load(tmp_path / "out" / "std.npz", allow_pickle=True)
    assert sorted(z.files) == ["features", "is_labeled", "labels", "paths", "scaler_mean", "scaler_scale"]
    assert z["features"].dtype == np.float32 and z["features"].shape == (700, 512)
    assert np.array_equal(z["features"][::7, ::5], golden["features_sample"])
    assert np.array_equal(z["scaler_mean"], golden["scaler_mean"]) and np.array_equal(z["scaler_scale"], golden["scaler_scale"])
    assert np.array_equal(np.asarray(z["labels"], dtype=str), golden["labels"]) and np.array_equal(z["is_labeled"], golden["is_labeled"])
    assert z["paths"][0] == "img_0000.png"  # rows re-aligned on the explicit index column
    with pytest.raises(FileNotFoundError):
        sf.standardize_embeddings(tmp_path / "nope.npy", tmp_path / "embeddings.csv", tmp_path / "x.npz")


def test_large_matrix_known_answers(eng):
    """Size-independent properties at C3-like scale (400k rows = 0.8 GB): exact column means / variances of a
    constructed matrix, and a planted duplicate found among 400k rows."""
    n, d = 400_000, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((n, d), device="cuda", generator=g)
    x[123_459] = x[399_999]
    x[:, 0] = 2.0                                               # constant
    x[:, 1] = (torch.arange(n, device="cuda") % 4).float()      # mean 1.5, var 1.25 exactly
    st, mean, std, var = eng.column_stats(x)
    assert mean[0].item() == 2.0 and var[0].item() == 0.0
    assert mean[1].item() == 1.5 and var[1].item() == 1.25
    assert abs(mean[2:].mean().item() - 0.5) < 1e-3 and abs(var[2:].mean().item() - 1 / 12) < 1e-3
    rows, sims = eng.neighbor_probe(x, [399_999, 123_459, 7])
    assert rows[0] == 123_459 and rows[1] == 399_999 and abs(sims[0] - 1.0) < 1e-6
    z = eng.standardize(x, mean, torch.where(var == 0, torch.ones_like(std), std))
    assert abs(z[:, 5].mean().item()) < 1e-5 and abs(z[:, 5].std(unbiased=False).item() - 1.0) < 1e-5
    assert z[:, 0].abs().max().item() == 0.0


def test_device_streamed_npy_is_np_save_byte_for_byte(tmp_path):
    """embeddings.npy written from the gathered matrix while it is still on the GPU (two page-locked
    buffers, copy of chunk k+1 under the write of chunk k) == np.save(matrix.astype(float32)), src/feature_extraction.py:416."""
    from ssip_b200 import _artifacts as A

    for n, chunk in ((0, 1 << 20), (1, 1 << 20), (700, 64 * 2048), (700, 1 << 30), (4099, 3 * 2048 + 100)):
        x = post_matrix(max(n, 1))[:n] if n <= 700 else np.random.default_rng(n).random((n, 512), dtype=np.float32)
        dev = torch.from_numpy(x).to(DEV)
        A.write_npy(tmp_path / "dev.npy", dev, chunk_bytes=chunk)
        np.save(tmp_path / "ref.npy", x.astype(np.float32))
        assert (tmp_path / "dev.npy").read_bytes() == (tmp_path / "ref.npy").read_bytes(), (n, chunk)
    with pytest.raises(ValueError):
        A.write_npy(tmp_path / "bad.npy", torch.zeros(4, 512, dtype=torch.float16, device=DEV))
