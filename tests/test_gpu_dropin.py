"""The drop-in module (ssip_b200.feature_extraction) on a GPU: same entry points, bookkeeping, error
behaviour and artifacts as src/feature_extraction.py (reference lines cited per test), results within
BASELINE.json's tolerance of the oracle port run on the same files."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import reference_path as rp
from ssip_b200 import feature_extraction as fx
from ssip_b200 import synthetic
from ssip_b200.engine import pack_images

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = tmp_path_factory.mktemp("data")
    imgs = synthetic.ragged_images([(224, 224)] * 10 + [(300, 500), (512, 512), (640, 480)], seed=31)
    imgs += list(synthetic.mri_like_images(3, 512, seed=5))
    synthetic.write_png_dataset(root, imgs, n_labeled=6)
    (root / "sans_label" / "zz_broken.png").write_bytes(b"this is not an image")  # decode failure (:281-284)
    return root, imgs


@pytest.fixture(scope="module", autouse=True)
def weights_env():
    old = os.environ.get(fx.WEIGHTS_ENV)
    os.environ[fx.WEIGHTS_ENV] = "random-bn:1234"
    yield
    if old is None:
        os.environ.pop(fx.WEIGHTS_ENV, None)
    else:
        os.environ[fx.WEIGHTS_ENV] = old


def _oracle(records, batch_size):
    port = [rp.PortRecord(r.absolute_path, r.relative_path, r.bucket, r.label) for r in records]
    # fx._seeded_backbone(1234, True) uses BN seed 1235 == rp.BN_SEED
    return rp.port_extract_embeddings(port, torch.device("cpu"), batch_size=batch_size, seed=1234, randomize_bn=True)


def test_extract_embeddings_matches_oracle_and_keeps_bookkeeping(dataset):
    root, imgs = dataset
    records = fx.discover_image_records(root)  # :125-181
    assert len(records) == len(imgs) + 1
    res = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=5)  # ragged last batch, 2 slots in flight
    ref = _oracle(records, 5)
    assert [r.relative_path for r in res.records] == [r.relative_path for r in ref.records]  # order, failures removed (:295)
    assert [p.name for p in res.failures] == ["zz_broken.png"] == [p.name for p in ref.failures]
    assert res.embeddings.dtype == np.float32 and res.embeddings.shape == (len(imgs), 512)
    assert len(res.per_file_times) == len(imgs) and all(t > 0 for t in res.per_file_times)  # :297-300
    rel = np.linalg.norm(res.embeddings - ref.embeddings, axis=1) / np.linalg.norm(ref.embeddings, axis=1)
    cos = (res.embeddings * ref.embeddings).sum(1) / (np.linalg.norm(res.embeddings, axis=1) * np.linalg.norm(ref.embeddings, axis=1))
    assert rel.max() <= 1e-2 and cos.min() >= 0.999, (rel.max(), cos.min())
    # batch size must not change a single bit (determinism contract)
    res2 = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=16)
    assert np.array_equal(res.embeddings, res2.embeddings)


def test_main_writes_the_five_artifacts(dataset, tmp_path, monkeypatch):
    root, imgs = dataset
    monkeypatch.chdir(tmp_path)  # output paths are CWD-relative constants (:53-62)
    fx.main(["--data-dir", str(root), "--device", "cuda:0", "--batch-size", "8"])
    emb = np.load(tmp_path / "outputs/features/embeddings.npy")
    assert emb.shape == (len(imgs), 512) and emb.dtype == np.float32
    import pandas as pd

    frame = pd.read_csv(tmp_path / "outputs/features/embeddings.csv")
    assert list(frame.columns) == ["index", "path", "bucket", "label"] and len(frame) == len(imgs)  # :418-431
    assert list(frame["index"]) == list(range(len(imgs)))
    assert set(frame["bucket"]) == {"labeled", "unlabeled"}
    meta = json.loads((tmp_path / "outputs/features/metadata.json").read_text())
    for key in ("backbone", "weights", "layer", "embedding_dimension", "input_resize", "input_crop", "normalization_mean",
                "normalization_std", "channel_policy", "date_utc", "num_images", "failed_images", "device", "dataset_dir",
                "dataset_digest", "sanity_checks", "neighbor_probe"):  # :433-451
        assert key in meta, key
    assert meta["num_images"] == len(imgs) and meta["failed_images"] == 1 and meta["embedding_dimension"] == 512
    assert set(meta["sanity_checks"]) == {"num_vectors", "dimension", "mean_abs_mean", "mean_std"}
    note = (tmp_path / "outputs/notes/feature_summary.md").read_text()
    assert "# Feature Extraction Summary" in note and "zz_broken.png" in note
    assert (tmp_path / "outputs/logs/feature_extraction.log").exists()  # (logging.basicConfig is a no-op under pytest, as for the reference)


def test_build_transform_and_load_model_are_drop_ins(dataset):
    root, imgs = dataset
    transform = fx.build_transform()  # :184-207
    want_t = rp.port_transform()
    for arr in (imgs[0], imgs[10], imgs[11]):
        pil = Image.fromarray(arr)
        assert torch.equal(transform(pil), want_t(pil))  # bit-exact
    model = fx.load_model(torch.device("cuda:0"))  # :210-227
    x = torch.stack([want_t(Image.fromarray(a)) for a in imgs[:4]])
    with torch.no_grad():
        got = model(x)
        want = rp.port_model(torch.device("cpu"), 1234, True)(x)
    assert tuple(got.shape) == (4, 512, 1, 1)
    g, w = got.flatten(1).cpu().numpy(), want.flatten(1).numpy()
    assert (np.linalg.norm(g - w, axis=1) / np.linalg.norm(w, axis=1)).max() <= 1e-2


def test_channel_policy_errors_propagate_like_the_reference(tmp_path):
    gray = tmp_path / "avec_labels" / "a"
    gray.mkdir(parents=True)
    Image.fromarray(np.zeros((64, 64), np.uint8)).save(gray / "g.png")  # true mode "L": Normalize raises
    records = fx.discover_image_records(tmp_path)
    with pytest.raises(RuntimeError, match="broadcast shape"):
        fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=2)


def test_pipelined_slots_equal_the_synchronous_call():
    from ssip_b200 import _native as N

    eng = fx.get_engine(torch.device("cuda:0"), min_batch=16)
    batches = [list(synthetic.noise_images(16, 224, 224, seed=s)) for s in range(7)]
    sync = []
    packed = []
    for b in batches:
        buf, descs, total = pack_images(b)
        pinned = torch.from_numpy(buf.copy()).pin_memory()
        packed.append((pinned, descs, total))
        sync.append(eng.embed_host(pinned, descs, 16, total))
    for nslots in (2, N.HOST_SLOTS):  # slots cycle over both lanes; more batches than slots -> every slot is reused
        outs = [torch.empty((16, 512)).pin_memory() for _ in batches]
        for i, (pinned, descs, total) in enumerate(packed):
            eng.embed_host_wait(i % nslots)
            eng.embed_host_async(i % nslots, pinned, descs, 16, total, outs[i])
        for s in range(nslots):
            eng.embed_host_wait(s)
        for i in range(len(batches)):
            assert np.array_equal(outs[i].numpy(), sync[i])
    with pytest.raises(N.FxError):
        eng.embed_host_async(N.HOST_SLOTS, packed[0][0], packed[0][1], 16, packed[0][2], outs[0])
    # uniform batches travel as one 2-D copy of the source rows the crop can touch (224 -> 256 -> centre 224: rows 14..211);
    # ragged batches as the whole buffer.  Either way the rows equal the device-resident path's (inputs fully on the device).
    from ssip_b200.engine import uniform_descs

    x = synthetic.noise_images(16, 224, 224, seed=99)
    pinned = torch.from_numpy(x.reshape(-1).copy()).pin_memory()
    before = eng.h2d_bytes
    got = eng.embed_host(pinned, uniform_descs(16, 224, 224), 16, x.size)
    copied = eng.h2d_bytes - before
    assert 0.85 * x.size < copied < 0.92 * x.size, (copied, x.size)
    want = eng.embed_device(torch.from_numpy(x.reshape(-1)).cuda(), uniform_descs(16, 224, 224), 16).cpu().numpy()
    assert np.array_equal(got, want)
    ragged = synthetic.ragged_images([(224, 224), (300, 500)], seed=5)
    buf, descs, total = pack_images(ragged)
    before = eng.h2d_bytes
    eng.embed_host(buf, descs, 2, total)
    assert eng.h2d_bytes - before == total


def test_config1_256_png_images_batch_32(tmp_path_factory):
    # BASELINE.json configs[0]: 256 synthetic 224x224 uint8 images, batch 32, against the CPU fp32 path on the same files
    root = tmp_path_factory.mktemp("c1")
    imgs = list(synthetic.noise_images(256, 224, 224, seed=0))
    synthetic.write_png_dataset(root, imgs, n_labeled=32)
    records = fx.discover_image_records(root)
    res = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=32)
    ref = _oracle(records, 32)
    assert res.embeddings.shape == (256, 512) and not res.failures
    rel = np.linalg.norm(res.embeddings - ref.embeddings, axis=1) / np.linalg.norm(ref.embeddings, axis=1)
    cos = (res.embeddings * ref.embeddings).sum(1) / (np.linalg.norm(res.embeddings, axis=1) * np.linalg.norm(ref.embeddings, axis=1))
    assert rel.max() <= 1e-2 and cos.min() >= 0.999, (rel.max(), cos.min())
    # the sanity statistics the reference logs (src/feature_extraction.py:334-356) agree to the same tolerance
    a, b = fx.run_sanity_checks(res.embeddings), fx.run_sanity_checks(ref.embeddings)
    assert abs(a["mean_abs_mean"] - b["mean_abs_mean"]) <= 1e-2 * b["mean_abs_mean"]
    assert abs(a["mean_std"] - b["mean_std"]) <= 2e-2 * b["mean_std"]


def test_size_independent_properties_at_scale():
    # 4096 device-resident images, batch 256 (the config-2 shape): rows depend on their own image only, so
    # (a) a permutation of the inputs permutes the rows bit-exactly, (b) duplicated images give identical rows,
    # (c) every value is finite.  No oracle needed at this size.
    from ssip_b200.engine import uniform_descs

    eng = fx.get_engine(torch.device("cuda:0"), min_batch=256)
    n, b = 4096, 256
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(0, 256, (n, 224 * 224 * 3), dtype=torch.uint8, device="cuda", generator=gen)
    x[1000] = x[7]  # duplicates in different batches and positions
    x[4095] = x[7]
    descs = uniform_descs(b, 224, 224)

    def run(t):
        out = torch.empty((n, 512), dtype=torch.float32, device="cuda")
        for s in range(0, n, b):
            eng.embed_device(t[s : s + b].reshape(-1), descs, b, out=out[s : s + b])
        torch.cuda.synchronize()
        return out

    base = run(x)
    # batch 64 takes different tilings (no split tail wave in layer4, other M-tile boxes): still the same bits,
    # and batch-64 results are the ones checked against the oracle elsewhere
    d64 = uniform_descs(64, 224, 224)
    small = torch.empty((512, 512), dtype=torch.float32, device="cuda")
    for s0 in range(0, 512, 64):
        eng.embed_device(x[s0 : s0 + 64].reshape(-1), d64, 64, out=small[s0 : s0 + 64])
    torch.cuda.synchronize()
    assert torch.equal(small, base[:512])
    perm = torch.randperm(n, device="cuda", generator=gen)
    assert torch.equal(run(x[perm].contiguous()), base[perm])
    assert torch.equal(base[1000], base[7]) and torch.equal(base[4095], base[7])
    assert bool(torch.isfinite(base).all())
    assert float((base[0] - base[1]).abs().max()) > 0  # different images do differ


def test_process_decode_pool_gives_the_same_result_as_threads(dataset, tmp_path, monkeypatch):
    """worker processes decoding into the shared page-locked staging buffer change nothing but
    the speed -- same rows, same failures (incl. a file whose header parses but whose pixel data is truncated), same
    exception for a channel count the reference cannot normalise."""
    from PIL import Image

    root, imgs = dataset
    full = tmp_path / "data"
    import shutil

    shutil.copytree(root, full)
    Image.fromarray(imgs[0]).save(full / "sans_label" / "a_good.jpg", quality=90)
    data = (full / "sans_label" / "a_good.jpg").read_bytes()
    (full / "sans_label" / "b_truncated.jpg").write_bytes(data[: len(data) // 3])
    records = fx.discover_image_records(full)
    monkeypatch.setenv(fx.DECODE_MODE_ENV, "thread")
    a = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=6)
    monkeypatch.setenv(fx.DECODE_MODE_ENV, "process")
    monkeypatch.setenv(fx.DECODE_THREADS_ENV, "4")
    b = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=6)
    assert np.array_equal(a.embeddings, b.embeddings)
    assert [r.relative_path for r in a.records] == [r.relative_path for r in b.records]
    # ... and both are the reference's rows: the oracle port on the very same files (JPEG included)
    ref = _oracle(records, 6)
    assert [r.relative_path for r in b.records] == [r.relative_path for r in ref.records]
    assert sorted(p.name for p in ref.failures) == ["b_truncated.jpg", "zz_broken.png"]
    rel = np.linalg.norm(b.embeddings - ref.embeddings, axis=1) / np.linalg.norm(ref.embeddings, axis=1)
    cos = (b.embeddings * ref.embeddings).sum(1) / (np.linalg.norm(b.embeddings, axis=1) * np.linalg.norm(ref.embeddings, axis=1))
    assert rel.max() <= 1e-2 and cos.min() >= 0.999, (rel.max(), cos.min())
    assert sorted(p.name for p in a.failures) == sorted(p.name for p in b.failures) == ["b_truncated.jpg", "zz_broken.png"]
    monkeypatch.setenv(fx.GRAY_CARRIAGE_ENV, "1")  # R==G==B files travel as one plane: same embeddings
    c = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=6)
    assert np.array_equal(a.embeddings, c.embeddings)
    monkeypatch.delenv(fx.GRAY_CARRIAGE_ENV)
    Image.fromarray(imgs[1][..., 0]).save(full / "sans_label" / "c_gray.png")  # true mode "L": the reference raises
    with pytest.raises(RuntimeError, match="broadcast shape"):
        fx.extract_embeddings(fx.discover_image_records(full), torch.device("cuda:0"), batch_size=6)


def test_real_mri_files_match_the_reference_goldens(golden_dir, monkeypatch):
    """16 JPEGs of the reference's own dataset (tests/golden/mri_real, minted by make_golden_mri.py from
    the unmodified src.feature_extraction): the fused transform is bit-exact on every file, the embeddings are within
    BASELINE.json's tolerance of the reference's, in bf16 and in the tight mode, with thread and process decode."""
    import hashlib

    g = np.load(golden_dir / "mri_real_golden.npz")
    records = fx.discover_image_records(golden_dir / "mri_real")
    assert [str(r.relative_path) for r in records] == g["paths"].tolist()
    transform = fx.build_transform()
    for r, pre in zip(records, g["pre_sha256"]):
        t = fx.preprocess_image(r.absolute_path, transform)
        assert hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest() == str(pre)
    want = g["emb_randbn"]  # weights_env: random-bn:1234

    def check(res, tol, cos_tol):
        assert not res.failures and [r.relative_path for r in res.records] == [r.relative_path for r in records]
        rel = np.linalg.norm(res.embeddings - want, axis=1) / np.linalg.norm(want, axis=1)
        cos = (res.embeddings * want).sum(1) / (np.linalg.norm(res.embeddings, axis=1) * np.linalg.norm(want, axis=1))
        assert rel.max() <= tol and cos.min() >= cos_tol, (rel.max(), cos.min())

    a = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=5)
    check(a, 1e-2, 0.999)
    monkeypatch.setenv(fx.DECODE_MODE_ENV, "process")
    monkeypatch.setenv(fx.DECODE_THREADS_ENV, "4")
    b = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=16)
    assert np.array_equal(a.embeddings, b.embeddings)
    monkeypatch.setenv(fx.DECODE_MODE_ENV, "thread")
    monkeypatch.setenv(fx.PRECISION_ENV, "fp32")
    check(fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=8), 1e-5, 0.999999)


def test_oversized_file_path_is_bit_identical_to_the_kernel_path(dataset, monkeypatch):
    """Lowering the host-resize threshold sends the 300x500 / 512x512 / 640x480 files through Pillow's resize on the host
    (the reference's own call) instead of the kernel's resampler: the embeddings must not change by a single bit -- which
    also says the kernel's integer resampler IS Pillow's."""
    from ssip_b200 import _decode_pool as dp

    root, imgs = dataset
    records = fx.discover_image_records(root)
    monkeypatch.setenv(fx.DECODE_MODE_ENV, "thread")
    a = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=8)
    monkeypatch.setattr(dp, "HOST_RESIZE_SHORT_SIDE", 280)
    monkeypatch.setattr(fx, "HOST_RESIZE_SHORT_SIDE", 280)
    b = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=8)
    assert np.array_equal(a.embeddings, b.embeddings)
    assert [r.relative_path for r in a.records] == [r.relative_path for r in b.records]
