"""CPU-side checks of the product: the C-ABI library loads and exports what include/fx_b200.h
declares, its host-side integer code (coefficient tables, resize size, crop offsets) equals the
oracle's, the drop-in module keeps the reference's surface and error behaviour, the N>1 sharding /
gather logic works on gloo with world_size 2.  No compute call needs a GPU here."""
import ctypes
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import reference_path as rp
from ssip_b200 import _native as N
from ssip_b200 import dist as fxdist
from ssip_b200 import engine as E
from ssip_b200 import feature_extraction as fx
from ssip_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "fx_b200.h").read_text()
    declared = set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", header)) - {"fx_handle"}
    assert declared == set(N.EXPORTED_SYMBOLS)
    lib = N.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fx_abi_version() == 1
    assert b"sm_100a" in lib.fx_version()


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.FxError) as err:
        E.Engine(0, 8)
    assert err.value.status == N.FX_ERR_UNSUPPORTED and "no CPU path" in err.value.text
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fx.extract_embeddings([], torch.device("cpu"), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fx.load_model(torch.device("cpu"))


def test_product_never_imports_the_oracle():
    pkg = ROOT / "semi-supervised-image-processing_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        assert not re.search(r"import\s+oracle|from\s+oracle|oracle/|oracle\.|libfx_oracle|fxo_", text), path


@pytest.mark.parametrize("sizes", [(512, 256), (224, 256), (300, 256), (500, 426), (1000, 365), (2048, 256), (100, 256), (256, 256), (7, 256)])
def test_host_coefficients_equal_oracle(sizes):
    xmin, cnt, taps = E.host_coeffs(*sizes)
    bounds, kk = rp.c_coeffs(*sizes)
    np.testing.assert_array_equal(xmin, bounds[:, 0])
    np.testing.assert_array_equal(cnt, bounds[:, 1])
    np.testing.assert_array_equal(taps, kk)
    assert (taps.sum(axis=1) - (1 << 22)).__abs__().max() <= taps.shape[1]  # normalised rows


def test_host_resize_and_crop_geometry_equal_torchvision():
    from torchvision.transforms import functional as F

    for (h, w) in [(224, 224), (512, 512), (300, 500), (500, 300), (514, 512), (777, 333), (1000, 700), (1, 1), (3000, 17)]:
        want = F._compute_resized_output_size((h, w), [256])
        assert list(E.host_resized_size(h, w)) == list(want)
    for size in range(224, 1200):
        assert E.host_crop_offset(size) == int(round((size - 224) / 2.0))


def test_pack_images_layout():
    imgs = synthetic.ragged_images([(5, 7), (3, 2)], seed=1) + [np.zeros((4, 4), np.uint8)]
    buf, descs, total = E.pack_images(imgs)
    assert total % 256 == 0 and descs[0].offset == 0 and descs[1].offset == 256 and descs[2].offset == 512
    assert (descs[2].channels, descs[1].height, descs[1].width) == (1, 3, 2)
    np.testing.assert_array_equal(buf[256 : 256 + 18].reshape(3, 2, 3), imgs[1])
    with pytest.raises(TypeError):
        E.pack_images([np.zeros((4, 4, 3), np.float32)])


def test_layer_table_matches_torchvision_state_dict():
    state = rp.make_backbone().state_dict()
    shapes = [tuple(state[k].shape) for k, *_ in E.LAYER_TABLE]
    assert shapes[0] == (64, 3, 7, 7) and shapes[7] == (128, 64, 1, 1) and shapes[19] == (512, 512, 3, 3)
    macs = 0
    hw = [112] + [56] * 4 + [28] * 5 + [14] * 5 + [7] * 5
    for (co, ci, kh, kw), o in zip(shapes, hw):
        macs += co * ci * kh * kw * o * o
    assert macs == 1_813_561_344


def _make_dataset(tmp_path, n=6):
    imgs = list(synthetic.noise_images(n, 32, 40, seed=3))
    synthetic.write_png_dataset(tmp_path, imgs, n_labeled=4)
    return imgs


def test_discover_image_records_contract(tmp_path, have_reference):
    _make_dataset(tmp_path)
    (tmp_path / "sans_label" / "notes.txt").write_text("not an image")
    recs = fx.discover_image_records(tmp_path)
    assert [r.bucket for r in recs] == ["labeled"] * 4 + ["unlabeled"] * 3
    assert [r.label for r in recs] == ["cancer", "cancer", "normal", "normal", None, None, None]
    assert str(recs[-1].relative_path) == "sans_label/notes.txt"  # every file is a record
    with pytest.raises(FileNotFoundError):
        fx.discover_image_records(tmp_path / "missing")
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(RuntimeError, match="No image files"):
        fx.discover_image_records(empty)
    if have_reference:
        sys.path.insert(0, "/root/reference")
        import src.feature_extraction as fe

        ref = fe.discover_image_records(tmp_path)
        assert [(r.relative_path, r.bucket, r.label) for r in ref] == [(r.relative_path, r.bucket, r.label) for r in recs]


def test_module_surface_matches_reference(have_reference):
    names = ["discover_image_records", "build_transform", "load_model", "preprocess_image", "batched", "extract_embeddings",
             "compute_dataset_digest", "run_sanity_checks", "nearest_neighbor_probe", "save_artifacts", "configure_logging",
             "parse_args", "main", "ImageRecord", "ExtractionResults"]
    for name in names:
        assert hasattr(fx, name), name
    consts = ["IMAGENET_MEAN", "IMAGENET_STD", "TARGET_RESIZE", "TARGET_CROP", "BATCH_SIZE", "NEIGHBOR_SAMPLE", "RNG_SEED",
              "LABELED_BUCKET", "UNLABELED_BUCKET", "BACKBONE_NAME", "BACKBONE_WEIGHTS", "BACKBONE_LAYER", "EMBEDDING_ARRAY_PATH",
              "EMBEDDING_CSV_PATH", "METADATA_PATH", "SUMMARY_NOTE_PATH", "LOG_PATH", "DEFAULT_DATA_DIR"]
    if have_reference:
        sys.path.insert(0, "/root/reference")
        import src.feature_extraction as fe

        for c in consts:
            assert getattr(fx, c) == getattr(fe, c), c
    args = fx.parse_args(["--data-dir", "d", "--batch-size", "7", "--verbose", "--device", "cuda:3"])
    assert (str(args.data_dir), args.batch_size, args.verbose, args.device) == ("d", 7, True, "cuda:3")
    assert list(fx.batched(list(range(7)), 3)) == [[0, 1, 2], [3, 4, 5], [6]]


def test_postprocessing_and_artifacts(tmp_path, monkeypatch, have_reference):
    _make_dataset(tmp_path)
    recs = fx.discover_image_records(tmp_path)
    emb = np.random.default_rng(0).random((len(recs), 512), dtype=np.float32)
    emb[3] = emb[1]  # a duplicate must be its twin's nearest neighbour
    res = fx.ExtractionResults(emb, recs, [tmp_path / "broken.jpg"], [0.01] * len(recs))
    stats = fx.run_sanity_checks(emb)
    probe = fx.nearest_neighbor_probe(emb, recs)
    assert len(probe) == 5 and all(set(p) == {"query", "neighbor", "similarity"} for p in probe)
    bad = emb.copy()
    bad[0, 0] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        fx.run_sanity_checks(bad)
    bad[0, 0] = np.inf
    with pytest.raises(ValueError, match="inf"):
        fx.run_sanity_checks(bad)
    monkeypatch.chdir(tmp_path)
    fx.save_artifacts(res, stats, probe, tmp_path, torch.device("cuda:0"))
    saved = np.load(tmp_path / fx.EMBEDDING_ARRAY_PATH)
    assert saved.dtype == np.float32 and np.array_equal(saved, emb)
    import pandas as pd

    frame = pd.read_csv(tmp_path / fx.EMBEDDING_CSV_PATH)
    assert list(frame.columns) == ["index", "path", "bucket", "label"] and len(frame) == len(recs)
    assert frame["label"].isna().sum() == 2
    meta = json.loads((tmp_path / fx.METADATA_PATH).read_text())
    assert list(meta) == ["backbone", "weights", "layer", "embedding_dimension", "input_resize", "input_crop", "normalization_mean",
                          "normalization_std", "channel_policy", "date_utc", "num_images", "failed_images", "device", "dataset_dir",
                          "dataset_digest", "sanity_checks", "neighbor_probe"]
    assert meta["failed_images"] == 1 and meta["num_images"] == len(recs) and meta["device"] == "cuda:0"
    note = (tmp_path / fx.SUMMARY_NOTE_PATH).read_text()
    assert "- Batch size: 32" in note and "- Failed decodes: 1" in note and "| Query | Neighbor | Cosine |" in note
    if have_reference:
        # identical numbers / text from the reference's own post-processing on the same inputs
        sys.path.insert(0, "/root/reference")
        import src.feature_extraction as fe

        ref_recs = fe.discover_image_records(tmp_path)
        assert fe.run_sanity_checks(emb) == stats
        assert fe.nearest_neighbor_probe(emb, ref_recs) == probe
        assert fe.compute_dataset_digest(ref_recs) == meta["dataset_digest"]
        ours = {p: (tmp_path / p).read_text() for p in (fx.SUMMARY_NOTE_PATH, fx.EMBEDDING_CSV_PATH)}
        fe.save_artifacts(fe.ExtractionResults(emb, ref_recs, [tmp_path / "broken.jpg"], [0.01] * len(recs)), stats, probe, tmp_path,
                          torch.device("cuda:0"))
        for p, text in ours.items():
            assert (tmp_path / p).read_text() == text, p
        ref_meta = json.loads((tmp_path / fx.METADATA_PATH).read_text())
        ref_meta.pop("date_utc"), meta.pop("date_utc")
        assert ref_meta == meta
        # the downstream consumer accepts the files
        from src.standardize_features import standardize_embeddings

        bundle = tmp_path / "outputs" / "features" / "standardized_features.npz"
        standardize_embeddings(tmp_path / fx.EMBEDDING_ARRAY_PATH, tmp_path / fx.EMBEDDING_CSV_PATH, bundle)
        assert bundle.exists()


def test_decode_rules_follow_reference_channel_policy(tmp_path):
    from PIL import Image

    rgb = tmp_path / "a.png"
    Image.fromarray(synthetic.noise_images(1, 8, 8)[0]).save(rgb)
    with Image.open(rgb) as im:
        assert fx._decoded_array(im).shape == (8, 8, 3)
    gray = tmp_path / "g.png"
    Image.fromarray(np.zeros((8, 8), np.uint8)).save(gray)
    with Image.open(gray) as im, pytest.raises(RuntimeError, match="broadcast shape"):
        fx._decoded_array(im)
    rgba = tmp_path / "r.png"
    Image.fromarray(np.zeros((8, 8, 4), np.uint8)).save(rgba)
    with Image.open(rgba) as im, pytest.raises(RuntimeError):
        fx._decoded_array(im)
    deep = tmp_path / "d.png"
    Image.fromarray(np.zeros((8, 8), np.uint16)).save(deep)
    with Image.open(deep) as im, pytest.raises(TypeError):
        fx._decoded_array(im)
    junk = tmp_path / "junk.jpg"
    junk.write_bytes(b"not a jpeg")
    assert isinstance(fx._load_file(junk), Exception)


def test_shard_bounds_cover_everything_in_order():
    for n in (0, 1, 7, 8, 9, 1000, 1506):
        for size in (1, 2, 3, 4, 8):
            spans = [fxdist.shard_bounds(n, r, size) for r in range(size)]
            flat = [i for lo, hi in spans for i in range(lo, hi)]
            assert flat == list(range(n))


_GLOO_WORKER = r"""
import os, sys, torch, numpy as np
sys.path.insert(0, {root!r})
from ssip_b200 import dist as fxdist
assert fxdist.ensure_process_group("gloo")
rank, size = fxdist.world()
n = 11
lo, hi = fxdist.shard_bounds(n, rank, size)
rows = torch.arange(n * 512, dtype=torch.float32).reshape(n, 512)
local = rows[lo:hi]
if rank == 1:
    local = local[:-2]   # two decode failures on rank 1 -> uneven counts
full = fxdist.allgather_rows(local.contiguous())
meta = fxdist.allgather_objects(list(range(lo, hi))[: local.shape[0]])
kept = fxdist.concat_in_rank_order(meta)
assert kept == [0, 1, 2, 3, 4, 5, 6, 7, 8], kept
assert torch.equal(full, rows[kept]), (full.shape,)
empty = fxdist.allgather_rows(torch.zeros((0, 512)))
assert empty.shape == (0, 512)
# the in-place form the drop-in uses: every rank's rows already sit in its slot of one [size * cap, D] buffer
cap = fxdist.shard_bounds(n, 0, size)[1]
for drop in (0, 2):  # full shards (result is a view of the buffer), then two decode failures on rank 0 (compaction)
    buf = torch.full((size * cap, 512), float("nan"))
    mine = rows[lo:hi]
    if rank == 0 and drop:
        mine = mine[:-drop]
    buf[rank * cap : rank * cap + mine.shape[0]] = mine
    full, counts = fxdist.allgather_inplace(buf, cap, mine.shape[0])
    want_counts = [cap - drop, n - cap]
    assert counts == want_counts, counts
    want_rows = list(range(0, cap - drop)) + list(range(cap, n))
    assert torch.equal(full, rows[want_rows]), (drop, full.shape)
    if not drop:
        assert full.data_ptr() == buf.data_ptr()  # no copy when the shards are full
torch.distributed.barrier()
torch.distributed.destroy_process_group()
sys.stdout.write("rank %d ok\n" % rank); sys.stdout.flush()
"""


def test_allgather_rows_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29611", str(script)],
        capture_output=True, text=True, timeout=240, env=env,
    )
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2 and "0" in out.stdout and "1" in out.stdout


def test_inference_drop_ins_keep_the_reference_signatures(have_reference):
    """Classifier-head inference loops: same names, parameters and defaults as the reference."""
    import inspect
    import types

    from ssip_b200 import inference

    for name in ("generate_pseudo_labels", "evaluate_model", "compute_probs", "build_transforms"):
        assert callable(getattr(inference, name))
    with pytest.raises(RuntimeError):  # no CPU path
        inference.get_classifier(torch.nn.Linear(1, 1), torch.device("cpu"))
    if not have_reference:
        return
    for name in ("matplotlib", "matplotlib.pyplot"):  # not installed here; only the plotting helpers use it
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, "/root/reference/src")
    import src.training.common as ref_common
    import src.training.semi_supervised as ref_semi
    import threshold_sweep as ref_sweep

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]

    assert params(inference.generate_pseudo_labels) == params(ref_semi.generate_pseudo_labels)
    assert params(inference.evaluate_model) == params(ref_common.evaluate_model)
    assert params(inference.compute_probs) == params(ref_sweep.compute_probs)
    assert params(inference.build_transforms) == params(ref_common.build_transforms)


def test_decode_pool_processes_write_pillow_pixels_into_shared_memory(tmp_path):
    """SURVEY.md 8f rank 1 (host decode pool): worker processes decode with the reference's own Pillow call straight
    into the shared staging buffer; failures come back per file with the reference's two tolerated exception types
    marked as such (src/feature_extraction.py:281-284)."""
    from PIL import Image

    from ssip_b200._decode_pool import DecodePool

    imgs = synthetic.ragged_images([(64, 80), (224, 224), (100, 130)], seed=8) + [synthetic.mri_like_images(1, 128, seed=2)[0]]
    paths = []
    for i, a in enumerate(imgs):
        p = tmp_path / f"ok_{i}.{'png' if i % 2 else 'jpg'}"
        Image.fromarray(a).save(p, quality=92) if p.suffix == ".jpg" else Image.fromarray(a).save(p)
        paths.append(str(p))
    (tmp_path / "junk.png").write_bytes(b"not an image")
    full = (tmp_path / "ok_0.jpg").read_bytes()
    (tmp_path / "cut.jpg").write_bytes(full[: len(full) // 3])  # header parses, pixel data is truncated
    Image.fromarray(imgs[0][..., 0]).save(tmp_path / "gray.png")  # mode L: the caller must raise like the reference
    paths += [str(tmp_path / "junk.png"), str(tmp_path / "cut.jpg"), str(tmp_path / "gray.png"), str(tmp_path / "missing.png")]
    pool = DecodePool(3, 2)
    try:
        metas = pool.probe(paths)
        assert metas[0] == (64, 80, 3, "RGB") and metas[3] == (128, 128, 3, "RGB")
        assert metas[4][:2] == ("decode", "UnidentifiedImageError") and metas[7][:2] == ("decode", "FileNotFoundError")
        assert metas[6] == (64, 80, 1, "L")
        with pytest.raises(RuntimeError):
            fx._check_mode("L", 1)
        jobs, off = [], 0
        for p, m in zip(paths[:4] + [paths[5]], metas[:4] + [metas[5]]):
            jobs.append((p, off, m[0], m[1], m[2], p.endswith("ok_3.png")))  # gray carriage for the R==G==B MRI-like image
            off += (m[0] * m[1] * m[2] + 255) // 256 * 256
        shm, created = pool.buffer(1, off)
        assert created and shm.size >= off
        res = pool.decode(1, jobs)
        assert res[:3] == [3, 3, 3] and res[3] == 1  # the gray image went over as one plane
        assert res[4][0] == "decode" and res[4][1] == "OSError"  # truncated JPEG: Pillow raises OSError at load
        buf = np.frombuffer(shm.buf, dtype=np.uint8)
        for (p, o, h, w, b, g), r in zip(jobs[:4], res[:4]):
            want = np.asarray(Image.open(p))
            want = want[..., 0] if r == 1 else want
            assert np.array_equal(buf[o : o + want.size].reshape(want.shape), want)
        del buf
        assert pool.buffer(1, off // 2)[1] is False  # reused, not re-created
    finally:
        pool.close()


# ---- artifact writers at scale: same bytes as the reference's np.save / pandas / serial stat ----
def test_streamed_npy_is_np_save_byte_for_byte(tmp_path):
    from ssip_b200 import _artifacts as A

    rng = np.random.default_rng(5)
    cases = [rng.random((0, 512), dtype=np.float32), rng.random((1, 512), dtype=np.float32), rng.random((1000, 512), dtype=np.float32),
             rng.random((37, 512)).astype(np.float64), np.asfortranarray(rng.random((64, 512), dtype=np.float32)),
             rng.random((130, 512), dtype=np.float32)[::2]]
    for k, emb in enumerate(cases):
        ours, ref = tmp_path / f"o{k}.npy", tmp_path / f"r{k}.npy"
        A.write_npy(ours, emb, chunk_bytes=(1 << 16) + 12)  # a chunk size that does not divide a row
        np.save(ref, emb.astype(np.float32))                # src/feature_extraction.py:416
        assert ours.read_bytes() == ref.read_bytes(), k
    A.write_npy(tmp_path / "t.npy", torch.from_numpy(cases[2]))
    assert (tmp_path / "t.npy").read_bytes() == (tmp_path / "r2.npy").read_bytes()


def test_csv_writer_is_pandas_to_csv_byte_for_byte(tmp_path):
    import pandas as pd

    from ssip_b200 import _artifacts as A

    names = ["plain.jpg", "with,comma.png", 'quote"inside.jpg', "new\nline.jpg", "  spaces  .jpg", "ünï/cödé.jpg", "'single'.jpg", "semi;colon.jpg",
             "tab\there.jpg", "cr\rhere.jpg", "", "None", "nan"]
    recs = [fx.ImageRecord(tmp_path / n, Path("avec_labels") / "cancer" / n if i % 3 else Path(n), "labeled" if i % 2 else "unlabeled",
                           ("cancer" if i % 4 else 'odd,"label"') if i % 2 else None) for i, n in enumerate(names)]
    for rows in (recs, recs[:1], [r for r in recs if r.label is None], []):
        ours, ref = tmp_path / "o.csv", tmp_path / "r.csv"
        A.write_embeddings_csv(ours, rows)
        # the reference's construction, src/feature_extraction.py:418-431
        frame = pd.DataFrame([{"index": i, "path": str(r.relative_path), "bucket": r.bucket, "label": r.label} for i, r in enumerate(rows)])
        if rows:
            frame.to_csv(ref, index=False)
            assert ours.read_bytes() == ref.read_bytes()
        else:  # the reference raises before it gets here (no embeddings); only the header makes sense
            assert ours.read_text() == "index,path,bucket,label\n"


def test_digest_equals_the_reference_loop(tmp_path):
    import hashlib

    from ssip_b200 import _artifacts as A

    recs = []
    for i in range(300):
        f = tmp_path / f"f{i:04d}.bin"
        f.write_bytes(b"x" * (i % 17))
        os.utime(f, (1_700_000_000 + i, 1_700_000_000 + i * 3 + 0.75))
        recs.append(fx.ImageRecord(f, Path("sans_label") / f.name, "unlabeled", None))
    recs = recs[::-1]  # the digest sorts by relative path itself
    h = hashlib.sha256()
    for r in sorted(recs, key=lambda r: str(r.relative_path)):  # src/feature_extraction.py:326-330
        st = r.absolute_path.stat()
        h.update(str(r.relative_path).encode("utf-8"))
        h.update(str(st.st_size).encode("utf-8"))
        h.update(str(int(st.st_mtime)).encode("utf-8"))
    assert A.dataset_digest(recs) == h.hexdigest() == fx.compute_dataset_digest(recs)
    (tmp_path / "f0007.bin").unlink()
    with pytest.raises(FileNotFoundError):  # as the reference: a vanished file aborts the run
        A.dataset_digest(recs)


def test_oversized_files_are_resized_on_the_host_exactly_as_the_reference_would(tmp_path, monkeypatch):
    """Files whose short side exceeds what the preprocess kernel's band buffer takes are resized right after the decode with
    the reference's own call (torchvision functional.resize -> PIL resize(BILINEAR), src/feature_extraction.py:200-203); the
    kernel then sees a 256-short-side image, for which Resize(256) is the identity.  Threshold lowered here to keep it small."""
    from PIL import Image
    from torchvision.transforms import functional as TF

    from ssip_b200 import _decode_pool as dp

    monkeypatch.setattr(dp, "HOST_RESIZE_SHORT_SIDE", 300)
    for h, w in [(640, 480), (333, 777), (301, 301)]:
        arr = synthetic.ragged_images([(h, w)], seed=h)[0]
        img = Image.fromarray(arr)
        got = dp.host_resize_if_oversized(img)
        want = TF.resize(img, 256)
        assert got.size == want.size == dp.resized_size(h, w)[::-1]
        assert np.array_equal(np.asarray(got), np.asarray(want))
        # and the whole reference transform of the pre-resized image is the transform of the original
        t = rp_transform()
        assert torch.equal(t(got), t(img))
    small = Image.fromarray(synthetic.ragged_images([(300, 500)], seed=1)[0])
    assert dp.host_resize_if_oversized(small) is small
    p = tmp_path / "big.png"
    Image.fromarray(synthetic.ragged_images([(640, 480)], seed=2)[0]).save(p)
    assert dp._probe_chunk([str(p)])[0][:2] == (341, 256)  # the header pass reports what the decode pass will deliver


def rp_transform():
    from oracle import reference_path as rp

    return rp.port_transform()


def test_worker_exceptions_come_back_as_their_own_class():
    from PIL import Image, UnidentifiedImageError

    from ssip_b200._decode_pool import rebuild_exception

    assert type(rebuild_exception("ValueError", "x")) is ValueError
    assert type(rebuild_exception("MemoryError", "x")) is MemoryError
    assert type(rebuild_exception("DecompressionBombError", "x")) is Image.DecompressionBombError
    assert type(rebuild_exception("UnidentifiedImageError", "x")) is UnidentifiedImageError
    assert type(rebuild_exception("NoSuchThing", "x")) is RuntimeError and "NoSuchThing" in str(rebuild_exception("NoSuchThing", "x"))


def test_cli_fan_out_starts_one_worker_per_visible_gpu(monkeypatch, tmp_path):
    """`--device cuda` on a multi-GPU box = all of them: the parent starts WORLD_SIZE workers with the
    torchrun environment and the same arguments, and reports the first failure.  (Workers replaced by a stub here.)"""
    import subprocess as sp

    calls = []

    class FakeProc:
        def __init__(self, cmd, env):
            calls.append((cmd, {k: env[k] for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}, env["PYTHONPATH"]))
            self.rank = int(env["RANK"])

            self.terminated = False

        def poll(self):
            return 3 if self.rank == 2 else (None if not self.terminated and self.rank == 3 else 0)  # rank 3 would hang in the collective

        def terminate(self):
            self.terminated = True

        def wait(self, timeout=None):
            return self.poll()

    monkeypatch.setattr(sp, "Popen", lambda cmd, env=None: FakeProc(cmd, env))
    rc = fx._fan_out(["--data-dir", str(tmp_path), "--device", "cuda", "--batch-size", "64"], 4)
    assert rc == 3 and len(calls) == 4
    assert [c[1]["RANK"] for c in calls] == ["0", "1", "2", "3"] and {c[1]["WORLD_SIZE"] for c in calls} == {"4"}
    assert len({c[1]["MASTER_PORT"] for c in calls}) == 1 and calls[0][1]["MASTER_ADDR"] == "127.0.0.1"
    assert calls[0][0][1:3] == ["-m", "ssip_b200.feature_extraction"] and calls[0][0][-2:] == ["--batch-size", "64"]
    assert str(ROOT) in calls[0][2].split(os.pathsep)
