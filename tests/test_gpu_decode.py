"""GPU JPEG decode feeding the preprocess kernel (replaces Image.open + Pillow's decode at
src/feature_extraction.py:238 for baseline RGB JPEGs).  nvJPEG's IDCT / chroma up-sampling are not libjpeg-turbo's, so
the bar is a tolerance, stated here and measured on the reference's 1506 MRI files in profiles/r02_nvjpeg_tolerance.md:
per pixel |difference| <= 6 with a mean <= 0.05 on photographic / MRI content, embeddings within BASELINE.json's
relative L2 <= 1e-2 and cosine >= 0.999 of the REFERENCE's rows (goldens of the real src.feature_extraction on the same
files).  Everything nvJPEG does not take goes through Pillow: identical rows, identical failure handling."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import reference_path as rp
from ssip_b200 import _native as N
from ssip_b200 import feature_extraction as fx
from ssip_b200 import synthetic
from ssip_b200.engine import Engine

pytestmark = pytest.mark.gpu


def _jpeg(arr, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", **kw)
    return buf.getvalue()


@pytest.fixture(scope="module")
def eng():
    e = Engine(0, max_batch=64, precision="bf16")
    e.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
    e.jpeg_init("auto")  # raises if nvJPEG cannot be had: there is no silent fallback
    yield e
    e.close()


def test_header_walk_sorts_files_into_gpu_and_host_decode(eng):
    rgb = synthetic.mri_like_images(1, 256, seed=1)[0]
    cases = {
        "baseline420": (_jpeg(rgb, quality=90), N.FILE_GPU_JPEG),
        "baseline444": (_jpeg(rgb, quality=95, subsampling=0), N.FILE_GPU_JPEG),
        "progressive": (_jpeg(rgb, quality=90, progressive=True), N.FILE_HOST_DECODE),
        "gray": (_jpeg(rgb[..., 0], quality=90), N.FILE_HOST_DECODE),  # mode "L": the reference raises on it
        "truncated": (_jpeg(rgb, quality=90)[:3000], N.FILE_HOST_DECODE),  # no end-of-image marker
        "png": (b"\x89PNG\r\n\x1a\n" + b"0" * 64, N.FILE_HOST_DECODE),
        "junk": (b"this is not an image", N.FILE_HOST_DECODE),
    }
    for name, (blob, want) in cases.items():
        info = eng.jpeg_probe(blob)
        assert info.status == want, name
        if want == N.FILE_GPU_JPEG:
            assert (info.height, info.width, info.components, info.encoding, info.precision) == (256, 256, 3, 0xC0, 8)
    assert eng.jpeg_probe(cases["baseline420"][0]).subsampling == 2 and eng.jpeg_probe(cases["baseline444"][0]).subsampling == 0


def test_read_files_only_pulls_jpegs_into_the_page_locked_buffer(eng, tmp_path):
    """A dataset directory may hold anything (the reference lists every file, src/feature_extraction.py:145-170): files that
    do not start with a JPEG SOI marker are left to the host decoder unread (length 0), missing ones are flagged."""
    rgb = synthetic.mri_like_images(1, 256, seed=1)[0]
    blob = _jpeg(rgb, quality=90)
    (tmp_path / "a.jpg").write_bytes(blob)
    (tmp_path / "b.png").write_bytes(b"\x89PNG\r\n\x1a\n" + b"0" * 4096)
    (tmp_path / "c.bin").write_bytes(b"")
    (tmp_path / "d.jpg").write_bytes(blob[:2000])  # starts like a JPEG, no end-of-image marker
    info = eng.jpeg_read_files(0, [str(tmp_path / n) for n in ("a.jpg", "b.png", "c.bin", "d.jpg", "missing.jpg")])
    assert [i.status for i in info] == [N.FILE_GPU_JPEG, N.FILE_HOST_DECODE, N.FILE_HOST_DECODE, N.FILE_HOST_DECODE, N.FILE_UNREADABLE]
    assert [i.length for i in info] == [len(blob), 0, 0, 2000, 0]
    assert (info[0].height, info[0].width) == (256, 256) and info[0].offset % 64 == 0 and info[3].offset % 64 == 0


def test_decoded_pixels_are_close_to_pillow(eng, golden_dir):
    blobs, want = [], []
    for p in sorted((golden_dir / "mri_real").rglob("*.jpg")):  # the reference's own files: 512x512, 4:2:0
        blobs.append(p.read_bytes())
        want.append(np.asarray(Image.open(p)))
    smooth = synthetic.mri_like_images(3, 512, seed=3)
    for a, kw in ((smooth[0], dict(quality=90)), (smooth[1], dict(quality=75)), (smooth[2][:301, :477], dict(quality=95, subsampling=0)),
                  (smooth[2][:334, :250], dict(quality=85, subsampling=1))):  # odd sizes: partial MCUs at the right / bottom edge
        b = _jpeg(np.ascontiguousarray(a), **kw)
        blobs.append(b)
        want.append(np.asarray(Image.open(io.BytesIO(b))))
    got = eng.jpeg_decode(blobs, [(w.shape[0], w.shape[1]) for w in want])
    worst, mean = 0, []
    for g, w in zip(got, want):
        d = np.abs(g.cpu().numpy().astype(np.int16) - w.astype(np.int16))
        worst = max(worst, int(d.max()))
        mean.append(float(d.mean()))
    print(f"nvJPEG vs Pillow over {len(want)} files: max |d| {worst}, mean |d| {np.mean(mean):.4f} (worst file {max(mean):.4f})")
    assert worst <= 6 and max(mean) <= 0.05


def _dataset(tmp_path, golden_dir):
    import shutil

    root = tmp_path / "data"
    shutil.copytree(golden_dir / "mri_real", root)
    imgs = synthetic.ragged_images([(224, 224), (300, 500)], seed=31)
    Image.fromarray(imgs[0]).save(root / "sans_label" / "p0.png")  # host decode, bit-exact
    Image.fromarray(imgs[1]).save(root / "sans_label" / "p1_progressive.jpg", quality=92, progressive=True)  # host decode
    good = (root / "sans_label" / sorted(p.name for p in (root / "sans_label").glob("*.jpg"))[0]).read_bytes()
    (root / "sans_label" / "zz_truncated.jpg").write_bytes(good[: len(good) // 3])  # failure, as the reference reports it
    (root / "sans_label" / "zz_broken.png").write_bytes(b"this is not an image")
    return root


def test_extract_embeddings_with_gpu_decode(tmp_path, golden_dir, monkeypatch):
    monkeypatch.setenv(fx.WEIGHTS_ENV, "random-bn:1234")
    root = _dataset(tmp_path, golden_dir)
    records = fx.discover_image_records(root)
    monkeypatch.setenv(fx.DECODE_MODE_ENV, "thread")
    host = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=7)
    monkeypatch.setenv(fx.DECODE_MODE_ENV, "nvjpeg")
    gpu = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=7)
    # bookkeeping and failure semantics are the reference's (src/feature_extraction.py:281-284,295)
    assert [r.relative_path for r in gpu.records] == [r.relative_path for r in host.records]
    assert sorted(p.name for p in gpu.failures) == sorted(p.name for p in host.failures) == ["zz_broken.png", "zz_truncated.jpg"]
    names = [r.relative_path.name for r in gpu.records]
    for name in ("p0.png", "p1_progressive.jpg"):  # not nvJPEG's: Pillow decoded them, rows identical to the host path
        i = names.index(name)
        assert np.array_equal(gpu.embeddings[i], host.embeddings[i]), name
    # the real MRI files against the REFERENCE's rows (minted by tests/golden/make_golden_mri.py)
    g = np.load(golden_dir / "mri_real_golden.npz")
    want = {p: row for p, row in zip(g["paths"].tolist(), g["emb_randbn"])}
    rel, cos = [], []
    for r, row in zip(gpu.records, gpu.embeddings):
        w = want.get(str(r.relative_path))
        if w is not None:
            rel.append(np.linalg.norm(row - w) / np.linalg.norm(w))
            cos.append(float(row @ w / (np.linalg.norm(row) * np.linalg.norm(w))))
    assert len(rel) == 16
    print(f"GPU-decoded MRI rows vs the reference: max relL2 {max(rel):.3e}, min cos {min(cos):.6f}")
    assert max(rel) <= 1e-2 and min(cos) >= 0.999
    # a true grayscale JPEG still raises like the reference (Normalize cannot broadcast one channel)
    Image.fromarray(synthetic.mri_like_images(1, 128, seed=2)[0][..., 0]).save(root / "sans_label" / "gray.jpg")
    with pytest.raises(RuntimeError, match="broadcast shape"):
        fx.extract_embeddings(fx.discover_image_records(root), torch.device("cuda:0"), batch_size=7)


def test_gpu_decode_writes_rows_straight_into_a_device_sink(tmp_path, golden_dir, monkeypatch):
    """The multi-GPU drop-in points the trunk's output at the rank's slot of the gather buffer (fx_embed_files_async with
    emb_dev): same rows as through the host, for the nvJPEG and for the host-decode staging alike."""
    monkeypatch.setenv(fx.WEIGHTS_ENV, "random-bn:1234")
    root = _dataset(tmp_path, golden_dir)
    records = fx.discover_image_records(root)
    eng = fx.get_engine(torch.device("cuda:0"), min_batch=8)
    for mode in ("nvjpeg", "thread"):
        monkeypatch.setenv(fx.DECODE_MODE_ENV, mode)
        host, kept_h, fail_h, _ = fx._extract_local(records, eng, 6)
        sink = torch.full((len(records), 512), float("nan"), dtype=torch.float32, device="cuda")
        none, kept_d, fail_d, _ = fx._extract_local(records, eng, 6, sink=sink)
        assert none is None and kept_d == kept_h and [p.name for p in fail_d] == [p.name for p in fail_h]
        assert torch.equal(sink[: len(kept_d)].cpu(), host), mode
        assert bool(torch.isnan(sink[len(kept_d):]).all())  # nothing written past the kept rows
