"""Tolerance and throughput study of the GPU JPEG decode: nvJPEG against Pillow / libjpeg-turbo, the
decoder the reference uses (src/feature_extraction.py:238).

    python tools/nvjpeg_study.py [--data DIR] [--out profiles/r02_nvjpeg_tolerance.md] [--limit N]

DIR defaults to tests/golden/_mri_local (a git-ignored copy of the reference's 1506 MRI JPEGs made in the build container)
and falls back to the 16 committed files of tests/golden/mri_real.  For every backend nvJPEG offers on this box:
per-pixel |difference| histogram over all files, then the embedding of every file from both decodes through the same
engine (bf16 and fp32) -> relative L2 between them, i.e. what the decoder swap costs against BASELINE.json's 1e-2 budget.
Then the drop-in's end-to-end rate on the same files: host process pool vs nvJPEG."""
import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import reference_path as rp  # noqa: E402  (weights only)
from ssip_b200 import _native as N  # noqa: E402
from ssip_b200 import feature_extraction as fx  # noqa: E402
from ssip_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--limit", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=3, help="copies of the dataset for the throughput run")
    args = ap.parse_args()
    data = Path(args.data) if args.data else ROOT / "tests/golden/_mri_local"
    if not data.exists():
        data = ROOT / "tests/golden/mri_real"
    files = sorted(p for p in data.rglob("*.jpg"))
    if args.limit:
        files = files[: args.limit]
    lines = [f"# nvJPEG vs Pillow (libjpeg-turbo) on {len(files)} JPEG files of `{data.relative_to(ROOT)}`", ""]

    def say(x=""):
        print(x, flush=True)
        lines.append(x)

    blobs = [p.read_bytes() for p in files]
    ref = [np.asarray(Image.open(p)) for p in files]
    sizes = [(a.shape[0], a.shape[1]) for a in ref]
    say(f"sizes: {sorted(set(sizes))[:5]}; modes all RGB: {all(a.ndim == 3 and a.shape[2] == 3 for a in ref)}")
    eng = Engine(0, max_batch=256, precision="bf16")
    eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
    eng32 = Engine(0, max_batch=64, precision="fp32")
    eng32.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
    emb_ref = eng.embed_images(ref)
    emb_ref32 = eng32.embed_images(ref[:256])
    for backend in ("hardware", "gpu"):
        try:
            got_backend = eng.jpeg_init(backend)
        except N.FxError as exc:
            say(f"\n## backend {backend}: unavailable ({exc.text})")
            continue
        say(f"\n## backend {got_backend}")
        info = [eng.jpeg_probe(b) for b in blobs]
        take = [i for i, fi in enumerate(info) if fi.status == N.FILE_GPU_JPEG]
        say(f"eligible for GPU decode: {len(take)} of {len(files)} (subsampling codes: {sorted(set(fi.subsampling for fi in info))})")
        if not take:
            continue
        hist = np.zeros(256, np.int64)
        worst, dec = [], []
        t0 = time.perf_counter()
        for s in range(0, len(take), 256):
            part = take[s : s + 256]
            out = eng.jpeg_decode([blobs[i] for i in part], [sizes[i] for i in part])
            dec.extend(o.cpu().numpy() for o in out)
        dt = time.perf_counter() - t0
        for i, d in zip(take, dec):
            diff = np.abs(d.astype(np.int16) - ref[i].astype(np.int16))
            hist += np.bincount(diff.reshape(-1), minlength=256)
            worst.append(int(diff.max()))
        tot = hist.sum()
        say(f"synchronous fx_jpeg_decode incl. D2H of the pixels: {len(take) / dt:.0f} images/s")
        say(f"pixels compared: {tot}; identical {hist[0] / tot:.4%}, |d|=1 {hist[1] / tot:.4%}, |d|=2 {hist[2] / tot:.4%}, |d|>=3 {hist[3:].sum() / tot:.4%}; "
            f"mean |d| {(hist * np.arange(256)).sum() / tot:.4f}; max |d| {max(worst)}")
        emb = eng.embed_images(dec)
        rel = np.linalg.norm(emb - emb_ref[take], axis=1) / np.linalg.norm(emb_ref[take], axis=1)
        say(f"embedding relL2 nvJPEG-vs-Pillow decode, bf16 engine: median {np.median(rel):.3e}, p99 {np.percentile(rel, 99):.3e}, max {rel.max():.3e} (budget 1e-2)")
        k = [j for j, i in enumerate(take) if i < 256]
        emb32 = eng32.embed_images([dec[j] for j in k])
        r32 = np.linalg.norm(emb32 - emb_ref32[[take[j] for j in k]], axis=1) / np.linalg.norm(emb_ref32[[take[j] for j in k]], axis=1)
        say(f"embedding relL2 nvJPEG-vs-Pillow decode, fp32 engine ({len(k)} files): median {np.median(r32):.3e}, max {r32.max():.3e}")
    eng.close()
    eng32.close()

    # ---- end to end through the drop-in on files --------------------------------------------------
    import shutil
    import tempfile

    tmp = Path(tempfile.mkdtemp())
    for rep in range(args.repeat):
        for p in files:
            dst = tmp / "sans_label" / f"r{rep}_{p.name}"
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(p, dst)
    os.environ[fx.WEIGHTS_ENV] = "random-bn:1234"
    records = fx.discover_image_records(tmp)
    say(f"\n## end to end, extract_embeddings on {len(records)} files ({args.repeat} copies), batch 256, host cores {len(os.sched_getaffinity(0))}")
    results = {}
    for mode, extra in (("process", {}), ("nvjpeg", {fx.JPEG_BACKEND_ENV: "hardware"}), ("nvjpeg", {fx.JPEG_BACKEND_ENV: "gpu"})):
        os.environ[fx.DECODE_MODE_ENV] = mode
        os.environ.update(extra)
        name = mode + ("/" + extra[fx.JPEG_BACKEND_ENV] if extra else "")
        try:
            fx.extract_embeddings(records[:512], torch.device("cuda:0"), batch_size=256)  # warm-up (workers, nvJPEG buffers)
            best = 0.0
            for _ in range(2):
                t0 = time.perf_counter()
                res = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=256)
                best = max(best, len(records) / (time.perf_counter() - t0))
            results[name] = res.embeddings
            say(f"- {name}: {best:.0f} images/s, failures {len(res.failures)}")
        except N.FxError as exc:
            say(f"- {name}: unavailable ({exc.text})")
    base = results.get("process")
    for name, emb in results.items():
        if name != "process" and base is not None:
            rel = np.linalg.norm(emb - base, axis=1) / np.linalg.norm(base, axis=1)
            say(f"- {name} vs process rows: relL2 median {np.median(rel):.3e}, max {rel.max():.3e}")
    shutil.rmtree(tmp, ignore_errors=True)
    if args.out:
        Path(args.out).write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
