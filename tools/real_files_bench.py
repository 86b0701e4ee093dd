"""End-to-end on FILES: where the time goes when the input is a directory of JPEGs, as in the
reference's dataset (512x512 gray-in-RGB JPEGs, ~24 KB each), instead of decoded arrays in host memory.

    SSIP_B200_WEIGHTS=random:1234 python tools/real_files_bench.py [n_files]

Writes n synthetic MRI-like JPEGs, then measures (1) Pillow decode alone on the drop-in's thread pool, (2) the
drop-in's extract_embeddings (decode pool -> pinned staging -> fx_embed_host_async slots), (3) the same images
already decoded (GPU path only).  Prints a markdown table.
"""
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("SSIP_B200_WEIGHTS", "random:1234")
import numpy as np
import torch
from PIL import Image

from ssip_b200 import feature_extraction as fx
from ssip_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
root = Path(tempfile.mkdtemp()) / "data"
(root / "avec_labels" / "cancer").mkdir(parents=True)
(root / "avec_labels" / "normal").mkdir(parents=True)
(root / "sans_label").mkdir(parents=True)
base = synthetic.mri_like_images(64, 512, seed=1)
t0 = time.perf_counter()
for i in range(n):
    sub = "avec_labels/cancer" if i < 50 else ("avec_labels/normal" if i < 100 else "sans_label")
    Image.fromarray(np.roll(base[i % 64], i // 64, axis=1)).save(root / sub / f"img_{i:05d}.jpg", quality=90)
size = sum(p.stat().st_size for p in root.rglob("*.jpg")) / n
records = fx.discover_image_records(root)
threads = min(32, os.cpu_count() or 8)

t0 = time.perf_counter()
with ThreadPoolExecutor(max_workers=threads) as pool:
    arrays = list(pool.map(fx._load_file, [r.absolute_path for r in records]))
t_decode = time.perf_counter() - t0

dev = torch.device("cuda:0")
t_files = {}
for mode in ("thread", "process"):
    os.environ["SSIP_B200_DECODE"] = mode
    fx.extract_embeddings(records[:512], dev, batch_size=256)  # warm-up: engine, weights, worker processes, buffers
    t0 = time.perf_counter()
    res = fx.extract_embeddings(records, dev, batch_size=256)
    t_files[mode] = time.perf_counter() - t0

eng = fx.get_engine(dev, min_batch=256)
t0 = time.perf_counter()
for rep in range(3):
    emb = eng.embed_images(arrays)
t_arrays = (time.perf_counter() - t0) / 3
assert np.array_equal(emb, res.embeddings)

print(f"## Files on disk: {n} synthetic 512x512 MRI-like JPEGs (avg {size/1024:.1f} KB), batch 256, 1xB200, {threads} decode threads of {os.cpu_count()} cores\n")
print("| stage | seconds | images/s |")
print("|---|---|---|")
print(f"| Pillow decode alone (thread pool) | {t_decode:.2f} | {n / t_decode:,.0f} |")
print(f"| drop-in extract_embeddings on the files, decode on {threads} THREADS (+ pack + H2D + kernels + D2H) | {t_files['thread']:.2f} | {n / t_files['thread']:,.0f} |")
print(f"| drop-in extract_embeddings on the files, decode in {threads} worker PROCESSES straight into shared pinned memory (default from 512 files) | {t_files['process']:.2f} | {n / t_files['process']:,.0f} |")
print(f"| same images already decoded (pack + H2D + kernels + D2H, synchronous embed_images) | {t_arrays:.2f} | {n / t_arrays:,.0f} |")
