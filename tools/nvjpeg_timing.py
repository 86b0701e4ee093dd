"""Where does the nvJPEG file path spend its host time?  (extract_embeddings on copies of the MRI files, batch 256.)"""
import os, sys, time, shutil, tempfile
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ssip_b200 import feature_extraction as fx
from ssip_b200 import _native as N
from ssip_b200.engine import Engine

data = ROOT / "tests/golden/_mri_local"
if not data.exists():
    data = ROOT / "tests/golden/mri_real"
files = sorted(data.rglob("*.jpg"))
tmp = Path(tempfile.mkdtemp())
reps = max(1, 4608 // len(files))
for rep in range(reps):
    for p in files:
        d = tmp / "sans_label" / f"r{rep}_{p.name}"
        d.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(p, d)
os.environ[fx.WEIGHTS_ENV] = "random-bn:1234"
os.environ[fx.DECODE_MODE_ENV] = "nvjpeg"
records = fx.discover_image_records(tmp)
eng = fx.get_engine(torch.device("cuda:0"), min_batch=256)
eng.jpeg_init(os.environ.get("SSIP_B200_NVJPEG_BACKEND", "auto"))
paths = [str(r.absolute_path) for r in records]
# phase timings of one slot, synchronously
import ctypes
t_read = t_py = t_submit = t_wait = 0.0
out = torch.empty((256, 512)).pin_memory()
for lo in range(0, len(paths) - 255, 256):
    chunk = paths[lo:lo + 256]
    t0 = time.perf_counter()
    info = eng.jpeg_read_files(0, chunk)
    t1 = time.perf_counter()
    n = len(chunk)
    sel = (N.FileInfo * n)(); descs = (N.ImageDesc * n)(); off = 0
    for j in range(n):
        sel[j] = info[j]
        descs[j].offset, descs[j].height, descs[j].width, descs[j].channels = off, info[j].height, info[j].width, 3
        off += (info[j].height * info[j].width * 3 + 255) // 256 * 256
    t2 = time.perf_counter()
    eng.embed_files_async(0, sel, [None] * n, descs, n, off, out)
    t3 = time.perf_counter()
    eng.embed_host_wait(0)
    t4 = time.perf_counter()
    t_read += t1 - t0; t_py += t2 - t1; t_submit += t3 - t2; t_wait += t4 - t3
nb = len(paths) // 256
print(f"per 256-file batch (synchronous, one slot): read+header {t_read / nb * 1e3:.2f} ms, python bookkeeping {t_py / nb * 1e3:.2f} ms, "
      f"fx_embed_files_async (host part of nvjpegDecodeBatched + enqueue) {t_submit / nb * 1e3:.2f} ms, wait for the GPU {t_wait / nb * 1e3:.2f} ms")
for _ in range(2):
    t0 = time.perf_counter()
    res = fx.extract_embeddings(records, torch.device("cuda:0"), batch_size=256)
    print(f"extract_embeddings: {len(records) / (time.perf_counter() - t0):,.0f} images/s")
shutil.rmtree(tmp, ignore_errors=True)
