"""Artifact writers sized for [N,512] matrices of 10^6 rows.

The reference's `save_artifacts` (src/feature_extraction.py:401-502) is written for 1.5 k images:
`np.save(path, embeddings.astype(np.float32))` copies the whole matrix before writing it, the CSV goes
through a list of dicts and a pandas DataFrame, and `compute_dataset_digest` (:316-331) runs after both.  At N = 1 M
that is a second 2 GB buffer (4.5 s) and 11 s of DataFrame construction and formatting, all after the GPUs have
gone idle (`profiles/r01_artifacts_at_scale.md`).  The writers here produce the SAME BYTES:

* `write_npy`          -- the `.npy` header numpy itself would write, then the rows in chunks: straight from a
                          C-contiguous float32 host matrix (no copy), or from the gathered matrix still on the
                          GPU through two page-locked buffers, the device->host copy of chunk k+1 running while
                          chunk k is written to the file;
* `write_embeddings_csv` -- the `csv` module with the dialect pandas' `to_csv(index=False)` uses (pandas drives
                          the same module underneath), without the DataFrame;
* `dataset_digest`     -- the reference's loop, run on a worker thread while the two files are being written.

`save_parallel` runs the three side by side.  Nothing here touches the embeddings' values.
"""
from __future__ import annotations

import csv
import hashlib
import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Callable, Sequence

import numpy as np
import torch

NPY_CHUNK_BYTES = 64 << 20   # rows per write are chosen so that a chunk is about this large
CSV_COLUMNS = ("index", "path", "bucket", "label")


def _npy_header(fh, shape, dtype=np.float32) -> None:
    """The header `np.save` writes for a C-ordered array of this shape (format 1.0, 64-byte aligned)."""
    np.lib.format.write_array_header_1_0(fh, {"descr": np.lib.format.dtype_to_descr(np.dtype(dtype)), "fortran_order": False,
                                              "shape": tuple(int(s) for s in shape)})


def write_npy(path: Path, matrix, chunk_bytes: int = NPY_CHUNK_BYTES) -> None:
    """`np.save(path, matrix.astype(np.float32))` (src/feature_extraction.py:416), streamed.

    `matrix`: numpy array (any dtype / layout; float32 C-contiguous is written without a copy) or a 2-D float32
    CUDA tensor, which is copied to the host chunk by chunk while the previous chunk is being written."""
    path = Path(path)
    if isinstance(matrix, torch.Tensor) and matrix.is_cuda:
        _write_npy_from_device(path, matrix, chunk_bytes)
        return
    if isinstance(matrix, torch.Tensor):
        matrix = matrix.numpy()
    arr = np.asarray(matrix)
    if not arr.flags.c_contiguous:  # strided / Fortran-ordered input (never produced by this package): the reference's own
        np.save(path, arr.astype(np.float32))  # expression decides the on-disk order
        return
    if arr.dtype != np.float32:
        arr = arr.astype(np.float32)
    with open(path, "wb") as fh:
        _npy_header(fh, arr.shape)
        flat = arr.reshape(-1).view(np.uint8)
        for lo in range(0, flat.size, chunk_bytes):
            fh.write(memoryview(flat[lo:lo + chunk_bytes]))


def _write_npy_from_device(path: Path, matrix: torch.Tensor, chunk_bytes: int) -> None:
    if matrix.dtype != torch.float32 or matrix.dim() != 2:
        raise ValueError("write_npy: the device matrix must be a 2-D float32 tensor")
    matrix = matrix.contiguous()
    n, d = matrix.shape
    rows = max(1, min(max(n, 1), chunk_bytes // max(1, d * 4)))
    bufs = [torch.empty((rows, d), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream = torch.cuda.Stream(device=matrix.device)
    copy_stream.wait_stream(torch.cuda.current_stream(matrix.device))  # the matrix was produced on the current stream
    starts = list(range(0, n, rows))

    def launch(k: int) -> None:
        lo = starts[k]
        hi = min(n, lo + rows)
        with torch.cuda.stream(copy_stream):
            bufs[k & 1][: hi - lo].copy_(matrix[lo:hi], non_blocking=True)
            done[k & 1].record(copy_stream)

    with open(path, "wb") as fh:
        _npy_header(fh, (n, d))
        if starts:
            launch(0)
        for k, lo in enumerate(starts):
            hi = min(n, lo + rows)
            done[k & 1].synchronize()
            if k + 1 < len(starts):
                launch(k + 1)  # into the other buffer, which the previous iteration has finished writing out
            fh.write(memoryview(bufs[k & 1][: hi - lo].numpy().reshape(-1).view(np.uint8)))


def write_embeddings_csv(path: Path, records: Sequence, relative_path: Callable = lambda r: str(r.relative_path)) -> None:
    """`DataFrame(rows).to_csv(path, index=False)` of src/feature_extraction.py:418-431, byte for byte.

    pandas writes through `csv.writer(lineterminator=os.linesep, delimiter=",", quoting=QUOTE_MINIMAL, doublequote=True,
    quotechar='"')`, UTF-8, a missing label as the empty string; so does this."""
    with open(path, "w", encoding="utf-8", newline="") as fh:
        w = csv.writer(fh, lineterminator=os.linesep, delimiter=",", quoting=csv.QUOTE_MINIMAL, doublequote=True, quotechar='"')
        w.writerow(CSV_COLUMNS)
        w.writerows((i, relative_path(r), r.bucket, "" if r.label is None else r.label) for i, r in enumerate(records))


def dataset_digest(records: Sequence) -> str:
    """sha256 over (relative path, size, int(mtime)) in path order (src/feature_extraction.py:316-331).

    One pass with `os.stat` on the path string.  (A thread pool over the `stat` calls was measured and dropped: with the
    dentry cache warm -- the files have just been decoded -- 100 k files take 0.5 s serially and 3.9 s through a pool.)"""
    h = hashlib.sha256()
    for rec in sorted(records, key=lambda r: str(r.relative_path)):
        st = os.stat(rec.absolute_path)
        h.update(str(rec.relative_path).encode("utf-8"))
        h.update(str(st.st_size).encode("utf-8"))
        h.update(str(int(st.st_mtime)).encode("utf-8"))
    return h.hexdigest()


def save_parallel(npy_path: Path, csv_path: Path, matrix, records: Sequence) -> str:
    """Writes embeddings.npy and embeddings.csv and returns the dataset digest, the three running side by side."""
    with ThreadPoolExecutor(max_workers=2) as pool:
        csv_job = pool.submit(write_embeddings_csv, csv_path, records)
        digest_job = pool.submit(dataset_digest, records)
        write_npy(npy_path, matrix)  # this thread: CUDA calls stay on the thread that owns the context
        csv_job.result()
        return digest_job.result()
