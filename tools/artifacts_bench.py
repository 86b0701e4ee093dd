#!/usr/bin/env python
"""SURVEY.md 8f rank 4 at scale: the reference's artifact-writing expressions against ssip_b200._artifacts on N rows.

    python tools/artifacts_bench.py [N=1000000] [files=100000]

Host-only (no GPU needed).  Writes under $TMPDIR, checks that both sides produced identical bytes, prints a markdown table."""
import hashlib
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import pandas as pd

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssip_b200 import _artifacts as A  # noqa: E402
from ssip_b200.feature_extraction import ImageRecord  # noqa: E402


def timed(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    tmp = Path(tempfile.mkdtemp(prefix="ssip_artifacts_"))
    emb = np.random.default_rng(0).random((n, 512), dtype=np.float32)
    recs = [ImageRecord(tmp / "data" / f"{i:07d}.jpg", Path("avec_labels") / ("cancer" if i % 2 else "normal") / f"img_{i:07d}.jpg" if i % 5
                        else Path("sans_label") / f"img_{i:07d}.jpg", "labeled" if i % 5 else "unlabeled", ("cancer" if i % 2 else "normal") if i % 5 else None)
            for i in range(n)]
    rows = []
    # --- embeddings.npy: src/feature_extraction.py:416 ---
    t_ref, _ = timed(lambda: np.save(tmp / "ref.npy", emb.astype(np.float32)))
    t_our, _ = timed(lambda: A.write_npy(tmp / "our.npy", emb))
    same = hashlib.sha256((tmp / "ref.npy").read_bytes()).digest() == hashlib.sha256((tmp / "our.npy").read_bytes()).digest()
    rows.append((f"embeddings.npy, [{n},512] fp32 ({emb.nbytes / 1e9:.2f} GB)", t_ref, t_our, same))
    (tmp / "ref.npy").unlink(), (tmp / "our.npy").unlink()

    # --- embeddings.csv: src/feature_extraction.py:418-431 ---
    def ref_csv():
        r = [{"index": i, "path": str(x.relative_path), "bucket": x.bucket, "label": x.label} for i, x in enumerate(recs)]
        pd.DataFrame(r).to_csv(tmp / "ref.csv", index=False)

    t_ref, _ = timed(ref_csv)
    t_our, _ = timed(lambda: A.write_embeddings_csv(tmp / "our.csv", recs))
    rows.append((f"embeddings.csv, {n} rows", t_ref, t_our, (tmp / "ref.csv").read_bytes() == (tmp / "our.csv").read_bytes()))

    # --- dataset digest over real files: src/feature_extraction.py:316-331 ---
    (tmp / "data").mkdir()
    sub = recs[:n_files]
    for r in sub:
        r.absolute_path.write_bytes(b"")

    def ref_digest():
        h = hashlib.sha256()
        for r in sorted(sub, key=lambda r: str(r.relative_path)):
            st = r.absolute_path.stat()
            h.update(str(r.relative_path).encode("utf-8"))
            h.update(str(st.st_size).encode("utf-8"))
            h.update(str(int(st.st_mtime)).encode("utf-8"))
        return h.hexdigest()

    t_ref, d_ref = timed(ref_digest)
    t_our, d_our = timed(lambda: A.dataset_digest(sub))
    rows.append((f"dataset digest, {n_files} files (warm dentry cache)", t_ref, t_our, d_ref == d_our))

    print(f"host: {os.cpu_count()} cores; N = {n}\n")
    print("| artifact | reference expression (s) | ssip_b200._artifacts (s) | identical bytes |")
    print("|---|---|---|---|")
    for name, a, b, same in rows:
        print(f"| {name} | {a:.2f} | {b:.2f} | {same} |")
    import shutil

    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
