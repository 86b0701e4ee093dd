"""Drop-in for ``src/standardize_features.py``: StandardScaler over the embedding matrix and
the ``standardized_features.npz`` bundle, with the fit and the transform done on the GPU.

Same CLI (``--embeddings-npy --embeddings-csv --output-npz --log-level``, src/standardize_features.py:64-100), same
checks and exceptions (:15-39), same bundle keys ``features, paths, is_labeled, labels, scaler_mean, scaler_scale``
(:49-58).  The scaler: column mean / variance with fp64 accumulation (fx_column_stats), near-constant columns get
scale 1 by scikit-learn's own rule (sklearn/preprocessing/_data.py _is_constant_feature / _handle_zeros_in_scale),
transform = scikit-learn's float32 arithmetic, bit for bit (fx_standardize).  No CPU fallback.
"""
from __future__ import annotations

import argparse
import logging
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import torch

from .feature_extraction import _NO_WEIGHTS, get_engine


def fit_transform_device(matrix: torch.Tensor) -> Tuple[torch.Tensor, np.ndarray, np.ndarray]:
    """StandardScaler().fit_transform on a CUDA fp32 matrix -> (Z on the device, mean_ fp64, scale_ fp64)."""
    eng = get_engine(matrix.device, min_batch=1, state_dict=_NO_WEIGHTS)
    with torch.cuda.device(matrix.device):
        _, mean, std, var = eng.column_stats(matrix)
        n = matrix.shape[0]
        eps = torch.finfo(torch.float64).eps
        constant = var <= n * eps * var + (n * mean * eps) ** 2  # _is_constant_feature
        scale = torch.where(constant, torch.ones_like(std), std)  # _handle_zeros_in_scale(constant_mask=...)
        z = eng.standardize(matrix, mean, scale)
    return z, mean.cpu().numpy(), scale.cpu().numpy()


def standardize_embeddings(embeddings_path: Path, csv_path: Path, output_path: Path, device: Optional[torch.device] = None) -> None:
    if not embeddings_path.exists():
        raise FileNotFoundError(f"Embeddings file not found: {embeddings_path}")
    if not csv_path.exists():
        raise FileNotFoundError(f"Embeddings CSV not found: {csv_path}")
    logging.info("Loading embeddings from %s", embeddings_path)
    E = np.load(embeddings_path)
    if E.ndim != 2:
        raise ValueError(f"Embeddings must be 2D [N, D], got shape {E.shape}")
    logging.info("Loading metadata from %s", csv_path)
    df = pd.read_csv(csv_path)
    required_cols = {"index", "path", "bucket", "label"}
    missing = required_cols - set(df.columns)
    if missing:
        raise KeyError(f"Embeddings CSV missing columns: {', '.join(sorted(missing))}")
    df = df.sort_values("index").reset_index(drop=True)
    if len(df) != E.shape[0]:
        raise ValueError(f"Row count mismatch between CSV ({len(df)}) and embeddings ({E.shape[0]})")
    logging.info("Fitting StandardScaler and transforming features")
    device = torch.device(device) if device is not None else torch.device("cuda")
    eng = get_engine(device, min_batch=1, state_dict=_NO_WEIGHTS)
    with torch.cuda.device(eng.device):
        Z, mean, scale = fit_transform_device(torch.from_numpy(np.ascontiguousarray(E, dtype=np.float32)).to(eng.device))
        Z = Z.cpu().numpy()
    paths = df["path"].astype(str).to_numpy()
    is_labeled = (df["bucket"].astype(str) == "labeled").to_numpy()
    labels = df["label"].fillna("").astype(str)
    labels = labels.where(is_labeled, "").to_numpy()
    output_path.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(output_path, features=Z.astype(np.float32), paths=paths, is_labeled=is_labeled, labels=labels,
                        scaler_mean=np.asarray(mean, dtype=np.float32), scaler_scale=np.asarray(scale, dtype=np.float32))
    logging.info("Wrote standardized bundle: %s (N=%d, D=%d)", output_path, Z.shape[0], Z.shape[1])


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description=(
        "Standardize embeddings and build feature bundle for clustering. Consumes outputs/features/embeddings.{npy,csv} "
        "and writes outputs/features/standardized_features.npz by default."))
    parser.add_argument("--embeddings-npy", type=Path, default=Path("outputs/features/embeddings.npy"), help="Path to embeddings .npy file")
    parser.add_argument("--embeddings-csv", type=Path, default=Path("outputs/features/embeddings.csv"),
                        help="Path to embeddings CSV file (paths + labels)")
    parser.add_argument("--output-npz", type=Path, default=Path("outputs/features/standardized_features.npz"),
                        help="Path to write the standardized feature bundle")
    parser.add_argument("--log-level", type=str, default="INFO", choices=["DEBUG", "INFO", "WARNING", "ERROR"], help="Logging level")
    return parser.parse_args(argv)


def main(argv: Optional[Sequence[str]] = None) -> None:
    args = parse_args(argv)
    logging.basicConfig(level=getattr(logging, args.log_level.upper()), format="%(asctime)s [%(levelname)s] %(message)s")
    standardize_embeddings(args.embeddings_npy, args.embeddings_csv, args.output_npz)


if __name__ == "__main__":
    main()
