"""Drop-in for ``src/standardize_features.py``: StandardScaler over the embedding matrix and
the ``standardized_features.npz`` bundle, with the fit and the transform done on the GPU.

Contract kept (reference lines): CLI flags and defaults (src/standardize_features.py:64-100), the input checks and the
exception types they raise (:15-39), the bundle keys ``features, paths, is_labeled, labels, scaler_mean, scaler_scale`` and
their dtypes (:49-58).  What is new underneath: the scaler is two HBM-bound passes on the device -- column mean /
variance with fp64 accumulation (``fx_column_stats``), near-constant columns get scale 1 by scikit-learn's own rule
(sklearn/preprocessing/_data.py ``_is_constant_feature`` / ``_handle_zeros_in_scale``), the transform is scikit-learn's
float32 arithmetic bit for bit (``fx_standardize``).  No CPU fallback.
"""
from __future__ import annotations

import argparse
import logging
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import torch

from .feature_extraction import _NO_WEIGHTS, get_engine

CSV_COLUMNS = ("index", "path", "bucket", "label")  # what save_artifacts writes (src/feature_extraction.py:418-431)
DEFAULTS = {
    "embeddings_npy": Path("outputs/features/embeddings.npy"),
    "embeddings_csv": Path("outputs/features/embeddings.csv"),
    "output_npz": Path("outputs/features/standardized_features.npz"),
}


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


@dataclass
class EmbeddingTable:
    """The extraction's two artifacts, row-aligned: matrix row i <-> CSV row with index i."""

    matrix: np.ndarray  # [N, D]
    paths: np.ndarray  # str
    is_labeled: np.ndarray  # bool
    labels: np.ndarray  # str, "" for unlabeled rows

    @classmethod
    def load(cls, npy: Path, csv: Path) -> "EmbeddingTable":
        for what, path in (("Embeddings file", npy), ("Embeddings CSV", csv)):
            if not path.exists():
                raise FileNotFoundError(f"{what} not found: {path}")
        logging.info("Loading embeddings from %s", npy)
        matrix = np.load(npy)
        if matrix.ndim != 2:
            raise ValueError(f"Embeddings must be 2D [N, D], got shape {matrix.shape}")
        logging.info("Loading metadata from %s", csv)
        frame = pd.read_csv(csv)
        absent = sorted(set(CSV_COLUMNS) - set(frame.columns))
        if absent:
            raise KeyError(f"Embeddings CSV missing columns: {', '.join(absent)}")
        frame = frame.sort_values("index", kind="quicksort").reset_index(drop=True)  # rows follow the explicit index column
        if len(frame) != matrix.shape[0]:
            raise ValueError(f"Row count mismatch between CSV ({len(frame)}) and embeddings ({matrix.shape[0]})")
        labeled = frame["bucket"].astype(str).to_numpy() == "labeled"
        names = frame["label"].fillna("").astype(str).to_numpy()
        return cls(matrix, frame["path"].astype(str).to_numpy(), labeled, np.where(labeled, names, "").astype(object))

    def write_bundle(self, target: Path, z: np.ndarray, mean: np.ndarray, scale: np.ndarray) -> None:
        target.parent.mkdir(parents=True, exist_ok=True)
        arrays = {
            "features": z.astype(np.float32),
            "paths": self.paths,
            "is_labeled": self.is_labeled,
            "labels": self.labels,
            "scaler_mean": np.asarray(mean, dtype=np.float32),
            "scaler_scale": np.asarray(scale, dtype=np.float32),
        }
        np.savez_compressed(target, **arrays)
        logging.info("Wrote standardized bundle: %s (N=%d, D=%d)", target, z.shape[0], z.shape[1])


def standardize_embeddings(embeddings_path: Path, csv_path: Path, output_path: Path, device: Optional[torch.device] = None) -> None:
    """Same call as the reference's (plus an optional device): load, check, scale on the GPU, write the bundle."""
    table = EmbeddingTable.load(Path(embeddings_path), Path(csv_path))
    logging.info("Fitting StandardScaler and transforming features")
    eng = get_engine(torch.device(device) if device is not None else torch.device("cuda"), min_batch=1, state_dict=_NO_WEIGHTS)
    with torch.cuda.device(eng.device):
        on_device = torch.from_numpy(np.ascontiguousarray(table.matrix, dtype=np.float32)).to(eng.device)
        z, mean, scale = fit_transform_device(on_device)
        z = z.cpu().numpy()
    table.write_bundle(Path(output_path), z, mean, scale)


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description=(
        "Standardize embeddings and build feature bundle for clustering. Consumes outputs/features/embeddings.{npy,csv} "
        "and writes outputs/features/standardized_features.npz by default."))
    for flag, key, text in (("--embeddings-npy", "embeddings_npy", "Path to embeddings .npy file"),
                            ("--embeddings-csv", "embeddings_csv", "Path to embeddings CSV file (paths + labels)"),
                            ("--output-npz", "output_npz", "Path to write the standardized feature bundle")):
        parser.add_argument(flag, type=Path, default=DEFAULTS[key], help=text)
    parser.add_argument("--log-level", type=str, default="INFO", choices=["DEBUG", "INFO", "WARNING", "ERROR"], help="Logging level")
    return parser.parse_args(argv)


def main(argv: Optional[Sequence[str]] = None) -> None:
    args = parse_args(argv)
    logging.basicConfig(level=getattr(logging, args.log_level.upper()), format="%(asctime)s [%(levelname)s] %(message)s")
    standardize_embeddings(args.embeddings_npy, args.embeddings_csv, args.output_npz)


if __name__ == "__main__":
    main()
