"""Importable alias of the ``semi-supervised-image-processing_b200/`` package directory.

The directory name required by the repo layout contains hyphens and so cannot be written in an
``import`` statement; this shim points ``ssip_b200``'s submodule search path at it, so
``python -m ssip_b200.feature_extraction`` is the drop-in for ``python -m src.feature_extraction``.
"""
from pathlib import Path as _Path

_PKG_DIR = _Path(__file__).resolve().parent.parent / "semi-supervised-image-processing_b200"
__path__.insert(0, str(_PKG_DIR))
__version__ = "0.1.0"
