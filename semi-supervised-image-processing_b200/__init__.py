"""B200-native drop-in for the ``src.feature_extraction`` pass of
Septimus4/semi-supervised-image-processing.

The directory name is fixed by the repository layout and is not an importable identifier; import
the package as ``ssip_b200`` (a path alias at the repository root):

    python -m ssip_b200.feature_extraction --data-dir ... --device cuda --batch-size 256
"""
__version__ = "0.1.0"
