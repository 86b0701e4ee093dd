"""Multi-GPU drop-in: byte-identical artifacts for 1 and N visible GPUs, through the CLI's own per-GPU
workers and through torchrun.  Needs >= 2 GPUs on the box (the driver's `-m gpu` run has one: skipped there; run by
`gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`, logs under profiles/)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


def test_one_and_n_gpus_give_byte_identical_artifacts():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible")
    import check_multigpu_dropin

    check_multigpu_dropin.run(min(n, 8))
