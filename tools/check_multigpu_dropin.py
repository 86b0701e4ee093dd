"""Multi-GPU check of the drop-in CLI: the same dataset through `python -m ssip_b200.feature_extraction --device cuda`
with ONE visible GPU and with N visible GPUs (the CLI starts one worker per visible GPU) must give
byte-identical embeddings.npy / embeddings.csv, and the same again under torchrun.

    python tools/check_multigpu_dropin.py N [log file]

The dataset holds ragged sizes, a JPEG, an undecodable file in the LAST shard and one in the FIRST (uneven shards ->
the compaction path of the in-place all-gather).  Also run by tests/test_gpu_multi.py when the box has >= 2 GPUs.
"""
import os
import socket
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ssip_b200 import synthetic  # noqa: E402


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run(n: int, log=None) -> None:
    from PIL import Image

    def say(*a):
        print(*a, flush=True)
        if log:
            with open(log, "a") as fh:
                print(*a, file=fh)

    tmp = Path(tempfile.mkdtemp())
    imgs = list(synthetic.noise_images(61, 224, 224, seed=2)) + synthetic.ragged_images([(300, 500), (512, 512)], seed=3)
    synthetic.write_png_dataset(tmp / "data", imgs, n_labeled=8)
    (tmp / "data" / "sans_label" / "zz_broken.png").write_bytes(b"nope")
    (tmp / "data" / "avec_labels" / "cancer" / "aa_broken.png").write_bytes(b"nope either")
    Image.fromarray(synthetic.mri_like_images(1, 512, seed=4)[0]).save(tmp / "data" / "sans_label" / "mri.jpg", quality=90)
    base_env = dict(os.environ, SSIP_B200_WEIGHTS="random-bn:1234", PYTHONPATH=str(ROOT))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_PORT", "MASTER_ADDR"):
        base_env.pop(k, None)
    cli = ["-m", "ssip_b200.feature_extraction", "--data-dir", str(tmp / "data"), "--device", "cuda", "--batch-size", "16"]
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    ids = visible.split(",") if visible else [str(i) for i in range(n)]
    assert len(ids) >= n, f"need {n} GPUs, CUDA_VISIBLE_DEVICES={visible}"
    runs = [("1 gpu", dict(base_env, CUDA_VISIBLE_DEVICES=ids[0]), [sys.executable] + cli),
            (f"{n} gpus, workers started by the CLI", dict(base_env, CUDA_VISIBLE_DEVICES=",".join(ids[:n])), [sys.executable] + cli),
            (f"{n} gpus, torchrun", dict(base_env, CUDA_VISIBLE_DEVICES=",".join(ids[:n])),
             [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
              "--master-port", str(free_port())] + cli)]
    outs = []
    for k, (name, env, cmd) in enumerate(runs):
        wd = tmp / f"run{k}"
        wd.mkdir()
        r = subprocess.run(cmd, cwd=wd, env=env, capture_output=True, text=True, timeout=900)
        say(f"{name}: rc={r.returncode}", r.stderr[-1500:] if r.returncode else "")
        assert r.returncode == 0, name
        outs.append((np.load(wd / "outputs/features/embeddings.npy"), (wd / "outputs/features/embeddings.csv").read_text(),
                     (wd / "outputs/logs/feature_extraction.log").read_text()))
    assert outs[0][0].shape == (64, 512), outs[0][0].shape
    for k in (1, 2):
        assert np.array_equal(outs[0][0], outs[k][0]), f"{runs[k][0]}: embeddings differ from the single-GPU run"
        assert outs[0][1] == outs[k][1], f"{runs[k][0]}: CSV differs"
        assert outs[k][2].count("Failed to decode") >= 1  # rank 0's own failure is in its log
    say(f"OK: 1 GPU, {n} GPUs (CLI workers) and {n} GPUs (torchrun) give byte-identical [64,512] embeddings and CSV")


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 2, sys.argv[2] if len(sys.argv) > 2 else None)
