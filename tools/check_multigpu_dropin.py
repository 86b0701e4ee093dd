"""Manual multi-GPU check of the drop-in: run the CLI with 1 rank and with N ranks (torchrun) on the same
synthetic PNG dataset and require byte-identical embeddings.npy / embeddings.csv.
    python tools/check_multigpu_dropin.py N
"""
import os, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ssip_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tmp = Path(tempfile.mkdtemp())
imgs = list(synthetic.noise_images(61, 224, 224, seed=2)) + synthetic.ragged_images([(300, 500), (512, 512)], seed=3)
synthetic.write_png_dataset(tmp / "data", imgs, n_labeled=8)
(tmp / "data" / "sans_label" / "broken.png").write_bytes(b"nope")
env = dict(os.environ, SSIP_B200_WEIGHTS="random-bn:1234", PYTHONPATH=str(ROOT))
outs = []
for world in (1, n):
    wd = tmp / f"run{world}"
    wd.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", "-m", "ssip_b200.feature_extraction", "--data-dir", str(tmp / "data"), "--device", "cuda", "--batch-size", "16"]
    r = subprocess.run(cmd, cwd=wd, env=env, capture_output=True, text=True, timeout=600)
    print(f"world {world}: rc={r.returncode}", r.stderr[-400:] if r.returncode else "")
    assert r.returncode == 0
    outs.append((np.load(wd / "outputs/features/embeddings.npy"), (wd / "outputs/features/embeddings.csv").read_text()))
assert outs[0][0].shape == (63, 512), outs[0][0].shape
assert np.array_equal(outs[0][0], outs[1][0]), "embeddings differ between world sizes"
assert outs[0][1] == outs[1][1]
print(f"OK: world 1 and world {n} give byte-identical [63,512] embeddings and CSV")
