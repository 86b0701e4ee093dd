"""FX_TC4=1 (the 4-CTA-cluster kernel with weight K-blocks multicast to two CTA pairs, csrc/conv_tc.cu) against the default
CTA-pair kernel: the same 300 images at batch 40 and batch 256 in two interpreters, rows must be byte-identical; prints
the per-launch times of the six layer3 / layer4 3x3 convs for both.    python tools/tc4_check.py"""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
WORKER = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
from oracle import reference_path as rp
from ssip_b200 import synthetic
from ssip_b200.engine import Engine, uniform_descs
eng = Engine(0, max_batch=256, precision="bf16")
eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
x = synthetic.noise_images(512, 224, 224, seed=21)
dev = torch.from_numpy(x.reshape(-1)).cuda()
per = 224 * 224 * 3
outs = []
for batch in (40, 256):
    o = torch.empty((512, 512), device="cuda")
    for s in range(0, 512 - batch + 1, batch):
        eng.embed_device(dev[s * per:(s + batch) * per], uniform_descs(batch, 224, 224), batch, out=o[s:s + batch])
    torch.cuda.synchronize()
    outs.append(o[: (512 // batch) * batch].cpu().numpy())
eng.profile(True)
ms = np.zeros(21)
for _ in range(10):
    eng.preprocess(dev[:256 * per], uniform_descs(256, 224, 224), 256)
    eng.forward(256, o[:256])
    torch.cuda.synchronize()
    ms += eng.profile_read()
print("layer ms (batch 256):", " ".join(f"{i}:{v / 10:.4f}" for i, v in enumerate(ms) if i in (11, 13, 14, 16, 18, 19)), flush=True)
np.savez(sys.argv[2], a=outs[0], b=outs[1])
eng.close()
"""

tmp = Path(tempfile.mkdtemp())
(tmp / "w.py").write_text(WORKER)
res = []
for knob in ("0", "1"):
    env = dict(os.environ, FX_TC4=knob)
    r = subprocess.run([sys.executable, str(tmp / "w.py"), str(ROOT), str(tmp / f"o{knob}.npz")], env=env, capture_output=True, text=True, timeout=300)
    print(f"FX_TC4={knob}: rc={r.returncode}", r.stdout.strip()[-400:], r.stderr.strip()[-600:] if r.returncode else "")
    if r.returncode:
        sys.exit(1)
    res.append(np.load(tmp / f"o{knob}.npz"))
for k in ("a", "b"):
    same = np.array_equal(res[0][k], res[1][k])
    print(f"rows {k}: identical={same} maxdiff={np.abs(res[0][k] - res[1][k]).max():.3e} finite={np.isfinite(res[1][k]).all()}")
    assert same
print("TC4 OK")
