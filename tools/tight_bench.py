"""Tight-tolerance mode: per-layer times and images/s at batch 256 (and the accuracy against the oracle port), for the
split-fp16 tensor-core kernel (default) and the CUDA-core kernel (FX_TIGHT_SIMT=1).   python tools/tight_bench.py"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
WORKER = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
from oracle import reference_path as rp
from ssip_b200 import synthetic
from ssip_b200.engine import Engine, uniform_descs
B = 256
eng = Engine(0, max_batch=B, precision="fp32")
eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
x = synthetic.noise_images(B, 224, 224, seed=3)
dev = torch.from_numpy(x.reshape(-1)).cuda()
descs = uniform_descs(B, 224, 224)
out = torch.empty((B, 512), device="cuda")
for _ in range(3):
    eng.embed_device(dev, descs, B, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    eng.embed_device(dev, descs, B, out=out)
b.record(); b.synchronize()
ms = a.elapsed_time(b) / 10
eng.profile(True)
acc = np.zeros(21)
for _ in range(5):
    eng.preprocess(dev, descs, B); eng.forward(B, out); torch.cuda.synchronize(); acc += eng.profile_read()
eng.profile(False)
want = rp.port_embed_arrays(list(x[:32]), randomize_bn=True)
got = out[:32].cpu().numpy()
rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
print(f"batch {B}: {ms:.3f} ms/step = {B / ms * 1e3:,.0f} images/s; max relL2 vs the CPU fp32 reference path {rel.max():.3e}")
print("layer ms:", " ".join(f"{i}:{v / 5:.3f}" for i, v in enumerate(acc)))
eng.close()
"""
import tempfile
if "--direct" in sys.argv:  # one interpreter, current FX_TIGHT_SIMT setting (what ncu is pointed at)
    sys.argv = [sys.argv[0], str(ROOT)]
    exec(compile(WORKER, "worker", "exec"))
    sys.exit(0)
tmp = Path(tempfile.mkdtemp()) / "w.py"
tmp.write_text(WORKER)
for knob in ("0", "1"):
    r = subprocess.run([sys.executable, str(tmp), str(ROOT)], env=dict(os.environ, FX_TIGHT_SIMT=knob), capture_output=True, text=True, timeout=600)
    print(f"## FX_TIGHT_SIMT={knob} ({'CUDA-core fp32 kernel' if knob == '1' else 'split-fp16 tensor-core kernel'})")
    print(r.stdout.strip()[-1500:], r.stderr.strip()[-800:] if r.returncode else "", flush=True)
