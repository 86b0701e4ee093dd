"""Ragged datasets: every image its own (height, width) -> a new coefficient table per image.  Images/s of the host-buffer
path on decoded arrays (batch 256, pipelined slots).  Set A warms everything up (allocations, kernel loading); set B (other
sizes) is then run twice: first pass = tables built and uploaded on the fly, second pass = tables cached."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import reference_path as rp
from ssip_b200 import _native as N
from ssip_b200.engine import Engine, pack_images

rng = np.random.default_rng(1)
n = 1024
shapes = set()
while len(shapes) < 2 * n:
    shapes.add((int(rng.integers(230, 700)), int(rng.integers(230, 700))))
shapes = sorted(shapes)
rng.shuffle(shapes)
sets = {name: [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in part] for name, part in (("A", shapes[:n]), ("B", shapes[n:]))}
eng = Engine(0, max_batch=256, precision="bf16")
eng.load_state_dict(rp.make_backbone(randomize_bn=True).state_dict())
outs = [torch.empty((256, 512)).pin_memory() for _ in range(4)]
for name, passes in (("A", 1), ("B", 3)):
  imgs = sets[name]
  batches = []
  for lo in range(0, n, 256):
    buf, descs, total = pack_images(imgs[lo:lo + 256])
    batches.append((torch.from_numpy(buf[:total].copy()).pin_memory(), descs, total))
  for p in range(passes):
    t0 = time.perf_counter()
    for i, (b, d, t) in enumerate(batches):
        eng.embed_host_wait(i % 4)
        eng.embed_host_async(i % 4, b, d, 256, t, outs[i % 4])
    for s in range(4):
        eng.embed_host_wait(s)
    dt = time.perf_counter() - t0
    print(f"set {name} pass {p}: {n / dt:,.0f} images/s ({n} images, {n} distinct sizes, {sum(t for _, _, t in batches) / 1e6:.0f} MB of pixels)", flush=True)
# parity spot check of the last batch against the oracle
want = rp.port_embed_arrays(imgs[-256:][:8], randomize_bn=True)
got = outs[(len(batches) - 1) % 4][:8].numpy()
print("max relL2 vs the oracle on 8 of them:", float((np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)).max()))
eng.close()
