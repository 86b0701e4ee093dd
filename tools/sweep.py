"""BASELINE.json configs 4 and 5 as measurements (one GPU):
  * C5: batch-size sweep 1..1024, bf16 (tcgen05) vs fp32 tight-tolerance mode (split-fp16 tcgen05 convs with
    register accumulation; FX_TIGHT_SIMT=1 for the CUDA-core kernel): latency per batch and images/s,
    device-resident inputs, CUDA events, 3 warm-ups, median of 10.
  * C4: 512x512 MRI-like gray sources (3 identical channels, and 1-channel gray carriage) -> 224: preprocess
    kernel GB/s (algorithmic bytes: source once + bf16 staging once) and end-to-end images/s.
Writes a markdown table to stdout (committed under profiles/).
"""
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from ssip_b200 import synthetic
from ssip_b200.engine import Engine, pack_images, uniform_descs
from ssip_b200.feature_extraction import _seeded_backbone

dev = torch.device("cuda", 0)
state = _seeded_backbone(1234, False).state_dict()


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return statistics.median(ms)


side = torch.cuda.Stream(dev)  # a capturable stream: steps of <= 64 images replay as one CUDA graph (FX_GRAPHS=0 disables)
torch.cuda.set_stream(side)
print("## C5 — batch sweep, 224x224x3 synthetic, 1xB200 (device-resident uint8 in, [B,512] out; median of 10)\n")
print("| batch | bf16 ms | bf16 img/s | fp32 ms | fp32 img/s |")
print("|---|---|---|---|---|")
pool = torch.randint(0, 256, (1024 * 150528,), dtype=torch.uint8, device=dev)
engines = {}
for prec, maxb in (("bf16", 1024), ("fp32", 1024)):
    e = Engine(0, maxb, prec)
    e.load_state_dict(state)
    engines[prec] = e
rows = []
for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
    descs = uniform_descs(b, 224, 224)
    cells = []
    for prec in ("bf16", "fp32"):
        e = engines[prec]
        if b > e.max_batch:
            cells += ["–", "–"]
            continue
        out = torch.empty((b, 512), dtype=torch.float32, device=dev)
        ms = timed(lambda: e.embed_device(pool[: b * 150528], descs, b, out=out))
        cells += [f"{ms:.3f}", f"{b / ms * 1e3:,.0f}"]
    print(f"| {b} | " + " | ".join(cells) + " |", flush=True)
for e in engines.values():
    e.close()

print("\n## C4 — 512x512 MRI-like sources -> 224 (Resize(256) antialiased, CenterCrop(224)), batch 256, bf16\n")
print("| carriage | source B/img | preprocess ms | algorithmic GB/s | frac of 6543 GB/s | end-to-end img/s (device-resident) |")
print("|---|---|---|---|---|---|")
e = Engine(0, 256, "bf16")
e.load_state_dict(state)
base = synthetic.mri_like_images(32, 512, seed=3)
imgs = np.concatenate([base] * 8, axis=0)  # 256 images
for name, arr in (("RGB (R==G==B), 3 channels", imgs), ("gray carriage, 1 channel", np.ascontiguousarray(imgs[..., 0]))):
    c = 3 if arr.ndim == 4 else 1
    d = torch.from_numpy(arr.reshape(-1)).to(dev)
    descs = uniform_descs(256, 512, 512, c)
    out = torch.empty((256, 512), dtype=torch.float32, device=dev)
    pre = timed(lambda: e.preprocess(d, descs, 256))
    full = timed(lambda: e.embed_device(d, descs, 256, out=out))
    byt = 512 * 512 * c + 224 * 224 * 3 * 2
    gbs = byt * 256 / (pre / 1e3) / 1e9
    print(f"| {name} | {512 * 512 * c:,} | {pre:.3f} | {gbs:,.0f} | {gbs / 6543.1:.3f} | {256 / full * 1e3:,.0f} |", flush=True)
e.close()
