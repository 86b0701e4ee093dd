"""tcgen05.mma issue rate vs N, swizzle width, A-view shift and number of interleaved accumulators
(cycles per M128xNxK16 MMA, all SMs busy)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from ssip_b200.engine import Engine

eng = Engine(0, 8, "bf16")
print("N rowb shift tapstride n_acc | cycles/MMA (mean over SMs, min, max) | floor")
for rowb in (128, 64, 32):
    for n in (64, 128, 256):
        for shift, ts, nacc in ((0, 0, 1), (3, 30, 1), (3, 30, 2), (3, 30, 4), (3, 30, 8)):
            if nacc * n > 512:
                continue
            r = eng.mma_rate(n, rowb, shift, ts + 1000 * nacc, 3000)[:148].cpu()
            print(f"{n:3d} {rowb:4d} {shift:4d} {ts:4d} {nacc:2d} | {r.mean():7.1f} {r.min():7.1f} {r.max():7.1f} | {n // 2}", flush=True)
