"""What does copying only the crop's source window buy on the H2D leg?  fx_embed_host_async with the compute skipped
(FX_DEBUG_E2E=4) and with it, for the copy shapes: whole buffer, the crop's rows (one 2-D copy, default).  One
interpreter per setting (the knobs are read once).   usage: python tools/h2d_window_probe.py [batch]
Measured on one B200 (batch 256): whole buffer 55.2 GB/s = 367 k images/s copy alone; crop rows 54.5 GB/s = 409 k images/s;
a third shape -- the crop's rows AND columns as one 3-D copy of 588-byte row pieces (117,612 B/image) -- was tried with
a temporary knob and ran at 16.4 GB/s = 139 k images/s: the copy engines do not like short pieces; not kept."""
import os, subprocess, sys, time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r"""
import sys, time
sys.path.insert(0, sys.argv[1])
import torch
from ssip_b200.engine import Engine, uniform_descs
from ssip_b200.feature_extraction import _seeded_backbone
B, IMG = int(sys.argv[2]), 150528
eng = Engine(0, max_batch=B, precision="bf16")
eng.load_state_dict(_seeded_backbone(1234, False).state_dict())
descs = uniform_descs(B, 224, 224)
host_in = [torch.randint(0, 256, (B * IMG,), dtype=torch.uint8).pin_memory() for _ in range(8)]
host_out = [torch.empty((B, 512)).pin_memory() for _ in range(4)]
def loop(k):
    for i in range(k):
        slot = i % 4
        eng.embed_host_wait(slot)
        eng.embed_host_async(slot, host_in[i % 8], descs, B, B * IMG, host_out[slot])
    for s in range(4): eng.embed_host_wait(s)
    torch.cuda.synchronize()
loop(20)
best = 1e9
for _ in range(3):
    b0 = eng.h2d_bytes; t0 = time.perf_counter(); loop(200); dt = time.perf_counter() - t0; nbytes = eng.h2d_bytes - b0
    best = min(best, dt)
print(f"{200 * B / best:,.0f} images/s, {nbytes / best / 1e9:.1f} GB/s, {nbytes // (200 * B)} B/image", flush=True)
eng.close()
"""
batch = sys.argv[1] if len(sys.argv) > 1 else "256"
for name, env in (("whole buffer", {"FX_H2D_ROWS": "0"}), ("crop rows (2-D copy)", {})):
    for what, dbg in (("copy alone", "4"), ("copy + compute", None)):
        e = dict(os.environ)
        for k in ("FX_H2D_ROWS", "FX_H2D_COLS", "FX_DEBUG_E2E"):
            e.pop(k, None)
        e.update(env)
        if dbg:
            e["FX_DEBUG_E2E"] = dbg
        out = subprocess.run([sys.executable, "-c", CHILD, str(ROOT), batch], env=e, capture_output=True, text=True)
        print(f"{name:32s} {what:15s}: {out.stdout.strip() or out.stderr[-300:]}", flush=True)
