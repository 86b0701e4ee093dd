"""Why is the host-buffer pipeline ~10 % below the device-resident loop?  Variants of the same 256-image step."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ssip_b200.engine import Engine, uniform_descs
from ssip_b200.feature_extraction import _seeded_backbone

B, IMG = 256, 150528
dev = torch.device("cuda", 0)
eng = Engine(0, max_batch=B, precision="bf16")
eng.load_state_dict(_seeded_backbone(1234, False).state_dict())
descs = uniform_descs(B, 224, 224)
pool = torch.randint(0, 256, (64 * B * IMG,), dtype=torch.uint8, device=dev)
host_in = [torch.randint(0, 256, (B * IMG,), dtype=torch.uint8).pin_memory() for _ in range(8)]
host_out = [torch.empty((B, 512)).pin_memory() for _ in range(4)]
out = torch.empty((4 * B, 512), device=dev)
s1 = torch.cuda.Stream(dev)
streams = [torch.cuda.current_stream(dev), s1]

def device_loop(k, sync_every=0):
    evs = []
    for i in range(k):
        lane = i % 2
        eng.select_lane(lane)
        with torch.cuda.stream(streams[lane]):
            eng.embed_device(pool[(i % 64) * B * IMG:((i % 64) + 1) * B * IMG], descs, B, out=out[(i % 4) * B:(i % 4 + 1) * B])
            if sync_every:
                e = torch.cuda.Event(); e.record(); evs.append(e)
                if len(evs) > sync_every: evs.pop(0).synchronize()
    eng.select_lane(0)
    torch.cuda.synchronize()

def host_loop(k):
    for i in range(k):
        slot = i % 4
        eng.embed_host_wait(slot)
        eng.embed_host_async(slot, host_in[i % 8], descs, B, B * IMG, host_out[slot])
    for s in range(4): eng.embed_host_wait(s)
    torch.cuda.synchronize()

for name, fn in (("device loop, free running", lambda k: device_loop(k)), ("device loop, host waits for step i-4", lambda k: device_loop(k, 4)),
                 ("host-buffer slots", host_loop)):
    for k in (50, 400):
        fn(10)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(k); best = min(best, time.perf_counter() - t0)
        print(f"{name:45s} k={k:4d}: {k * B / best:,.0f} images/s ({best / k * 1e3:.3f} ms/step)", flush=True)
eng.close()
