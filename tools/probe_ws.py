"""Exploratory: tcgen05.mma.ws (weight-stationary) correctness of D layout and issue rate with B re-use."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from ssip_b200.engine import Engine

eng = Engine(0, 8, "bf16")
for kb in (64, 16):
    g = torch.Generator().manual_seed(kb)
    a = torch.randint(-4, 5, (256, kb), generator=g).float()
    b = torch.randint(-4, 5, (64, kb), generator=g).float()
    ad, bd = a.to(torch.bfloat16).cuda(), b.to(torch.bfloat16).cuda()
    for shift in (0, 5):
        want = a[shift:shift + 128] @ b.T
        got = eng.umma_shift(ad, bd, shift, 100).cpu()
        ok = torch.equal(got, want)
        print(f"ws kb={kb} shift={shift}: {'OK' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            # is it a row permutation / transposition of the expected result?
            rows = {tuple(r.tolist()): i for i, r in enumerate(want)}
            perm = [rows.get(tuple(r.tolist()), -1) for r in got]
            print("  row map (got row -> want row):", perm[:40])
            print("  got[0,:8]", got[0, :8].tolist(), "want[0,:8]", want[0, :8].tolist())
print("N rowb nacc ws | cycles/MMA")
for rowb in (128, 32):
    for n in (64, 128):
        for nacc, ws in ((1, 0), (2, 0), (1, 1), (2, 1), (2, 2), (4, 2)):
            if nacc * n > 512:
                continue
            r = eng.mma_rate(n, rowb, 3, 30 + 1000 * nacc + 100000 * ws, 4000)[:148].cpu()
            print(f"{n:3d} {rowb:4d} {nacc:2d} {ws} | {r.mean():7.1f} {r.min():7.1f} {r.max():7.1f}", flush=True)
