"""Exploratory: which (shift, base_offset) combinations give a correct row-shifted UMMA operand."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from ssip_b200.engine import Engine

eng = Engine(0, 8, "bf16")
for kb in (64, 32, 16):
    g = torch.Generator().manual_seed(kb)
    a = torch.randint(-4, 5, (256, kb), generator=g).float()
    b = torch.randint(-4, 5, (64, kb), generator=g).float()
    ad, bd = a.to(torch.bfloat16).cuda(), b.to(torch.bfloat16).cuda()
    for shift in (0, 1, 2, 3, 5, 7, 8, 9, 13, 30, 58, 59, 116, 128):
        want = a[shift:shift + 128] @ b.T
        res = []
        for bo in sorted({0, shift & 7, (8 - shift) & 7}):
            got = eng.umma_shift(ad, bd, shift, bo).cpu()
            res.append(f"bo={bo}:{'OK' if torch.equal(got, want) else 'BAD(%d)' % int((got != want).sum())}")
        print(f"kb={kb} shift={shift:3d} " + " ".join(res), flush=True)
