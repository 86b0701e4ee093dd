"""How exact is the tensor cores' fp32 accumulation?  (Input to the tight-mode design)
bf16 x bf16 products are exact in fp32, so a bf16 conv whose fp32 accumulator is written out unrounded differs from an
fp64 evaluation of the same bf16-rounded operands ONLY by the accumulation: round-to-nearest adds give ~1e-7 relative
error whatever K is, truncating adds (as reported for earlier tensor-core generations) a bias that grows with K / 16.
    python tools/probe_accum.py"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ssip_b200.engine import Engine  # noqa: E402

eng = Engine(0, max_batch=64, precision="bf16")
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)  # noqa: E731
for cin, cout, hin, n, positive in [(64, 64, 56, 2, False), (128, 128, 28, 4, False), (256, 256, 14, 16, False), (512, 512, 7, 32, False),
                                    (512, 512, 7, 32, True), (64, 64, 56, 2, True)]:
    g = torch.Generator().manual_seed(cin + hin)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    x = torch.randn(n, hin, hin, cin, generator=g)
    if positive:  # every product has the same sign: the partial sums grow monotonically, the worst case for truncation
        w, x = w.abs(), x.abs()
    got = eng.debug_conv(w, None, 1, 1, x.cuda(), None, relu=False, f32_out=True).cpu().double()
    # the library folds BN (identity here: gamma 1, var 1 - 1e-5 + eps) in fp64 and rounds the folded weight to bf16
    s = 1.0 / torch.sqrt(torch.ones(cout, dtype=torch.float64) - 1e-5 + 1e-5)
    wf = bf((w.double() * s[:, None, None, None]).float())
    want = F.conv2d(bf(x).permute(0, 3, 1, 2).double(), wf.double(), None, stride=1, padding=1).permute(0, 2, 3, 1)
    f32 = F.conv2d(bf(x).permute(0, 3, 1, 2), wf, None, stride=1, padding=1).permute(0, 2, 3, 1).double()
    rel = float((got - want).norm() / want.norm())
    rel32 = float((f32 - want).norm() / want.norm())
    signed = float(((got - want) * want.sign()).mean() / want.abs().mean())
    print(f"K={cin * 9:5d} ({cin}->{cout}, {hin}x{hin}, {'all-positive' if positive else 'zero-mean'}): tensor-core accumulate relL2 {rel:.3e}, "
          f"signed mean error {signed:+.3e} of mean |y|;  CPU fp32 conv relL2 {rel32:.3e}", flush=True)
eng.close()
