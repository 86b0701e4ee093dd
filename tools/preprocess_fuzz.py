"""Randomised bit-exactness sweep of the GPU preprocess (both output forms) against the C restatement of the reference
transform (oracle/preprocess_oracle.c = Pillow's fixed-point resample + CenterCrop + ToTensor + Normalize,
src/feature_extraction.py:200-207,233-240).  Sizes: log-uniform 1..2600 per axis plus the boundaries around 224 / 256 / 512;
content: uniform noise (every tap matters); RGB and one-plane (gray carriage) sources; Resize(256)+CenterCrop(224) and the
classifier path's Resize((224, 224)).   usage: python tools/preprocess_fuzz.py [n_shapes] [out.md]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import reference_path as rp  # noqa: E402  (checker only)
from ssip_b200 import _native as N  # noqa: E402
from ssip_b200.engine import Engine, pack_images  # noqa: E402

n_shapes = int(sys.argv[1]) if len(sys.argv) > 1 else 600
out_md = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "preprocess_fuzz.md"
rng = np.random.default_rng(20261018)
edges = [1, 2, 3, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 111, 112, 113, 127, 128, 129, 223, 224, 225, 226, 255, 256, 257, 447, 448, 449, 511, 512,
         513, 1023, 1024, 1025]
shapes = []
while len(shapes) < n_shapes:
    pick = lambda: int(rng.choice(edges)) if rng.random() < 0.3 else int(np.exp(rng.uniform(0, np.log(2600))))
    h, w = max(1, pick()), max(1, pick())
    if h * w <= 2600 * 2600:
        shapes.append((h, w))

f32 = Engine(0, max_batch=64, precision="fp32")
b16 = Engine(0, max_batch=64, precision="bf16")
stats = {"images": 0, "nchw_exact": 0, "staging_exact": 0, "square_exact": 0, "unsupported": 0, "mismatch": []}
t0 = time.time()
for lo in range(0, n_shapes, 32):
    chunk = shapes[lo : lo + 32]
    imgs = []
    for k, (h, w) in enumerate(chunk):
        gray = (lo + k) % 5 == 4
        imgs.append(rng.integers(0, 256, (h, w) if gray else (h, w, 3), dtype=np.uint8))
    # geometries outside the kernels' range are refused loudly (the Python host resizes such files on the CPU first)
    keep, refs = [], []
    for im in imgs:
        buf, descs, total = pack_images([im])
        try:
            f32.preprocess_nchw(torch.from_numpy(buf[:total]).cuda(), descs, 1)
            keep.append(im)
        except N.FxError as exc:
            assert exc.status == N.FX_ERR_UNSUPPORTED, exc
            stats["unsupported"] += 1
    if not keep:
        continue
    buf, descs, total = pack_images(keep)
    dev = torch.from_numpy(buf[:total]).cuda()
    got = f32.preprocess_nchw(dev, descs, len(keep)).cpu()
    b16.preprocess(dev, descs, len(keep))
    _, padded = b16.staged_crop(len(keep))
    padded = padded.cpu()
    for i, im in enumerate(keep):
        rgb = im if im.ndim == 3 else np.repeat(im[:, :, None], 3, 2)
        want = torch.from_numpy(rp.c_preprocess_rgb(rgb))
        stats["images"] += 1
        ok1 = torch.equal(got[i], want)
        ok2 = torch.equal(padded[i, :, 3:227, 3:227], want.to(torch.bfloat16))
        stats["nchw_exact"] += ok1
        stats["staging_exact"] += ok2
        if not (ok1 and ok2):
            stats["mismatch"].append((im.shape, bool(ok1), bool(ok2)))
    # the classifier path's transform, Resize((224, 224)) with no crop (src/training/common.py:111-117), RGB sources only
    rgbs = [im for im in keep if im.ndim == 3]
    if rgbs:
        buf, descs, total = pack_images(rgbs)
        f32.set_transform(N.TRANSFORM_SQUARE224)
        try:
            sq = f32.preprocess_nchw(torch.from_numpy(buf[:total]).cuda(), descs, len(rgbs)).cpu()
            for i, im in enumerate(rgbs):
                ok3 = torch.equal(sq[i], torch.from_numpy(rp.c_preprocess_square224(im)))
                stats["square_n"] = stats.get("square_n", 0) + 1
                stats["square_exact"] += ok3
                if not ok3:
                    stats["mismatch"].append((im.shape, "square224"))
        except N.FxError as exc:
            assert exc.status == N.FX_ERR_UNSUPPORTED, exc
            stats["square_refused"] = stats.get("square_refused", 0) + 1
        finally:
            f32.set_transform(N.TRANSFORM_EXTRACT)
print(f"{stats['images']} images in {time.time() - t0:.1f}s: fp32 NCHW exact {stats['nchw_exact']}, bf16 staging exact {stats['staging_exact']}, "
      f"refused (FX_ERR_UNSUPPORTED) {stats['unsupported']}, Resize((224,224)) exact {stats['square_exact']} of {stats.get('square_n', 0)} "
      f"(batches refused: {stats.get('square_refused', 0)}), mismatches {stats['mismatch'][:10]}")
hs = np.array([s[0] for s in shapes]); ws = np.array([s[1] for s in shapes])
out_md.parent.mkdir(parents=True, exist_ok=True)
out_md.write_text(
    "# Preprocess: randomised bit-exactness sweep (B200)\n\n"
    f"`python tools/preprocess_fuzz.py {n_shapes}`: {n_shapes} random source sizes (heights {hs.min()}..{hs.max()}, widths {ws.min()}..{ws.max()}, "
    f"aspect ratios {float((hs / ws).min()):.3f}..{float((hs / ws).max()):.1f}; 30 % of the axes on the boundaries around 224 / 256 / 512 / 1024), uniform-noise "
    "content, every fifth image a one-plane (gray carriage) source; checker = `oracle/preprocess_oracle.c` (pinned against the real "
    "torchvision / Pillow transform by `tests/test_oracle_cpu.py`).\n\n"
    f"- images checked: {stats['images']}\n- `fx_preprocess_nchw_f32` bit-exact: {stats['nchw_exact']}\n"
    f"- bf16 conv1 staging tensor == bf16(reference tensor), padding zero: {stats['staging_exact']}\n"
    f"- `Resize((224, 224))` (classifier path, RGB sources) bit-exact: {stats['square_exact']} of {stats.get('square_n', 0)} (batches refused as a whole because one member is out of range: {stats.get('square_refused', 0)})\n"
    f"- refused with FX_ERR_UNSUPPORTED (down-scaling factor beyond the kernels' band; the Python host resizes those on the CPU first): {stats['unsupported']}\n"
    f"- mismatches: {stats['mismatch'] if stats['mismatch'] else 'none'}\n")
f32.close(); b16.close()
sys.exit(1 if stats["mismatch"] else 0)
