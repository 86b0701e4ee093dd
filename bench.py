#!/usr/bin/env python
"""Benchmark of the feature-extraction hot path (BASELINE.json metric: ResNet-18 embed images/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (fused preprocess -> ResNet-18 trunk -> [B,512] embeddings) over
one batch of synthetic 224x224x3 uint8 images.  One JSON line on stdout (rank 0).

  value     images/s, whole job, inputs already resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the host-buffer C-ABI call fx_embed_host: pinned host uint8 in, host
            fp32 [B,512] out, both copies inside the timed region
  roofline  the trunk's implicit-GEMM kernels against the measured dense-bf16 peak
            (+ roofline_preprocess: the fused preprocess kernel against the measured HBM peak)
  cpu_baseline  the oracle port of the reference's CPU path on this box's host cores (N=1, rank 0)

--impl reference times that CPU port alone (the reference is pure Python + torchvision and
/root/reference does not exist on the GPU box; oracle/reference_path.py restates its loop over the
same third-party calls).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

IMG_H = IMG_W = 224
IMG_BYTES = IMG_H * IMG_W * 3
POOL_IMAGES = 50_000  # BASELINE.json configs[1]
TRUNK_FLOP_PER_IMAGE = 2 * 1_813_561_344  # 1.8136 GMAC, 2 FLOP/MAC
PRE_BYTES_BF16 = IMG_BYTES + 224 * 224 * 3 * 2  # source read once + bf16 NHWC(3) written once
WEIGHT_SEED = 1234

# MACs per image of the 20 conv groups in fx_load_weights order and the kernel family that runs each
_M = 115_605_504
LAYER_MACS = [118_013_952, _M, _M, _M, _M, _M // 2, _M, 6_422_528, _M, _M, _M // 2, _M, 6_422_528, _M, _M, _M // 2, _M, 6_422_528, _M, _M]
assert sum(LAYER_MACS) == 1_813_561_344
LAYER_FAMILY = (["stemw2_conv_kernel<cta_group::2> (stem conv1+bn+relu+maxpool, s2d 4x4, two outputs per accumulator row, N=128)"] + ["flat2w_conv_kernel<cta_group::2> (layer1 3x3/s1, two outputs per accumulator row, N=128; residual prefetched into registers)"] * 4 +
                ["tc_conv_kernel + tc2_conv_kernel<cta_group::2> (stride-2 3x3, 1x1/s2, layer3, layer4; per-tap TMA)", "flat128x2_conv_kernel<cta_group::2> (layer2 3x3/s1, N=128)",
                 "tc_conv_kernel + tc2_conv_kernel<cta_group::2> (stride-2 3x3, 1x1/s2, layer3, layer4; per-tap TMA)", "flat128x2_conv_kernel<cta_group::2> (layer2 3x3/s1, N=128)",
                 "flat128x2_conv_kernel<cta_group::2> (layer2 3x3/s1, N=128)"] + ["tc_conv_kernel + tc2_conv_kernel<cta_group::2> (stride-2 3x3, 1x1/s2, layer3, layer4; per-tap TMA)"] * 10)
EXEC_ORDER = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 13, 14, 15, 16, 18, 19]  # launch order of the slots inside fx_forward (the 1x1 downsample slots 7, 12, 17 ride in the launches of slots 5, 10, 15)
LAUNCH_PROFILE = ROOT / "profiles" / "r02_launches.csv"  # ncu launch list (batch 256) used for the DRAM-traffic column


def profiled_traffic():
    """DRAM bytes (read + write) per launch from the committed ncu launch list: {slot: bytes}, {'preprocess': bytes}."""
    import csv

    if not LAUNCH_PROFILE.exists():
        return None
    rows = list(csv.DictReader(l for l in open(LAUNCH_PROFILE) if not l.startswith("==")))
    per_id = {}
    for r in rows:
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            per_id.setdefault(r["ID"], [r["Kernel Name"], 0.0])[1] += float(r["Metric Value"].replace(",", ""))
    launches = list(per_id.values())
    if len(launches) != 2 + len(EXEC_ORDER) or "preprocess" not in launches[0][0]:
        return None
    out = {"preprocess": launches[0][1], 7: 0.0, 12: 0.0, 17: 0.0}
    for k, slot in enumerate(EXEC_ORDER):
        out[slot] = launches[1 + k][1]
    return out


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]),
                "tf_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_port_rate(images: np.ndarray, batch: int, budget_s: float, threads: int):
    """Oracle port of the reference loop (preprocess_image per file + stack + model forward,
    src/feature_extraction.py:272-300) on decoded arrays; returns (images/s, images timed, seconds)."""
    from oracle import reference_path as rp

    torch.set_num_threads(threads)
    transform = rp.port_transform()
    model = rp.port_model(torch.device("cpu"), WEIGHT_SEED, False)

    def one_batch(chunk):
        x = torch.stack([rp.port_preprocess_array(a, transform) for a in chunk])
        with torch.no_grad():
            return torch.flatten(model(x), 1).numpy()

    one_batch(images[:batch])  # warm-up (oneDNN primitive creation)
    done, pos, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:  # cycle through the sample until the time budget is used
        chunk = images[pos : pos + batch]
        one_batch(chunk)
        done += len(chunk)
        pos = pos + batch if pos + batch < len(images) else 0
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def synth_host(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, IMG_H, IMG_W, 3), dtype=np.uint8)


def run_reference(args, rank: int, json_fd: int):
    """--impl reference: the CPU path alone, rank 0 only."""
    if rank != 0:
        return
    cores = host_cores()
    batch = 32  # the reference's own default batch (src/feature_extraction.py:68)
    from oracle import reference_path as rp

    torch.set_num_threads(cores)
    transform = rp.port_transform()
    model = rp.port_model(torch.device("cpu"), WEIGHT_SEED, False)
    imgs = synth_host(batch, 0)

    def step(k):
        x = torch.stack([rp.port_preprocess_array(a, transform) for a in imgs[:k]])
        with torch.no_grad():
            return torch.flatten(model(x), 1).numpy()

    t0 = time.perf_counter()
    step(batch)
    first = time.perf_counter() - t0
    # size the per-step sample so that steps+warmup end within ~150 s
    t0 = time.perf_counter()
    step(batch)
    per_img = (time.perf_counter() - t0) / batch
    total_steps = args.steps + args.warmup
    k = int(max(4, min(batch, 150.0 / max(total_steps, 1) / per_img)))
    for _ in range(args.warmup):
        step(k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(k)
    dt = time.perf_counter() - t0
    value = args.steps * k / dt
    sample = (f"{k} synthetic 224x224x3 images per step x {args.steps} steps, fp32, torch threads={cores} (first call {first:.2f}s untimed); "
              "decoded arrays in memory, no PNG/JPEG decode (the GPU arm also starts from decoded bytes)")
    line = {
        "impl": "reference", "metric": "ResNet-18 embed images/sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line, json_fd)


def workload_config(args, world: int):
    if world == 1:
        name = f"ResNet-18 embedding extraction, 50k synthetic 224x224 images, batch {args.batch}, {args.precision} on 1xB200"
    else:
        name = (f"Image-sharded extraction of synthetic 224x224 images at {world}xB200, batch {args.batch}/GPU, "
                f"NCCL all-gather of [N,512] embeddings")
    return {"workload": name, "batch_per_gpu": args.batch, "image": "224x224x3 uint8", "precision": args.precision,
            "weights": f"torchvision resnet18 random-init seed {WEIGHT_SEED}",
            "l2": "every step reads a different batch of a resident uint8 pool larger than L2",
            "parallelism": f"image-sharded dp{world}", "lanes": max(1, min(2, args.lanes))}


def bind_to_gpu_numa_node(gpu_index: int) -> None:
    """Pin this process to the CPU cores NVML reports as local to its GPU, so that the pinned host buffers
    (first touch) and the H2D copies stay on the GPU's socket.  Best effort: silently skipped without NVML."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def emit(line: dict, fd: int) -> None:
    """The ONE JSON line, written to the process's original stdout (everything else, including what NCCL or
    other native libraries print on fd 1, has been redirected to stderr)."""
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    # keep stdout clean for the one JSON line: native libraries (e.g. "NCCL version ...") write to fd 1 as well
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default 256 at 1 GPU, 512 at N>1)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--pool", type=int, default=0, help="resident images per GPU (default 50000 / enough for L2 rule)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline work (N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=2, help="engine lanes (1 or 2) the device-resident steps alternate between")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.batch <= 0:
        args.batch = 256 if world == 1 else 512

    if args.impl == "reference":
        run_reference(args, rank, json_fd)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference for the CPU arm)")

    from ssip_b200.engine import Engine, uniform_descs
    from ssip_b200.feature_extraction import _seeded_backbone

    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, K, W = args.batch, args.steps, args.warmup
    B512 = 512  # per-GPU batch of the N > 1 runs (BASELINE.json configs[2]); also timed at N = 1 for a like-for-like scaling base
    eng = Engine(local_rank, max_batch=max(B, B512) if world == 1 else B, precision=args.precision)
    eng.load_state_dict(_seeded_backbone(WEIGHT_SEED, False).state_dict())

    # resident synthetic pool (seeded, generated on the device): C2 keeps 50k images = 7.5 GB in HBM
    pool_n = args.pool or POOL_IMAGES
    pool_n = max(B, (pool_n // B) * B)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    pool = torch.empty(pool_n * IMG_BYTES, dtype=torch.uint8, device=dev)
    chunk = 2048 * IMG_BYTES
    for off in range(0, pool.numel(), chunk):
        hi = min(pool.numel(), off + chunk)
        pool[off:hi] = torch.randint(0, 256, (hi - off,), dtype=torch.uint8, device=dev, generator=gen)
    n_batches = pool_n // B
    descs = uniform_descs(B, IMG_H, IMG_W)
    out_local = torch.empty((K * B, 512), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * K * B, 512), dtype=torch.float32, device=dev) if world > 1 else None

    def batch_view(i):
        j = i % n_batches
        return pool[j * B * IMG_BYTES : (j + 1) * B * IMG_BYTES]

    n_lanes = max(1, min(2, args.lanes))
    lane_streams = [torch.cuda.current_stream(dev)] + [torch.cuda.Stream(dev) for _ in range(1, n_lanes)]

    def device_steps(k0, k1, out):
        # consecutive steps alternate between the engine's lanes (fx_select_lane: independent staging / activation
        # buffers) and streams: the kernels of step i+1 fill the SMs that the tail of each of step i's kernels leaves
        # idle (every kernel is a persistent one-CTA-per-SM grid)
        for i in range(k0, k1):
            lane = i % n_lanes
            eng.select_lane(lane)
            with torch.cuda.stream(lane_streams[lane]):
                eng.embed_device(batch_view(i), descs, B, out=out[(i - k0) * B : (i - k0 + 1) * B])
        eng.select_lane(0)
        for st in lane_streams[1:]:
            lane_streams[0].wait_stream(st)

    # ---- value: device-resident inputs -----------------------------------------------------------
    scratch = torch.empty((W * B, 512), dtype=torch.float32, device=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load0 = time.perf_counter()
    device_steps(0, W, scratch)
    barrier()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t_host0 = time.perf_counter()
    device_steps(W, W + K, out_local)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / K  # CPU time to queue one step (launch-bound if ~ ms_per_step)
    if world > 1:
        dist.all_gather_into_tensor(gathered, out_local)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    # keep the same load running until nvidia-smi has had >= 1.5 s of it to sample -- and time it: under the 1 kW power cap
    # the SM clock settles within a few hundred ms, so a K-step burst reads higher than the rate the GPU can hold
    # (DESIGN.md 4.7); `value_sustained` is the same loop over that >= 1 s window (rank 0's GPU)
    sus0, sus1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sus_steps = 0
    sus0.record()
    while time.perf_counter() - t_load0 < 1.5 or sus_steps < 4 * W:
        device_steps(0, W, scratch)
        sus_steps += W
        torch.cuda.synchronize()
    sus1.record()
    sus1.synchronize()
    value_sustained = world * sus_steps * B / (sus0.elapsed_time(sus1) / 1e3)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * B / (ms_max / 1e3)

    # ---- per-launch timing (separate pass, same stream, CUDA events inside the library: fx_profile_*) ----
    k_prof = min(K, 30)
    pre_ms, trunk_ms = [], []
    layer_ms = np.zeros(21, np.float64)
    emb = scratch[:B]
    eng.profile(True)
    for i in range(k_prof):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        src = batch_view(W + K + i)
        a.record()
        eng.preprocess(src, descs, B)
        b.record()
        eng.forward(B, emb)
        c.record()
        c.synchronize()
        pre_ms.append(a.elapsed_time(b))
        trunk_ms.append(b.elapsed_time(c))
        layer_ms += eng.profile_read()
    eng.profile(False)
    layer_ms /= k_prof
    # the trunk alone, back to back, without events between its launches (those inhibit the programmatic-dependent-
    # launch overlap of one kernel's prologue with the previous kernel's tail, so the per-launch times above are
    # upper bounds): this is the number the trunk roofline is computed from
    eng.forward(B, emb)
    torch.cuda.synchronize()
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    for _ in range(k_prof):
        eng.forward(B, emb)
    t1e.record()
    t1e.synchronize()
    trunk_b2b = t0e.elapsed_time(t1e) / k_prof
    # the same for the preprocess kernel: back to back over different batches (a launch bracketed by events on its own
    # also times the launch gap, ~10 us on a 40 us kernel)
    eng.preprocess(batch_view(0), descs, B)
    torch.cuda.synchronize()
    t0e.record()
    for i in range(k_prof):
        eng.preprocess(batch_view(W + K + i), descs, B)
    t1e.record()
    t1e.synchronize()
    pre_b2b = t0e.elapsed_time(t1e) / k_prof
    pre_avg, trunk_avg = statistics.mean(pre_ms), statistics.mean(trunk_ms)
    peaks = load_peaks()
    tf = TRUNK_FLOP_PER_IMAGE * B / (trunk_b2b / 1e3) / 1e12
    gbs = PRE_BYTES_BF16 * B / (pre_b2b / 1e3) / 1e9
    traffic = profiled_traffic() if B == 256 and args.precision == "bf16" else None
    families = {}
    for slot in range(20):
        f = families.setdefault(LAYER_FAMILY[slot], {"launches": 0, "ms": 0.0, "flop": 0.0, "traffic": 0.0})
        f["launches"] += 0 if slot in (7, 12, 17) else 1  # the downsample convs share the launch of their block's conv1
        f["ms"] += float(layer_ms[slot])
        f["flop"] += 2.0 * LAYER_MACS[slot] * B
        if traffic:
            f["traffic"] += traffic[slot]
    kernels = []
    for name, f in families.items():
        ach = f["flop"] / (f["ms"] / 1e3) / 1e12
        kernels.append({"kernel": name, "launches_per_step": f["launches"], "avg_launch_ms": f["ms"] / f["launches"],
                        "share_of_trunk": f["ms"] / float(layer_ms.sum()), "bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"],
                        "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"], "flop_per_launch": f["flop"] / f["launches"],
                        "traffic": (f["traffic"] / f["launches"]) if traffic else None})
    kernels.sort(key=lambda k: -k["share_of_trunk"])
    dominant = kernels[0]

    # ---- e2e: host buffers through fx_embed_host ---------------------------------------------------
    k_e2e = min(K, 200)  # the pipeline's fill and drain (one un-overlapped copy, one un-overlapped step) are part of the number: amortise them
    n_host = min(8, n_batches)
    host_in = [torch.from_numpy(synth_host(B, 77 + rank * 100 + j).reshape(-1)).pin_memory() for j in range(n_host)]
    n_slots = 4  # FX_HOST_SLOTS
    host_out = [torch.empty((B, 512), dtype=torch.float32).pin_memory() for _ in range(n_slots)]
    for j in range(3):
        eng.embed_host(host_in[j % n_host], descs, B, B * IMG_BYTES, out=host_out[0])
    h2d0 = eng.h2d_bytes
    e2e_runs = []
    for _ in range(3):  # median of three passes of k_e2e steps: one host hiccup (page fault, scheduler) must not decide the number
        barrier()
        t0 = time.perf_counter()
        for i in range(k_e2e):  # four pipeline slots over two lanes: two batches compute while the next copies are in flight
            slot = i % n_slots
            eng.embed_host_wait(slot)
            eng.embed_host_async(slot, host_in[i % n_host], descs, B, B * IMG_BYTES, host_out[slot])
        for slot in range(n_slots):
            eng.embed_host_wait(slot)
        torch.cuda.synchronize()
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = statistics.median(e2e_runs)
    h2d_per_step = (eng.h2d_bytes - h2d0) // (3 * k_e2e)  # counted by the library: the source rows the crop can touch
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * k_e2e * B / float(t.item())
    finite = bool(torch.isfinite(out_local).all().item()) and all(bool(np.isfinite(h.numpy()).all()) for h in host_out)

    # ---- the same e2e pass for R==G==B data carried as ONE plane (fx_image_desc.channels = 1, the MRI case): a third of the
    # H2D bytes, the one-plane preprocess kernel.  Extra line; the headline e2e above is the RGB workload. ----------------
    gray_descs = uniform_descs(B, IMG_H, IMG_W, 1)
    gray_in = [torch.from_numpy(np.random.default_rng(500 + rank * 10 + j).integers(0, 256, B * IMG_H * IMG_W, dtype=np.uint8)).pin_memory()
               for j in range(n_host)]
    for j in range(3):
        eng.embed_host(gray_in[j % n_host], gray_descs, B, B * IMG_H * IMG_W, out=host_out[0])
    gray_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for i in range(k_e2e):
            slot = i % n_slots
            eng.embed_host_wait(slot)
            eng.embed_host_async(slot, gray_in[i % n_host], gray_descs, B, B * IMG_H * IMG_W, host_out[slot])
        for slot in range(n_slots):
            eng.embed_host_wait(slot)
        torch.cuda.synchronize()
        gray_runs.append(time.perf_counter() - t0)
    t = torch.tensor([statistics.median(gray_runs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_gray_value = world * k_e2e * B / float(t.item())

    # ---- parity of the timed rows (outside the timed region): a seeded subsample of this rank's rows of the timed
    # steps against the CPU port of the reference path on the same pool images; max over ranks -----------------------
    from oracle import reference_path as rp  # the checker, never the thing measured

    prng = np.random.default_rng(4242 + rank)
    n_par = 32
    rows = sorted(prng.choice(K * B, size=min(n_par, K * B), replace=False).tolist())
    imgs = []
    for r in rows:
        step, j = divmod(r, B)
        img = ((W + step) % n_batches) * B + j
        imgs.append(pool[img * IMG_BYTES : (img + 1) * IMG_BYTES].cpu().numpy().reshape(IMG_H, IMG_W, 3))
    torch.set_num_threads(max(1, len(all_cpus) // world))
    want = rp.port_embed_arrays(imgs, seed=WEIGHT_SEED, randomize_bn=False)
    got = out_local[rows].cpu().numpy()
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
    par = torch.tensor([float(rel.max()), -float(cos.min())], dtype=torch.float64, device=dev)
    # the gathered matrix must hold every rank's rows, bit for bit: checksum of the int32 view per block
    gather_ok = None
    if world > 1:
        dist.all_reduce(par, op=dist.ReduceOp.MAX)
        mine = out_local.view(torch.int32).to(torch.int64).sum().reshape(1)
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, mine)
        blocks = gathered.view(world, K * B * 512).view(torch.int32).to(torch.int64).sum(dim=1)
        gather_ok = bool(torch.equal(blocks, sums)) and bool(torch.equal(gathered[rank * K * B : (rank + 1) * K * B], out_local))
    parity_max_rel_l2, parity_min_cos = float(par[0].item()), -float(par[1].item())

    # ---- N = 1 only: the same device-resident pass at batch 512 (the per-GPU batch of the N > 1 runs), so that scaling
    # efficiency can be computed like for like -----------------------------------------------------------------------
    value_b512 = None
    if world == 1 and B != B512 and pool_n >= 2 * B512:
        d512 = uniform_descs(B512, IMG_H, IMG_W)
        nb512 = pool_n // B512
        o512 = torch.empty((2 * B512, 512), dtype=torch.float32, device=dev)

        def steps512(k):
            for i in range(k):
                lane = i % n_lanes
                eng.select_lane(lane)
                with torch.cuda.stream(lane_streams[lane]):
                    j = i % nb512
                    eng.embed_device(pool[j * B512 * IMG_BYTES : (j + 1) * B512 * IMG_BYTES], d512, B512, out=o512[lane * B512 : (lane + 1) * B512])
            eng.select_lane(0)
            for st in lane_streams[1:]:
                lane_streams[0].wait_stream(st)

        steps512(4)
        torch.cuda.synchronize()
        k512 = K  # the N > 1 runs time K steps of this batch size: same region length, same share of it under the power cap
        a512, b512 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a512.record()
        steps512(k512)
        b512.record()
        b512.synchronize()
        value_b512 = k512 * B512 / (a512.elapsed_time(b512) / 1e3)

    # ---- the host leg of e2e: what the pinned-host -> device copies alone can deliver, all ranks copying at once
    # (plain cudaMemcpyAsync of one batch per call on one stream; CUDA events) ---------------------------------------
    h2d_dst = torch.empty(B * IMG_BYTES, dtype=torch.uint8, device=dev)
    for j in range(2):
        h2d_dst.copy_(host_in[j % n_host], non_blocking=True)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_copies = 24
    h0.record()
    for j in range(n_copies):
        h2d_dst.copy_(host_in[j % n_host], non_blocking=True)
    h1.record()
    h1.synchronize()
    h2d = torch.tensor([n_copies * B * IMG_BYTES / (h0.elapsed_time(h1) / 1e3) / 1e9], dtype=torch.float64, device=dev)
    h2d_min = h2d.clone()
    if world > 1:
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
        dist.all_reduce(h2d_min, op=dist.ReduceOp.MIN)
    h2d_mean_gbs, h2d_min_gbs = float(h2d.item()) / world, float(h2d_min.item())
    e2e_bytes_per_image = IMG_BYTES + 512 * 4

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)  # the CPU arm gets every host core again
        cores = host_cores()
        sample = synth_host(512, 5)
        rate, done, secs = cpu_port_rate(list(sample), 32, args.cpu_budget, cores)
        cpu_baseline = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": f"{done} synthetic 224x224x3 images, batch 32, fp32, {secs:.1f}s, oracle port of src/feature_extraction.py:272-300 on decoded arrays (no PNG/JPEG decode)"}

    if rank == 0:
        line = {
            "metric": "ResNet-18 embed images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": B * 512 * 4,
                    "h2d_note": f"of the {B * IMG_BYTES} B of a batch only the source rows Resize(256)+CenterCrop(224) reads are copied (one 2-D copy)",
                    "steps": k_e2e, "api": "fx_embed_host_async/wait, 4 slots over 2 lanes (pinned host uint8 in, host fp32 [B,512] out); median of 3 passes",
                    "pass_seconds": [round(x, 5) for x in e2e_runs]},
            "e2e_gray_carriage": {"value": e2e_gray_value, "unit": "images/s", "h2d_bytes_per_step": B * IMG_H * IMG_W, "d2h_bytes_per_step": B * 512 * 4,
                                  "note": "R==G==B sources carried as one plane (channels = 1): same path, a third of the host -> device bytes"},
            "gpu_launches": int(launches * world),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "lanes": n_lanes,
            # dominant kernel family of the step (largest share of device time), timed per launch by CUDA events
            "roofline": {"bound": "tensor", "achieved": dominant["achieved"], "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": dominant["frac"], "traffic": dominant["traffic"], "kernel": dominant["kernel"],
                         "flop_per_launch": dominant["flop_per_launch"], "avg_launch_ms": dominant["avg_launch_ms"],
                         "launches_per_step": dominant["launches_per_step"], "share_of_trunk": dominant["share_of_trunk"],
                         "peak_src": peaks["src"] + " sustained bf16 (kernel timed inside a long step)",
                         "traffic_src": str(LAUNCH_PROFILE.relative_to(ROOT)) if dominant["traffic"] else None},
            "roofline_trunk": {"bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                               "frac": tf / peaks["tf_sustained"], "frac_of_burst_peak": tf / peaks["tf_burst"],
                               "kernel": "whole trunk (17 conv launches incl. grouped downsamples + avgpool per batch), run back to back",
                               "flop_per_image": TRUNK_FLOP_PER_IMAGE, "avg_ms": trunk_b2b, "avg_ms_with_per_launch_events": trunk_avg},
            "roofline_kernels": kernels,
            "layer_ms": [round(float(x), 5) for x in layer_ms],  # slots 0..19 = conv groups in fx_load_weights order, 20 = avgpool
            "roofline_preprocess": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                    "traffic": traffic["preprocess"] if traffic else None, "kernel": "preprocess_s2d_kernel (bf16 space-to-depth staging)", "bytes_per_image": PRE_BYTES_BF16,
                                    "avg_ms": pre_b2b, "avg_ms_single_launch_with_events": pre_avg, "peak_src": peaks["src"]},
            # the host-side ceiling of e2e: every image costs 150,528 B of H2D (+ 2 KB of D2H); `peak` is what plain pinned copies
            # deliver per GPU with all ranks copying at once, `achieved` what the pipelined e2e pass moved per GPU
            "roofline_e2e": {"bound": "pcie-h2d", "achieved": e2e_value / world * (h2d_per_step / B) / 1e9, "peak": h2d_mean_gbs, "unit": "GB/s per GPU",
                             "frac": (e2e_value / world * (h2d_per_step / B) / 1e9) / h2d_mean_gbs, "peak_min_over_ranks": h2d_min_gbs,
                             "images_per_s_ceiling": world * h2d_mean_gbs * 1e9 / (h2d_per_step / B), "bytes_per_image": e2e_bytes_per_image,
                             "how": f"{n_copies} cudaMemcpyAsync of {B * IMG_BYTES} B from pinned host memory per rank, all ranks at once, CUDA events"},
            "value_sustained": value_sustained, "value_sustained_steps": sus_steps,
            "value_b512_n1": value_b512,
            "parity_max_rel_l2": parity_max_rel_l2, "parity_min_cos": parity_min_cos,
            "parity": {"rows_per_rank": len(rows), "against": "oracle port of src/feature_extraction.py:272-300 (CPU fp32) on the same pool images, "
                       "rows of the timed steps, max over ranks", "tolerance": "relL2 <= 1e-2, cos >= 0.999 (bf16); 1e-5 (fp32)",
                       "gathered_blocks_equal_rank_rows": gather_ok},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "finite": finite,
        }
        emit(line, json_fd)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
