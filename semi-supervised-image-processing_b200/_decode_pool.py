"""Host decode pool for the drop-in's file path ("multi-process host decode pool writing u8
into pinned memory").

The reference decodes every file inside its batch loop on the one Python thread (``Image.open`` ... ``transform``,
src/feature_extraction.py:233-240, 276-284).  Pillow's JPEG decoder does not scale on threads (the GIL is held for
most of a 512x512 decode: 2 threads are no faster than 1), so this pool is PROCESSES: spawned workers that import
only numpy + Pillow, receive file paths plus byte offsets, decode with the very same Pillow call the reference makes
and write the pixels straight into a shared-memory staging buffer that the parent has page-locked for the H2D copy.
The decoded bytes are Pillow's own: nothing about parity changes.

Two passes per batch: ``probe`` reads the headers (size / mode: lazy ``Image.open``, no pixel decode) so that the
parent can lay the images out; ``decode`` fills the buffer.  A file whose pixel data is broken fails in the second
pass and simply leaves a hole (its descriptor is dropped), with the reference's error handling preserved by the
caller: ``UnidentifiedImageError`` / ``OSError`` are reported per file, everything else propagates.
This module must stay importable without torch (spawn start-up time).
"""
from __future__ import annotations

import os
import pickle
import queue
import struct
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from multiprocessing import shared_memory
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from PIL import Image, UnidentifiedImageError

_ATTACHED: Dict[str, shared_memory.SharedMemory] = {}

# The preprocess kernel stages the source rows one output band needs in shared memory, which bounds the down-scaling
# factor it can take (csrc/preprocess.cu build_geom: about 23x, a short side near 5900 px).  Larger files are resized
# to the Resize(256) size right after the decode with the very call the reference's transform makes (torchvision
# functional.resize -> PIL Image.resize(BILINEAR)); the kernel then sees a 256-short-side image, for which Resize(256)
# is the identity, so the result is the reference's, bit for bit, for any size.
HOST_RESIZE_SHORT_SIDE = 4096
_RESIZE = 256


def resized_size(h: int, w: int) -> Tuple[int, int]:
    """(height, width) torchvision's Resize(256) gives an h x w image (transforms/functional.py:353-384)."""
    short, long = (w, h) if w <= h else (h, w)
    new_long = int(_RESIZE * long / short)
    return (new_long, _RESIZE) if w <= h else (_RESIZE, new_long)


def host_resize_if_oversized(img: Image.Image) -> Image.Image:
    if min(img.height, img.width) <= HOST_RESIZE_SHORT_SIDE:
        return img
    oh, ow = resized_size(img.height, img.width)
    return img.resize((ow, oh), Image.BILINEAR)


def rebuild_exception(name: str, text: str) -> BaseException:
    """The exception a worker reported as (class name, text), as an instance of the same class when it is a builtin
    or a Pillow one (so that process mode raises what thread mode and the reference raise), else RuntimeError."""
    import builtins

    import PIL

    for ns in (builtins, PIL, Image):
        cls = getattr(ns, name, None)
        if isinstance(cls, type) and issubclass(cls, BaseException):
            try:
                return cls(text)
            except Exception:  # noqa: BLE001 - constructor with another signature
                break
    return RuntimeError(f"{name}: {text}")


def _failure(exc: BaseException) -> Tuple[str, str, str]:
    kind = "decode" if isinstance(exc, (UnidentifiedImageError, OSError)) else "other"
    return (kind, type(exc).__name__, str(exc))


def _probe_chunk(paths: Sequence[str]):
    """Header pass: (height, width, bands, mode) per file, or a failure triple."""
    out = []
    for p in paths:
        try:
            with Image.open(p) as img:
                h, w = img.height, img.width
                if min(h, w) > HOST_RESIZE_SHORT_SIDE:  # the decode pass resizes these on the host (see above)
                    h, w = resized_size(h, w)
                out.append((h, w, len(img.getbands()), img.mode))
        except BaseException as exc:  # noqa: BLE001 - reported to the parent, which re-raises what the reference would
            out.append(_failure(exc))
    return out


def _decode_chunk(shm_name: str, jobs: Sequence[Tuple[str, int, int, int, int, bool]]):
    """jobs: (path, byte offset, height, width, bands, gray_carriage).  Returns per job the channel count written
    (3, or 1 for a gray carriage) or a failure triple."""
    shm = _ATTACHED.get(shm_name)
    if shm is None:
        for old in list(_ATTACHED):  # the parent replaced its buffer: drop stale mappings
            _ATTACHED.pop(old).close()
        shm = _ATTACHED[shm_name] = shared_memory.SharedMemory(name=shm_name)
        try:  # Python < 3.13 registers attached segments with this process's resource tracker, which would unlink
            from multiprocessing import resource_tracker  # the parent's buffer when the worker exits

            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:  # noqa: BLE001
            pass
    buf = np.frombuffer(shm.buf, dtype=np.uint8)
    out = []
    for path, off, h, w, bands, gray in jobs:
        try:
            with Image.open(path) as img:
                arr = np.asarray(host_resize_if_oversized(img))
            if arr.dtype != np.uint8 or arr.shape != ((h, w, bands) if bands > 1 else (h, w)):
                raise OSError(f"decoded array {arr.dtype}{arr.shape} does not match the header ({h}, {w}, {bands})")
            c = bands
            if gray and bands == 3 and (arr[..., 0] == arr[..., 1]).all() and (arr[..., 1] == arr[..., 2]).all():
                arr, c = arr[..., 0], 1
            buf[off : off + arr.size] = np.ascontiguousarray(arr).reshape(-1)
            out.append(c)
        except BaseException as exc:  # noqa: BLE001
            out.append(_failure(exc))
    return out


def _read_msg(stream):
    head = stream.read(8)
    if len(head) < 8:
        return None
    return pickle.loads(stream.read(struct.unpack("<Q", head)[0]))


def _write_msg(stream, obj) -> None:
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    stream.write(struct.pack("<Q", len(data)))
    stream.write(data)
    stream.flush()


def _worker_main() -> None:
    """Worker process: length-prefixed pickled (op, args) requests on stdin, results on stdout."""
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr  # nothing but protocol frames may reach the pipe
    while True:
        msg = _read_msg(inp)
        if msg is None:
            break
        op, args = msg
        _write_msg(out, _probe_chunk(*args) if op == "probe" else _decode_chunk(*args))
    for shm in _ATTACHED.values():
        shm.close()


class DecodePool:
    """Worker PROCESSES (plain ``python -c`` children talking pickle over pipes: no fork of a CUDA process, no
    re-import of the parent's ``__main__`` and therefore no torch import in the workers) + one shared-memory staging
    buffer per pipeline slot."""

    def __init__(self, workers: int, slots: int):
        self.workers = max(1, workers)
        root = str(Path(__file__).resolve().parent.parent)
        code = f"import sys; sys.path.insert(0, {root!r}); from ssip_b200._decode_pool import _worker_main; _worker_main()"
        self._procs = [subprocess.Popen([sys.executable, "-c", code], stdin=subprocess.PIPE, stdout=subprocess.PIPE)
                       for _ in range(self.workers)]
        self._free: "queue.Queue[subprocess.Popen]" = queue.Queue()
        for pr in self._procs:
            self._free.put(pr)
        self._threads = ThreadPoolExecutor(max_workers=self.workers)  # one blocking pipe conversation per worker
        self._shm: List[Optional[shared_memory.SharedMemory]] = [None] * slots
        self.chunk = 8  # files per request: amortises the round trip, still balances 256-file batches over 32 workers

    def _call(self, op: str, args):
        pr = self._free.get()
        try:
            _write_msg(pr.stdin, (op, args))
            res = _read_msg(pr.stdout)
        finally:
            self._free.put(pr)
        if res is None:
            raise RuntimeError("a decode worker exited unexpectedly")
        return res

    # -- staging buffers ------------------------------------------------------------------------
    def buffer(self, slot: int, nbytes: int) -> Tuple[shared_memory.SharedMemory, bool]:
        """Shared staging buffer of slot `slot` with room for nbytes; (buffer, True if it was (re)created)."""
        shm = self._shm[slot]
        if shm is not None and shm.size >= nbytes:
            return shm, False
        self.release(slot)
        shm = self._shm[slot] = shared_memory.SharedMemory(create=True, size=int(nbytes * 1.25) + 4096)
        return shm, True

    def release(self, slot: int) -> None:
        shm = self._shm[slot]
        if shm is not None:
            self._shm[slot] = None
            try:
                shm.close()
            except BufferError:  # a view of the buffer is still alive somewhere: leave the mapping, drop the name
                pass
            shm.unlink()

    # -- the two passes -------------------------------------------------------------------------
    def _chunks(self, items: Sequence):
        return [items[i : i + self.chunk] for i in range(0, len(items), self.chunk)]

    def probe(self, paths: Sequence[str]) -> list:
        res: list = []
        for part in self._threads.map(lambda c: self._call("probe", (c,)), self._chunks(list(paths))):
            res.extend(part)
        return res

    def decode(self, slot: int, jobs: Sequence[Tuple[str, int, int, int, int, bool]]) -> list:
        name = self._shm[slot].name
        res: list = []
        for part in self._threads.map(lambda c: self._call("decode", (name, c)), self._chunks(list(jobs))):
            res.extend(part)
        return res

    def close(self) -> None:
        self._threads.shutdown(wait=True)
        for pr in self._procs:
            try:
                pr.stdin.close()
                pr.wait(timeout=10)
            except Exception:  # noqa: BLE001
                pr.kill()
        for s in range(len(self._shm)):
            self.release(s)


def default_workers() -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    return max(1, min(32, n))
