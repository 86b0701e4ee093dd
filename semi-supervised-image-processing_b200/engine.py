"""Python handle on the CUDA engine (libfx_b200.so).  PyTorch is used here only for device memory
and streams; all compute happens inside the library's hand-written sm_100a kernels.

Stands in for ``load_model`` + the body of the batch loop of the reference
(src/feature_extraction.py:210-227, 289-294).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N

# fx_load_weights order (include/fx_b200.h): (conv weight key, bn prefix, stride, pad)
LAYER_TABLE: List[Tuple[str, str, int, int]] = [("conv1.weight", "bn1", 2, 3)]
for _s in range(1, 5):
    for _b in range(2):
        _stride = 2 if (_s > 1 and _b == 0) else 1
        LAYER_TABLE.append((f"layer{_s}.{_b}.conv1.weight", f"layer{_s}.{_b}.bn1", _stride, 1))
        LAYER_TABLE.append((f"layer{_s}.{_b}.conv2.weight", f"layer{_s}.{_b}.bn2", 1, 1))
        if _s > 1 and _b == 0:
            LAYER_TABLE.append((f"layer{_s}.{_b}.downsample.0.weight", f"layer{_s}.{_b}.downsample.1", 2, 0))
assert len(LAYER_TABLE) == N.NUM_CONV_LAYERS

_PRECISIONS = {"bf16": N.PRECISION_BF16, "fp32": N.PRECISION_FP32}
_ALIGN = 256


def pack_images(images: Sequence[np.ndarray], out: Optional[np.ndarray] = None):
    """Lay decoded HWC uint8 images end to end (256-byte aligned starts).

    Returns (buffer uint8 [total], descs ctypes array, total_bytes).  A 2-D array is a 1-channel
    gray carriage (see fx_image_desc in include/fx_b200.h).
    """
    n = len(images)
    descs = (N.ImageDesc * max(n, 1))()
    off = 0
    for i, a in enumerate(images):
        if a.dtype != np.uint8 or a.ndim not in (2, 3):
            raise TypeError(f"image {i}: expected HWC uint8, got {a.dtype} with shape {a.shape}")
        c = 1 if a.ndim == 2 else a.shape[2]
        descs[i].offset, descs[i].height, descs[i].width, descs[i].channels = off, a.shape[0], a.shape[1], c
        off += (a.size + _ALIGN - 1) // _ALIGN * _ALIGN
    total = off
    if out is None:
        out = np.empty(max(total, 1), np.uint8)
    elif out.size < total:
        raise ValueError("staging buffer too small")
    for i, a in enumerate(images):
        o = descs[i].offset
        out[o : o + a.size] = np.ascontiguousarray(a).reshape(-1)
    return out, descs, total


def uniform_descs(n: int, h: int, w: int, c: int = 3):
    """Descriptor table of n same-size images stored back to back without padding."""
    descs = (N.ImageDesc * max(n, 1))()
    for i in range(n):
        descs[i].offset, descs[i].height, descs[i].width, descs[i].channels = i * h * w * c, h, w, c
    return descs


class Engine:
    """One engine per GPU (``fx_create``)."""

    def __init__(self, device_index: int = 0, max_batch: int = 256, precision: str = "bf16"):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self._lib = N.lib()
        self.device_index = int(device_index)
        self.max_batch = int(max_batch)
        self.precision = precision
        self._h = ctypes.c_void_p()
        N.check(self._lib.fx_create(ctypes.byref(self._h), self.device_index, self.max_batch, _PRECISIONS[precision]))
        self.device = torch.device("cuda", self.device_index)
        self._slot_keep = {}

    # -- lifetime -----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.fx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status: int) -> None:
        N.check(status, self._h)

    def _stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def select_lane(self, lane: int) -> None:
        """Target lane (0/1) of the following preprocess/forward/embed_device calls: an independent set of staging and
        activation buffers, so that batches queued on different lanes and different CUDA streams overlap (fx_select_lane)."""
        self._check(self._lib.fx_select_lane(self._h, int(lane)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.fx_launch_count(self._h))

    @property
    def h2d_bytes(self) -> int:
        """Bytes the host-buffer calls have copied host -> device so far (fx_h2d_bytes)."""
        return int(self._lib.fx_h2d_bytes(self._h))

    def profile(self, on: bool) -> None:
        """Bracket every trunk launch of forward() with CUDA events (fx_profile_enable)."""
        self._check(self._lib.fx_profile_enable(self._h, int(on)))

    def profile_read(self) -> np.ndarray:
        """ms per launch of the last forward(): [0..19] conv groups in LAYER_TABLE order (0 = fused stem), [20] avgpool."""
        ms = np.zeros(N.NUM_CONV_LAYERS + 1, np.float32)
        self._check(self._lib.fx_profile_read(self._h, ms.ctypes.data, ms.size))
        return ms

    # -- weights ------------------------------------------------------------------------------
    def load_state_dict(self, state: Dict[str, torch.Tensor]) -> None:
        """Takes a torchvision resnet18 ``state_dict`` (fc.* ignored); BN is folded in the library."""
        table = (N.ConvBn * N.NUM_CONV_LAYERS)()
        keep = []

        def fptr(t: torch.Tensor):
            a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
            keep.append(a)
            return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))

        for i, (wkey, bn, stride, pad) in enumerate(LAYER_TABLE):
            w = state[wkey]
            e = table[i]
            e.weight = fptr(w)
            e.gamma, e.beta = fptr(state[bn + ".weight"]), fptr(state[bn + ".bias"])
            e.mean, e.var = fptr(state[bn + ".running_mean"]), fptr(state[bn + ".running_var"])
            e.eps = 1e-5  # torchvision BatchNorm2d default, resnet.py:43
            e.cout, e.cin, e.kh, e.kw = (int(x) for x in w.shape)
            e.stride, e.pad = stride, pad
        self._check(self._lib.fx_load_weights(self._h, table, N.NUM_CONV_LAYERS))

    def folded(self, layer: int, shape: Tuple[int, int, int, int]):
        """Folded (weight [cout,kh,kw,cin], bias [cout]) of a loaded layer as the kernels see them."""
        cout, kh, kw, cin = shape
        w = np.empty((cout, kh, kw, cin), np.float32)
        b = np.empty(cout, np.float32)
        self._check(self._lib.fx_debug_folded(self._h, layer, w.ctypes.data, b.ctypes.data))
        return w, b

    # -- preprocess ---------------------------------------------------------------------------
    def preprocess_nchw(self, packed_dev: torch.Tensor, descs, n: int) -> torch.Tensor:
        """Fused Resize/CenterCrop/ToTensor/Normalize -> fp32 [n,3,224,224] (the reference's batch tensor)."""
        self._dev_u8(packed_dev)
        out = torch.empty((n, 3, N.CROP, N.CROP), dtype=torch.float32, device=self.device)
        self._check(self._lib.fx_preprocess_nchw_f32(self._h, packed_dev.data_ptr(), descs, n, out.data_ptr(), self._stream()))
        return out

    # -- trunk --------------------------------------------------------------------------------
    def forward_nchw(self, x_dev: torch.Tensor) -> torch.Tensor:
        """Trunk only, on an already normalised fp32 [n,3,224,224] CUDA tensor."""
        if x_dev.dtype != torch.float32 or not x_dev.is_cuda or not x_dev.is_contiguous() or tuple(x_dev.shape[1:]) != (3, N.CROP, N.CROP):
            raise TypeError("expected a contiguous fp32 CUDA tensor [n,3,224,224]")
        n = x_dev.shape[0]
        out = torch.empty((n, N.EMBED_DIM), dtype=torch.float32, device=self.device)
        self._check(self._lib.fx_stage_nchw_f32(self._h, x_dev.data_ptr(), n, self._stream()))
        self._check(self._lib.fx_forward(self._h, n, out.data_ptr(), self._stream()))
        return out

    def preprocess(self, packed_dev: torch.Tensor, descs, n: int) -> None:
        """Fused transform into the engine's conv1 staging tensor (follow with forward())."""
        self._dev_u8(packed_dev)
        self._check(self._lib.fx_preprocess(self._h, packed_dev.data_ptr(), descs, n, self._stream()))

    def forward(self, n: int, out: torch.Tensor) -> torch.Tensor:
        """Trunk on the n staged images -> fp32 [n,512] written at `out` (CUDA, contiguous)."""
        self._check(self._lib.fx_forward(self._h, n, out.data_ptr(), self._stream()))
        return out

    # -- classifier head --------------------------------------------------
    def set_transform(self, transform: int) -> None:
        """N.TRANSFORM_EXTRACT (Resize(256)+CenterCrop(224)) or N.TRANSFORM_SQUARE224 (Resize((224,224)), the
        evaluation transform of src/training/common.py:111-117) for the following preprocess calls."""
        self._check(self._lib.fx_set_transform(self._h, int(transform)))

    def load_head(self, weight: torch.Tensor, bias: torch.Tensor) -> None:
        """fc of the fine-tuned model: weight [classes,512], bias [classes] (src/training/common.py:299-304)."""
        w = np.ascontiguousarray(weight.detach().to("cpu", torch.float32).numpy())
        b = np.ascontiguousarray(bias.detach().to("cpu", torch.float32).numpy())
        if w.ndim != 2 or w.shape[1] != N.EMBED_DIM or b.shape != (w.shape[0],):
            raise ValueError(f"head must be [classes,{N.EMBED_DIM}] + [classes], got {w.shape} + {b.shape}")
        self._check(self._lib.fx_load_head(self._h, w.ctypes.data, b.ctypes.data, int(w.shape[0])))
        self.num_classes = int(w.shape[0])

    def classify(self, n: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Trunk + head + softmax on the n staged images -> (embeddings [n,512], logits [n,C], probs [n,C]) on the device."""
        c = getattr(self, "num_classes", 0)
        emb = torch.empty((n, N.EMBED_DIM), dtype=torch.float32, device=self.device)
        logits = torch.empty((n, c), dtype=torch.float32, device=self.device)
        probs = torch.empty((n, c), dtype=torch.float32, device=self.device)
        self._check(self._lib.fx_classify(self._h, n, emb.data_ptr(), logits.data_ptr(), probs.data_ptr(), self._stream()))
        return emb, logits, probs

    def classify_nchw(self, x_dev: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """`model(inputs)` + softmax for an already transformed fp32 [n,3,224,224] CUDA batch (what the reference's
        DataLoader yields)."""
        if x_dev.dtype != torch.float32 or not x_dev.is_cuda or not x_dev.is_contiguous() or tuple(x_dev.shape[1:]) != (3, N.CROP, N.CROP):
            raise TypeError("expected a contiguous fp32 CUDA tensor [n,3,224,224]")
        self._check(self._lib.fx_stage_nchw_f32(self._h, x_dev.data_ptr(), x_dev.shape[0], self._stream()))
        return self.classify(x_dev.shape[0])

    def classify_device(self, packed_dev: torch.Tensor, descs, n: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Fused path: packed uint8 images on the device -> preprocess (current transform) -> trunk -> head."""
        self.preprocess(packed_dev, descs, n)
        return self.classify(n)

    # -- on-device post-processing of an [n,d] fp32 matrix ------------------
    @staticmethod
    def _matrix(x: torch.Tensor) -> Tuple[int, int]:
        if x.dtype != torch.float32 or not x.is_cuda or not x.is_contiguous() or x.dim() != 2:
            raise TypeError("expected a contiguous fp32 CUDA matrix [n,d]")
        return int(x.shape[0]), int(x.shape[1])

    def column_stats(self, x: torch.Tensor):
        """run_sanity_checks' scalars + StandardScaler's fit in one call: (stats dict, mean, std, var) with the column
        arrays as fp64 CUDA tensors [d] (fx_column_stats; fp64 accumulation, fixed reduction order)."""
        n, d = self._matrix(x)
        mean, std, var = (torch.empty(d, dtype=torch.float64, device=self.device) for _ in range(3))
        st = N.MatrixStats()
        self._check(self._lib.fx_column_stats(self._h, x.data_ptr(), n, d, mean.data_ptr(), std.data_ptr(), var.data_ptr(),
                                              ctypes.byref(st), self._stream()))
        return ({"nan_count": int(st.nan_count), "inf_count": int(st.inf_count), "mean_abs_mean": float(st.mean_abs_mean),
                 "mean_std": float(st.mean_std)}, mean, std, var)

    def standardize(self, x: torch.Tensor, mean: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
        """z = (x - fp32(mean)) / fp32(scale) in fp32, as scikit-learn's StandardScaler.transform on a float32 matrix."""
        n, d = self._matrix(x)
        if mean.dtype != torch.float64 or scale.dtype != torch.float64 or mean.numel() != d or scale.numel() != d:
            raise TypeError("mean / scale must be fp64 CUDA tensors [d]")
        out = torch.empty_like(x)
        self._check(self._lib.fx_standardize(self._h, x.data_ptr(), n, d, mean.data_ptr(), scale.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def neighbor_probe(self, x: torch.Tensor, query_rows: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
        """Cosine nearest neighbour (self excluded, first maximum) of each query row: (rows int64 [q], similarities fp32 [q])."""
        n, d = self._matrix(x)
        q = np.ascontiguousarray(np.asarray(query_rows, dtype=np.int64))
        nbr = np.empty(q.size, np.int64)
        sim = np.empty(q.size, np.float32)
        for lo in range(0, q.size, 32):  # fx_neighbor_probe keeps at most 32 queries in shared memory per pass
            qc, nc, sc = q[lo : lo + 32], nbr[lo : lo + 32], sim[lo : lo + 32]
            self._check(self._lib.fx_neighbor_probe(self._h, x.data_ptr(), n, d, qc.ctypes.data, int(qc.size), nc.ctypes.data,
                                                    sc.ctypes.data, self._stream()))
        return nbr, sim

    # -- whole path ---------------------------------------------------------------------------
    def embed_device(self, packed_dev: torch.Tensor, descs, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Preprocess + trunk on device-resident uint8 images -> fp32 [n,512] on the device."""
        self._dev_u8(packed_dev)
        if out is None:
            out = torch.empty((n, N.EMBED_DIM), dtype=torch.float32, device=self.device)
        elif out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous() or out.numel() < n * N.EMBED_DIM:
            raise TypeError("out must be a contiguous fp32 CUDA tensor with room for [n,512]")
        self._check(self._lib.fx_embed(self._h, packed_dev.data_ptr(), descs, n, out.data_ptr(), self._stream()))
        return out

    def embed_host(self, packed_host, descs, n: int, total_bytes: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host uint8 images in, host fp32 [n,512] out; both copies happen inside the call."""
        if isinstance(packed_host, torch.Tensor):
            src_ptr = packed_host.data_ptr()
        else:
            src_ptr = packed_host.ctypes.data
        if out is None:
            out = np.empty((n, N.EMBED_DIM), np.float32)
        dst_ptr = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
        self._check(self._lib.fx_embed_host(self._h, src_ptr, total_bytes, descs, n, dst_ptr))
        return out

    def embed_host_async(self, slot: int, packed_host, descs, n: int, total_bytes: int, out) -> None:
        """Queue one host batch on pipeline slot 0..HOST_SLOTS-1 (H2D overlaps earlier batches' kernels, which run on lane
        slot % 2); pair with embed_host_wait."""
        src_ptr = packed_host.data_ptr() if isinstance(packed_host, torch.Tensor) else packed_host.ctypes.data
        dst_ptr = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
        self._slot_keep[slot] = (packed_host, descs, out)  # the library reads/writes them until the wait
        self._check(self._lib.fx_embed_host_async(self._h, slot, src_ptr, total_bytes, descs, n, dst_ptr))

    def embed_host_async_dev(self, slot: int, packed_host, descs, n: int, total_bytes: int, out_dev: torch.Tensor) -> None:
        """embed_host_async with the rows left on the device: written straight at `out_dev` (a contiguous fp32 CUDA view
        with room for [n,512], e.g. this rank's slot of the all-gather buffer); no device->host copy."""
        if out_dev.dtype != torch.float32 or not out_dev.is_cuda or not out_dev.is_contiguous() or out_dev.numel() < n * N.EMBED_DIM:
            raise TypeError("out_dev must be a contiguous fp32 CUDA tensor with room for [n,512]")
        src_ptr = packed_host.data_ptr() if isinstance(packed_host, torch.Tensor) else packed_host.ctypes.data
        self._slot_keep[slot] = (packed_host, descs, out_dev)
        self._check(self._lib.fx_embed_host_async_dev(self._h, slot, src_ptr, total_bytes, descs, n, out_dev.data_ptr()))

    def embed_host_wait(self, slot: int) -> None:
        self._check(self._lib.fx_embed_host_wait(self._h, slot))
        self._slot_keep.pop(slot, None)

    # -- GPU JPEG decode ---------------------------------------------------
    _JPEG_BACKENDS = {"auto": N.JPEG_BACKEND_AUTO, "hardware": N.JPEG_BACKEND_HARDWARE, "gpu": N.JPEG_BACKEND_GPU}

    def jpeg_init(self, backend: str = "auto") -> str:
        """Load nvJPEG and create the batched decoder (fx_jpeg_init); returns the backend it runs on ("hardware" | "gpu").
        Raises FxError (FX_ERR_UNSUPPORTED) when nvJPEG cannot be had."""
        self._check(self._lib.fx_jpeg_init(self._h, self._JPEG_BACKENDS[backend]))
        return {N.JPEG_BACKEND_HARDWARE: "hardware", N.JPEG_BACKEND_GPU: "gpu"}[int(self._lib.fx_jpeg_backend(self._h))]

    def jpeg_probe(self, data: bytes) -> "N.FileInfo":
        info = N.FileInfo()
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
        self._check(self._lib.fx_jpeg_probe(self._h, ctypes.addressof(buf), len(data), ctypes.byref(info)))
        return info

    def jpeg_read_files(self, slot: int, paths: Sequence[str]):
        """Read one batch of files into the slot's page-locked bitstream buffer with the library's thread pool; returns
        the fx_file_info array (status: N.FILE_GPU_JPEG | N.FILE_HOST_DECODE | N.FILE_UNREADABLE, geometry)."""
        n = len(paths)
        arr = (ctypes.c_char_p * max(n, 1))(*[p.encode() if isinstance(p, str) else bytes(p) for p in paths])
        info = (N.FileInfo * max(n, 1))()
        self._check(self._lib.fx_jpeg_read_files(self._h, slot, arr, n, info))
        return info

    def jpeg_decode(self, blobs: Sequence[bytes], sizes: Sequence[Tuple[int, int]]) -> List[torch.Tensor]:
        """Decode JPEG bitstreams on the GPU (nvJPEG) -> list of uint8 CUDA tensors [h,w,3].  Test / study entry point."""
        n = len(blobs)
        descs = (N.ImageDesc * max(n, 1))()
        off = 0
        for i, (h, w) in enumerate(sizes):
            descs[i].offset, descs[i].height, descs[i].width, descs[i].channels = off, h, w, 3
            off += (h * w * 3 + 255) // 256 * 256
        out = torch.empty(max(off, 1), dtype=torch.uint8, device=self.device)
        keep = [(ctypes.c_char * len(b)).from_buffer_copy(b) for b in blobs]
        ptrs = (ctypes.c_void_p * max(n, 1))(*[ctypes.addressof(k) for k in keep])
        lens = (ctypes.c_size_t * max(n, 1))(*[len(b) for b in blobs])
        self._check(self._lib.fx_jpeg_decode(self._h, ptrs, lens, n, out.data_ptr(), descs, self._stream()))
        return [out[descs[i].offset : descs[i].offset + h * w * 3].view(h, w, 3) for i, (h, w) in enumerate(sizes)]

    def embed_files_async(self, slot: int, info, host_pixels: Sequence[Optional[np.ndarray]], descs, n: int, total_bytes: int, out) -> None:
        """One pipelined step over the files last read into `slot` (jpeg_read_files): info / descs list the n files that
        take part, host_pixels[i] is the caller-decoded array of an entry that is not N.FILE_GPU_JPEG (else None).
        `out`: pinned host tensor / array for the rows, or a CUDA tensor to leave them on the device."""
        ptrs = (ctypes.c_void_p * max(n, 1))(*[(a.ctypes.data if a is not None else None) for a in host_pixels])
        on_dev = isinstance(out, torch.Tensor) and out.is_cuda
        dst = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
        self._slot_keep[slot] = (info, host_pixels, descs, out, ptrs)
        self._check(self._lib.fx_embed_files_async(self._h, slot, info, ptrs, descs, n, total_bytes, None if on_dev else dst, dst if on_dev else None))

    def embed_images(self, images: Sequence[np.ndarray]) -> np.ndarray:
        """Convenience: list of decoded HWC uint8 arrays -> [n,512] (chunks of max_batch)."""
        chunks = []
        for s in range(0, len(images), self.max_batch):
            part = images[s : s + self.max_batch]
            buf, descs, total = pack_images(part)
            chunks.append(self.embed_host(buf, descs, len(part), total))
        return np.concatenate(chunks, axis=0) if chunks else np.empty((0, N.EMBED_DIM), np.float32)

    # -- test hooks ---------------------------------------------------------------------------
    def staged_crop(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """What fx_preprocess handed the trunk (bf16 engines): the raw space-to-depth staging tensor
        [n,115,116,16] and its rearrangement back to the zero-padded crop [n,3,230,232] (both bf16, on the device)."""
        if self.precision != "bf16":
            raise TypeError("staged_crop: bf16 engines only")
        raw = torch.empty((n, 115, 116, 16), dtype=torch.bfloat16, device=self.device)
        self._check(self._lib.fx_debug_staging(self._h, n, raw.data_ptr(), raw.numel() * 2, self._stream()))
        # channel (dy*2+dx)*3+c of s2d pixel (Y, X) = padded pixel (2Y+dy, 2X+dx), channel c
        v = raw[..., :12].reshape(n, 115, 116, 2, 2, 3).permute(0, 5, 1, 3, 2, 4).reshape(n, 3, 230, 232)
        return raw, v

    def debug_conv(self, weight: torch.Tensor, bn: Optional[Dict[str, torch.Tensor]], stride: int, pad: int,
                   x_nhwc: torch.Tensor, residual_nhwc: Optional[torch.Tensor] = None, relu: bool = False, f32_out: bool = False) -> torch.Tensor:
        """Run one conv+bn group through the library: x fp32 NHWC CUDA -> fp32 NHWC CUDA.  f32_out (bf16 engines): the
        fp32 accumulator (+bias, +residual, ReLU) itself instead of its bf16 rounding."""
        cout, cin, kh, kw = (int(v) for v in weight.shape)
        n, hin, win, _ = (int(v) for v in x_nhwc.shape)
        keep = []

        def fptr(t):
            a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
            keep.append(a)
            return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))

        if bn is None:
            bn = {"weight": torch.ones(cout), "bias": torch.zeros(cout), "running_mean": torch.zeros(cout),
                  "running_var": torch.ones(cout) - 1e-5}
        e = N.ConvBn()
        e.weight = fptr(weight)
        e.gamma, e.beta, e.mean, e.var = fptr(bn["weight"]), fptr(bn["bias"]), fptr(bn["running_mean"]), fptr(bn["running_var"])
        e.eps = 1e-5
        e.cout, e.cin, e.kh, e.kw, e.stride, e.pad = cout, cin, kh, kw, stride, pad
        ho, wo = (hin + 2 * pad - kh) // stride + 1, (win + 2 * pad - kw) // stride + 1
        out = torch.empty((n, ho, wo, cout), dtype=torch.float32, device=self.device)
        x = x_nhwc.contiguous()
        r = residual_nhwc.contiguous() if residual_nhwc is not None else None
        self._check(self._lib.fx_debug_conv(self._h, ctypes.byref(e), hin, win, x.data_ptr(), r.data_ptr() if r is not None else None,
                                            n, int(relu) | (2 if f32_out else 0), out.data_ptr(), self._stream()))
        return out

    def debug_stem_pool(self, weight: torch.Tensor, bn: Dict[str, torch.Tensor], x_nhwc: torch.Tensor) -> torch.Tensor:
        """conv1+bn1+relu+maxpool through the fused stem kernel: fp32 NHWC [n,224,224,3] -> [n,56,56,64]."""
        n = int(x_nhwc.shape[0])
        keep = []

        def fptr(t):
            a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
            keep.append(a)
            return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))

        e = N.ConvBn()
        e.weight = fptr(weight)
        e.gamma, e.beta, e.mean, e.var = fptr(bn["weight"]), fptr(bn["bias"]), fptr(bn["running_mean"]), fptr(bn["running_var"])
        e.eps = 1e-5
        e.cout, e.cin, e.kh, e.kw, e.stride, e.pad = 64, 3, 7, 7, 2, 3
        out = torch.empty((n, 56, 56, 64), dtype=torch.float32, device=self.device)
        x = x_nhwc.contiguous()
        self._check(self._lib.fx_debug_stem_pool(self._h, ctypes.byref(e), x.data_ptr(), n, out.data_ptr(), self._stream()))
        return out

    def mma_rate(self, n_cols: int, rowb: int, shift_rows: int = 0, tap_stride_rows: int = 0, iters: int = 2000) -> torch.Tensor:
        """SM cycles per tcgen05.mma (M128 x n_cols x K16) on every SM; see fx_debug_mma_rate."""
        out = torch.zeros(256, dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        self._check(self._lib.fx_debug_mma_rate(self._h, n_cols, rowb, shift_rows, tap_stride_rows, iters, out.data_ptr()))
        return out

    def tma_probe(self, base: torch.Tensor, dims, strides_bytes, box, elem_strides, swizzle: int, coords, nbytes: int) -> torch.Tensor:
        out = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        a = (ctypes.c_uint64 * 4)(*dims)
        s = (ctypes.c_uint64 * 3)(*strides_bytes)
        b = (ctypes.c_uint32 * 4)(*box)
        es = (ctypes.c_uint32 * 4)(*elem_strides)
        c = (ctypes.c_int * 4)(*coords)
        torch.cuda.synchronize(self.device)
        self._check(self._lib.fx_debug_tma_probe(self._h, base.data_ptr(), a, s, b, es, swizzle, c, nbytes, out.data_ptr()))
        return out

    def umma_shift(self, a: torch.Tensor, b: torch.Tensor, shift_rows: int, base_offset: int = 0) -> torch.Tensor:
        """a: bf16 [256,kb] CUDA, b: bf16 [64,kb] CUDA -> fp32 [128,64] = a[shift:shift+128] @ b.T on the tensor cores."""
        out = torch.empty((128, 64), dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        self._check(self._lib.fx_debug_umma_shift(self._h, a.data_ptr(), b.data_ptr(), int(a.shape[1]), shift_rows, base_offset, out.data_ptr()))
        return out

    def _dev_u8(self, t: torch.Tensor) -> None:
        if t.dtype != torch.uint8 or not t.is_cuda or not t.is_contiguous() or t.device.index != self.device_index:
            raise TypeError(f"expected a contiguous uint8 tensor on cuda:{self.device_index}")


def host_resized_size(h: int, w: int) -> Tuple[int, int]:
    oh, ow = ctypes.c_int(), ctypes.c_int()
    N.check(N.lib().fx_host_resized_size(h, w, ctypes.byref(oh), ctypes.byref(ow)))
    return oh.value, ow.value


def host_crop_offset(size: int) -> int:
    return int(N.lib().fx_host_crop_offset(size))


def host_coeffs(in_size: int, out_size: int):
    """Pillow's fixed-point coefficient table as the library computes it (host code, no GPU)."""
    L = N.lib()
    ks = L.fx_host_coeffs(in_size, out_size, None, None, None, 0)
    if ks < 0:
        raise N.FxError(ks, "fx_host_coeffs")
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    taps = np.zeros((out_size, ks), np.int32)
    rc = L.fx_host_coeffs(in_size, out_size, xmin.ctypes.data, cnt.ctypes.data, taps.ctypes.data, taps.size)
    if rc < 0:
        raise N.FxError(rc, "fx_host_coeffs")
    return xmin, cnt, taps
