// GPU JPEG decode feeding the fused preprocess kernel: replaces `Image.open` + Pillow's decode
// inside the reference's serial batch loop (src/feature_extraction.py:238, 276-284) for the files nvJPEG can take --
// baseline, 8-bit, three-component (YCbCr -> Pillow mode "RGB") JPEGs with a complete bitstream.  Everything else (PNG,
// progressive / CMYK / grayscale JPEGs, truncated files, non-images) is reported back as "host decode" and goes through
// the very Pillow call the reference makes, so the reference's per-file error semantics are untouched.
//
//   * fx_jpeg_read_files: a small native thread pool reads the files of one batch straight into the slot's page-locked
//     bitstream buffer and parses the frame header (SOF marker walk) -- no Python per file.
//   * fx_embed_files_async: nvjpegDecodeBatched (hardware JPEG engines when the handle could be created for them, else
//     the CUDA "GPU hybrid" decoder) writes interleaved RGB straight into the slot's device image buffer at the
//     offsets of the descriptor table, host-decoded stragglers are copied in beside them, then preprocess + trunk run
//     on the lane's stream exactly as for fx_embed_host_async.
//
// nvJPEG is loaded with dlopen at fx_jpeg_init: the library itself has no link-time dependency on it, and without it
// the entry points fail with FX_ERR_UNSUPPORTED (the Python host then keeps using its host decode pool).
// Not bit-exact against libjpeg-turbo (IDCT and chroma up-sampling differ): opt-in, tolerance measured in
// profiles/r02_nvjpeg_tolerance.md; the host decode pool stays the bit-exact default.
#include <dlfcn.h>
#include <fcntl.h>
#include <nvjpeg.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <thread>

#include "fx_common.cuh"

namespace fx {

int embed_slot_prepare(fx_engine* e, int slot, size_t total_bytes);
int embed_slot_compute(fx_engine* e, int slot, const fx_image_desc* descs, int n, float* emb_host, float* emb_dev_out);

namespace {

struct NvJpegApi {
    void* so = nullptr;
    decltype(&nvjpegCreateEx) CreateEx = nullptr;
    decltype(&nvjpegDestroy) Destroy = nullptr;
    decltype(&nvjpegJpegStateCreate) JpegStateCreate = nullptr;
    decltype(&nvjpegJpegStateDestroy) JpegStateDestroy = nullptr;
    decltype(&nvjpegDecodeBatchedInitialize) DecodeBatchedInitialize = nullptr;
    decltype(&nvjpegDecodeBatched) DecodeBatched = nullptr;
    decltype(&nvjpegJpegStreamCreate) JpegStreamCreate = nullptr;
    decltype(&nvjpegJpegStreamDestroy) JpegStreamDestroy = nullptr;
    decltype(&nvjpegJpegStreamParseHeader) JpegStreamParseHeader = nullptr;
    decltype(&nvjpegDecodeBatchedSupported) DecodeBatchedSupported = nullptr;
};

struct JpegCtx {
    NvJpegApi api;
    nvjpegHandle_t handle = nullptr;
    int backend = 0;  // FX_JPEG_BACKEND_HARDWARE or FX_JPEG_BACKEND_GPU, whichever the handle was created for
    nvjpegJpegStream_t probe_stream = nullptr;
    struct Slot {
        nvjpegJpegState_t state = nullptr;
        int init_batch = 0;
        uint8_t* bits = nullptr;  // page-locked bitstream buffer of the slot (fx_jpeg_read_files)
        size_t bits_cap = 0;
    } slots[FX_HOST_SLOTS + 1];  // the extra one serves the synchronous fx_jpeg_decode
    int io_threads = 8;
};

template <typename T>
bool load_sym(void* so, const char* name, T& fn) {
    fn = reinterpret_cast<T>(dlsym(so, name));
    return fn != nullptr;
}

const char* nvjpeg_text(nvjpegStatus_t s) {
    switch (s) {
        case NVJPEG_STATUS_SUCCESS: return "success";
        case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
        case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
        case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
        case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "jpeg not supported";
        case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
        case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
        case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
        case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
        case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "implementation not supported";
        case NVJPEG_STATUS_INCOMPLETE_BITSTREAM: return "incomplete bitstream";
    }
    return "unknown status";
}

JpegCtx* ctx_of(fx_engine* e) { return static_cast<JpegCtx*>(e->jpeg_state); }

// Frame header of a JPEG bitstream by a marker walk (ITU-T T.81 B.1.1): fills height / width / components /
// subsampling / encoding; returns false when `data` is not a JPEG whose header can be read.
bool parse_sof(const uint8_t* d, size_t len, fx_file_info& fi) {
    if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return false;
    size_t pos = 2;
    while (pos + 4 <= len) {
        if (d[pos] != 0xFF) return false;
        uint8_t m = d[pos + 1];
        if (m == 0xFF) {  // fill byte
            ++pos;
            continue;
        }
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) {  // markers without a segment
            pos += 2;
            continue;
        }
        if (m == 0xD9 || m == 0xDA) return false;  // end of image / start of scan before any frame header
        const size_t seg = ((size_t)d[pos + 2] << 8) | d[pos + 3];
        if (seg < 2 || pos + 2 + seg > len) return false;
        const bool sof = m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC;
        if (sof) {
            if (seg < 8) return false;
            const uint8_t* s = d + pos + 4;
            fi.encoding = m;
            fi.precision = s[0];
            fi.height = (s[1] << 8) | s[2];
            fi.width = (s[3] << 8) | s[4];
            fi.components = s[5];
            if (seg < 8 + 3u * fi.components) return false;
            fi.subsampling = -1;
            if (fi.components == 1) {
                fi.subsampling = NVJPEG_CSS_GRAY;
            } else if (fi.components == 3) {
                const int h0 = s[7] >> 4, v0 = s[7] & 15, h1 = s[10] >> 4, v1 = s[10] & 15, h2 = s[13] >> 4, v2 = s[13] & 15;
                if (h1 == 1 && v1 == 1 && h2 == 1 && v2 == 1) {
                    if (h0 == 1 && v0 == 1) fi.subsampling = NVJPEG_CSS_444;
                    if (h0 == 2 && v0 == 1) fi.subsampling = NVJPEG_CSS_422;
                    if (h0 == 2 && v0 == 2) fi.subsampling = NVJPEG_CSS_420;
                    if (h0 == 1 && v0 == 2) fi.subsampling = NVJPEG_CSS_440;
                    if (h0 == 4 && v0 == 1) fi.subsampling = NVJPEG_CSS_411;
                    if (h0 == 4 && v0 == 2) fi.subsampling = NVJPEG_CSS_410;
                }
            }
            return fi.height > 0 && fi.width > 0;
        }
        pos += 2 + seg;
    }
    return false;
}

// Pillow raises OSError("image file is truncated") when the entropy-coded data ends before the image is complete; a
// complete file carries the EOI marker (possibly followed by padding / trailing bytes).  Files without one within
// the last 4 KB go to the host decoder, which then reports exactly what the reference would.
bool has_eoi(const uint8_t* d, size_t len) {
    const size_t lo = len > 4096 ? len - 4096 : 0;
    for (size_t i = len; i >= lo + 2; --i)
        if (d[i - 2] == 0xFF && d[i - 1] == 0xD9) return true;
    return false;
}

int ensure_bits(fx_engine* e, JpegCtx::Slot& s, size_t need) {
    if (need <= s.bits_cap) return FX_OK;
    if (s.bits) cudaFreeHost(s.bits);
    s.bits = nullptr;
    s.bits_cap = 0;
    const size_t cap = need + need / 4 + 65536;
    cudaError_t a = cudaHostAlloc(reinterpret_cast<void**>(&s.bits), cap, cudaHostAllocDefault);
    if (a != cudaSuccess) return set_error(e, FX_ERR_NOMEM, std::string("cudaHostAlloc(bitstream buffer): ") + cudaGetErrorString(a));
    s.bits_cap = cap;
    return FX_OK;
}

constexpr uint64_t kMaxJpegFileBytes = 1ull << 28;  // larger "JPEGs" go to the host decoder (a 65535 x 65535 frame would not fit the device buffers anyway)

template <typename F>
void parallel_for(int n, int threads, F&& body) {
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) {
        for (int i = 0; i < n; ++i) body(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i);
        });
    for (auto& th : pool) th.join();
}

int batched_decode(fx_engine* e, JpegCtx* jc, JpegCtx::Slot& s, const std::vector<const unsigned char*>& ptrs, const std::vector<size_t>& lens,
                   std::vector<nvjpegImage_t>& dests, cudaStream_t stream) {
    const int n = (int)ptrs.size();
    if (n == 0) return FX_OK;
    if (s.init_batch != n) {
        // max_cpu_threads = 1: measured 4.9 ms of host time per 256 files against 7.3 ms with 8.  Cutting a batch into
        // sub-batches on their own host threads / states / streams was measured too (round 2): 31.5 k -> 26 k (2 parts) -> 17 k
        // images/s (4 parts) -- the library serialises them and re-initialises per size -- so one call per batch it stays;
        // the GPU-side Huffman decode (~8 ms per 256 files of 512x512) is what paces the file path.  NVJPEG_BACKEND_HYBRID (Huffman
        // on the host) was measured as well: its batched decode walks the files on ONE host thread whatever max_cpu_threads
        // says -- 100 ms per 256 files, 2.6 k images/s.
        nvjpegStatus_t st = jc->api.DecodeBatchedInitialize(jc->handle, s.state, n, 1, NVJPEG_OUTPUT_RGBI);
        if (st != NVJPEG_STATUS_SUCCESS) {
            s.init_batch = 0;
            return set_error(e, FX_ERR_CUDA, std::string("nvjpegDecodeBatchedInitialize: ") + nvjpeg_text(st));
        }
        s.init_batch = n;
    }
    nvjpegStatus_t st = jc->api.DecodeBatched(jc->handle, s.state, ptrs.data(), lens.data(), dests.data(), stream);
    if (st != NVJPEG_STATUS_SUCCESS) {
        s.init_batch = 0;  // nvjpegDecodeBatchedInitialize also resets a failed batch
        return set_error(e, FX_ERR_UNSUPPORTED, std::string("nvjpegDecodeBatched: ") + nvjpeg_text(st));
    }
    return FX_OK;
}

}  // namespace

void jpeg_free(fx_engine* e) {
    JpegCtx* jc = ctx_of(e);
    if (!jc) return;
    for (auto& s : jc->slots) {
        if (s.state) jc->api.JpegStateDestroy(s.state);
        if (s.bits) cudaFreeHost(s.bits);
    }
    if (jc->probe_stream) jc->api.JpegStreamDestroy(jc->probe_stream);
    if (jc->handle) jc->api.Destroy(jc->handle);
    // the shared object stays mapped: unloading a CUDA library at exit is asking for trouble
    delete jc;
    e->jpeg_state = nullptr;
}

}  // namespace fx

using namespace fx;

extern "C" {

int fx_jpeg_init(fx_handle e, int backend) {
    if (!e) return FX_ERR_INVALID;
    if (backend != FX_JPEG_BACKEND_AUTO && backend != FX_JPEG_BACKEND_HARDWARE && backend != FX_JPEG_BACKEND_GPU)
        return set_error(e, FX_ERR_INVALID, "fx_jpeg_init: unknown backend");
    if (e->jpeg_state) {
        JpegCtx* jc = ctx_of(e);
        if (backend == FX_JPEG_BACKEND_AUTO || backend == jc->backend) return FX_OK;
        FX_CUDA(e, cudaSetDevice(e->device));
        FX_CUDA(e, cudaDeviceSynchronize());
        jpeg_free(e);
    }
    FX_CUDA(e, cudaSetDevice(e->device));
    JpegCtx* jc = new JpegCtx();
    const char* names[] = {getenv("FX_NVJPEG_LIB"), "libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        jc->api.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (jc->api.so) break;
    }
    if (!jc->api.so) {
        delete jc;
        return set_error(e, FX_ERR_UNSUPPORTED, "fx_jpeg_init: libnvjpeg.so.12 could not be loaded (set FX_NVJPEG_LIB); GPU JPEG decode unavailable");
    }
    NvJpegApi& a = jc->api;
    const bool ok = load_sym(a.so, "nvjpegCreateEx", a.CreateEx) && load_sym(a.so, "nvjpegDestroy", a.Destroy) &&
                    load_sym(a.so, "nvjpegJpegStateCreate", a.JpegStateCreate) && load_sym(a.so, "nvjpegJpegStateDestroy", a.JpegStateDestroy) &&
                    load_sym(a.so, "nvjpegDecodeBatchedInitialize", a.DecodeBatchedInitialize) && load_sym(a.so, "nvjpegDecodeBatched", a.DecodeBatched) &&
                    load_sym(a.so, "nvjpegJpegStreamCreate", a.JpegStreamCreate) && load_sym(a.so, "nvjpegJpegStreamDestroy", a.JpegStreamDestroy) &&
                    load_sym(a.so, "nvjpegJpegStreamParseHeader", a.JpegStreamParseHeader) &&
                    load_sym(a.so, "nvjpegDecodeBatchedSupported", a.DecodeBatchedSupported);
    if (!ok) {
        delete jc;
        return set_error(e, FX_ERR_UNSUPPORTED, "fx_jpeg_init: libnvjpeg lacks a required entry point");
    }
    nvjpegStatus_t st = NVJPEG_STATUS_ARCH_MISMATCH;
    if (backend != FX_JPEG_BACKEND_GPU) {
        st = a.CreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &jc->handle);
        if (st == NVJPEG_STATUS_SUCCESS) jc->backend = FX_JPEG_BACKEND_HARDWARE;
    }
    if (st != NVJPEG_STATUS_SUCCESS && backend != FX_JPEG_BACKEND_HARDWARE) {
        // interpolated chroma up-sampling is what libjpeg-turbo's default ("fancy") up-sampling does
        st = a.CreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION, &jc->handle);
        if (st != NVJPEG_STATUS_SUCCESS) st = a.CreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &jc->handle);
        if (st == NVJPEG_STATUS_SUCCESS) jc->backend = FX_JPEG_BACKEND_GPU;
    }
    if (st != NVJPEG_STATUS_SUCCESS) {
        delete jc;
        cudaGetLastError();
        return set_error(e, FX_ERR_UNSUPPORTED, std::string("fx_jpeg_init: nvjpegCreateEx: ") + nvjpeg_text(st));
    }
    e->jpeg_state = jc;
    for (auto& s : jc->slots)
        if ((st = a.JpegStateCreate(jc->handle, &s.state)) != NVJPEG_STATUS_SUCCESS) break;
    if (st == NVJPEG_STATUS_SUCCESS) st = a.JpegStreamCreate(jc->handle, &jc->probe_stream);
    if (st != NVJPEG_STATUS_SUCCESS) {
        jpeg_free(e);
        return set_error(e, FX_ERR_CUDA, std::string("fx_jpeg_init: nvjpeg state: ") + nvjpeg_text(st));
    }
    if (const char* t = getenv("FX_IO_THREADS")) jc->io_threads = std::max(1, atoi(t));
    return FX_OK;
}

int fx_jpeg_backend(fx_handle e) {
    if (!e || !e->jpeg_state) return FX_JPEG_BACKEND_AUTO;
    return ctx_of(e)->backend;
}

int fx_jpeg_probe(fx_handle e, const uint8_t* data, size_t length, fx_file_info* info) {
    if (!e) return FX_ERR_INVALID;
    if (!data || !info) return set_error(e, FX_ERR_INVALID, "fx_jpeg_probe: null pointer");
    JpegCtx* jc = ctx_of(e);
    if (!jc) return set_error(e, FX_ERR_STATE, "fx_jpeg_probe: fx_jpeg_init has not succeeded");
    fx_file_info fi;
    std::memset(&fi, 0, sizeof(fi));
    fi.length = length;
    fi.status = FX_FILE_HOST_DECODE;
    if (parse_sof(data, length, fi) && fi.encoding == 0xC0 && fi.precision == 8 && fi.components == 3 && fi.subsampling >= 0 &&
        has_eoi(data, length)) {
        int unsupported = 1;
        if (jc->api.JpegStreamParseHeader(jc->handle, data, length, jc->probe_stream) == NVJPEG_STATUS_SUCCESS &&
            jc->api.DecodeBatchedSupported(jc->handle, jc->probe_stream, &unsupported) == NVJPEG_STATUS_SUCCESS && unsupported == 0)
            fi.status = FX_FILE_GPU_JPEG;
    }
    *info = fi;
    return FX_OK;
}

int fx_jpeg_read_files(fx_handle e, int slot, const char* const* paths, int n, fx_file_info* info) {
    if (!e) return FX_ERR_INVALID;
    if (slot < 0 || slot >= FX_HOST_SLOTS || n < 0 || (n > 0 && (!paths || !info))) return set_error(e, FX_ERR_INVALID, "fx_jpeg_read_files: bad arguments");
    JpegCtx* jc = ctx_of(e);
    if (!jc) return set_error(e, FX_ERR_STATE, "fx_jpeg_read_files: fx_jpeg_init has not succeeded");
    FX_CUDA(e, cudaSetDevice(e->device));
    fx_engine::HostSlot& hs = e->slots[slot];
    if (hs.busy) {  // the previous batch of this slot still reads the bitstream buffer
        FX_CUDA(e, cudaEventSynchronize(hs.done));
        hs.busy = false;
    }
    JpegCtx::Slot& s = jc->slots[slot];
    std::memset(info, 0, sizeof(fx_file_info) * (size_t)n);
    // pass 1: sizes -> layout (64-byte aligned starts).  Only files that start like a JPEG (SOI marker + the first byte of
    // the next marker) get room in the bitstream buffer: a dataset directory may hold anything (the reference lists every
    // file, src/feature_extraction.py:145-170), and a 2 GB volume or archive must not be pulled into page-locked memory just
    // to learn that Pillow cannot identify it.
    parallel_for(n, jc->io_threads, [&](int i) {
        struct stat sb;
        const int fd = paths[i] ? open(paths[i], O_RDONLY | O_CLOEXEC) : -1;
        if (fd < 0 || fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
            if (fd >= 0) close(fd);
            info[i].status = FX_FILE_UNREADABLE;
            return;
        }
        uint8_t head[3] = {0, 0, 0};
        const ssize_t r = pread(fd, head, 3, 0);
        close(fd);
        if (r == 3 && head[0] == 0xFF && head[1] == 0xD8 && head[2] == 0xFF && (uint64_t)sb.st_size <= kMaxJpegFileBytes)
            info[i].length = (uint64_t)sb.st_size;
        else
            info[i].status = FX_FILE_HOST_DECODE;  // length 0: not read here, the host decoder opens it itself
    });
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        info[i].offset = total;
        total += (info[i].length + 63) & ~(size_t)63;
    }
    int rc = ensure_bits(e, s, total + 64);
    if (rc != FX_OK) return rc;
    // pass 2: read + frame header
    parallel_for(n, jc->io_threads, [&](int i) {
        fx_file_info& fi = info[i];
        if (fi.status != FX_FILE_GPU_JPEG) return;  // (0 = still a candidate) unreadable, or left to the host decoder by pass 1
        fi.status = FX_FILE_HOST_DECODE;
        const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
        if (fd < 0) {
            fi.status = FX_FILE_UNREADABLE;
            return;
        }
        uint8_t* dst = s.bits + fi.offset;
        size_t got = 0;
        while (got < fi.length) {
            const ssize_t r = read(fd, dst + got, fi.length - got);
            if (r <= 0) break;
            got += (size_t)r;
        }
        close(fd);
        if (got != fi.length) {  // changed under our feet: let the host decoder have a look
            fi.length = got;
            return;
        }
        if (parse_sof(dst, got, fi) && fi.encoding == 0xC0 && fi.precision == 8 && fi.components == 3 && fi.subsampling >= 0 && has_eoi(dst, got))
            fi.status = FX_FILE_GPU_JPEG;
    });
    // the hardware engines take a subset of baseline JPEG (no 4:1:0 / 4:1:1, single scan): ask nvJPEG per candidate
    if (jc->backend == FX_JPEG_BACKEND_HARDWARE)
        for (int i = 0; i < n; ++i) {
            if (info[i].status != FX_FILE_GPU_JPEG) continue;
            int unsupported = 1;
            if (jc->api.JpegStreamParseHeader(jc->handle, s.bits + info[i].offset, info[i].length, jc->probe_stream) != NVJPEG_STATUS_SUCCESS ||
                jc->api.DecodeBatchedSupported(jc->handle, jc->probe_stream, &unsupported) != NVJPEG_STATUS_SUCCESS || unsupported != 0)
                info[i].status = FX_FILE_HOST_DECODE;
        }
    return FX_OK;
}

int fx_jpeg_decode(fx_handle e, const uint8_t* const* data, const size_t* lengths, int n, uint8_t* dst_dev, const fx_image_desc* descs,
                   void* stream_) {
    if (!e) return FX_ERR_INVALID;
    if (n < 0 || (n > 0 && (!data || !lengths || !dst_dev || !descs))) return set_error(e, FX_ERR_INVALID, "fx_jpeg_decode: bad arguments");
    JpegCtx* jc = ctx_of(e);
    if (!jc) return set_error(e, FX_ERR_STATE, "fx_jpeg_decode: fx_jpeg_init has not succeeded");
    if (n == 0) return FX_OK;
    FX_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    std::vector<const unsigned char*> ptrs(data, data + n);
    std::vector<size_t> lens(lengths, lengths + n);
    std::vector<nvjpegImage_t> dests(n);
    for (int i = 0; i < n; ++i) {
        if (descs[i].channels != 3) return set_error(e, FX_ERR_INVALID, "fx_jpeg_decode: the output is interleaved RGB, descriptors must say 3 channels");
        std::memset(&dests[i], 0, sizeof(nvjpegImage_t));
        dests[i].channel[0] = dst_dev + descs[i].offset;
        dests[i].pitch[0] = (size_t)descs[i].width * 3;
    }
    int rc = batched_decode(e, jc, jc->slots[FX_HOST_SLOTS], ptrs, lens, dests, stream);
    if (rc != FX_OK) return rc;
    FX_CUDA(e, cudaStreamSynchronize(stream));  // the bitstreams are the caller's and may go away once this returns
    return FX_OK;
}

int fx_embed_files_async(fx_handle e, int slot, const fx_file_info* info, const uint8_t* const* host_pixels, const fx_image_desc* descs, int n,
                         size_t total_bytes, float* emb_host, float* emb_dev) {
    if (!e) return FX_ERR_INVALID;
    if (slot < 0 || slot >= FX_HOST_SLOTS || n < 0 || n > e->max_batch || (n > 0 && (!info || !descs || (!emb_host && !emb_dev))))
        return set_error(e, FX_ERR_INVALID, "fx_embed_files_async: bad arguments");
    JpegCtx* jc = ctx_of(e);
    if (!jc) return set_error(e, FX_ERR_STATE, "fx_embed_files_async: fx_jpeg_init has not succeeded");
    JpegCtx::Slot& s = jc->slots[slot];
    for (int i = 0; i < n; ++i) {
        const size_t bytes = (size_t)descs[i].height * descs[i].width * descs[i].channels;
        if (descs[i].height < 1 || descs[i].width < 1 || descs[i].offset + bytes > total_bytes)
            return set_error(e, FX_ERR_INVALID, "fx_embed_files_async: image " + std::to_string(i) + " lies outside the buffer");
        if (info[i].status == FX_FILE_GPU_JPEG) {
            if (descs[i].channels != 3 || descs[i].height != info[i].height || descs[i].width != info[i].width || !s.bits ||
                info[i].offset + info[i].length > s.bits_cap)
                return set_error(e, FX_ERR_INVALID, "fx_embed_files_async: descriptor " + std::to_string(i) + " does not match the JPEG read into this slot");
        } else if (!host_pixels || !host_pixels[i]) {
            return set_error(e, FX_ERR_INVALID, "fx_embed_files_async: file " + std::to_string(i) + " needs host-decoded pixels");
        }
    }
    FX_CUDA(e, cudaSetDevice(e->device));
    int rc = embed_slot_prepare(e, slot, total_bytes);
    if (rc != FX_OK || n == 0) return rc;
    fx_engine::HostSlot& hs = e->slots[slot];
    std::vector<const unsigned char*> ptrs;
    std::vector<size_t> lens;
    std::vector<nvjpegImage_t> dests;
    for (int i = 0; i < n; ++i) {
        if (info[i].status == FX_FILE_GPU_JPEG) {
            nvjpegImage_t d;
            std::memset(&d, 0, sizeof(d));
            d.channel[0] = hs.src_dev + descs[i].offset;
            d.pitch[0] = (size_t)descs[i].width * 3;
            ptrs.push_back(s.bits + info[i].offset);
            lens.push_back((size_t)info[i].length);
            dests.push_back(d);
        } else {
            FX_CUDA(e, cudaMemcpyAsync(hs.src_dev + descs[i].offset, host_pixels[i], (size_t)descs[i].height * descs[i].width * descs[i].channels,
                                       cudaMemcpyHostToDevice, e->copy_stream));
        }
    }
    if ((rc = batched_decode(e, jc, s, ptrs, lens, dests, e->copy_stream)) == FX_OK) {
        cudaError_t ev = cudaEventRecord(hs.copied, e->copy_stream);
        rc = ev == cudaSuccess ? embed_slot_compute(e, slot, descs, n, emb_host, emb_dev) : set_error(e, FX_ERR_CUDA, cudaGetErrorString(ev));
    }
    if (rc != FX_OK) cudaStreamSynchronize(e->copy_stream);  // the slot is not marked busy: host pixels / bitstreams must be free of DMA when the error returns
    return rc;
}

}  // extern "C"
