// xgb_runtime.cpp -- implementation of include/xgrid_b200.h.
//
// Host-side runtime of the B200 backend: device memory, streams/events, NVRTC
// JIT (CUDA C -> sm_100a cubin), module/function handles, launches, CUDA graph
// capture/replay and NCCL halo exchange.  Links the static CUDA
// runtime only; the driver API (cuModule*, cuLaunchKernelEx),
// NVRTC and NCCL are resolved at run time so that the
// library loads on a machine without a GPU (symbol / ABI checks) and fails
// loudly -- never silently -- when a GPU call is made there.
#include "../../include/xgrid_b200.h"
#include "xgb_internal.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::mutex g_mu;
int g_device = -1;
cudaStream_t g_stream0 = nullptr;

int fail(const std::string &msg) {
    g_err = msg;
    return 1;
}

#define XGB_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            return fail(std::string(#expr) + ": " + cudaGetErrorName(e_) + " (" +        \
                        cudaGetErrorString(e_) + ")");                                   \
        }                                                                                \
    } while (0)

// ---- driver API entry points, resolved through the runtime ------------------
struct Driver {
    bool ready = false;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*FuncGetAttribute)(int *, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*LaunchKernelEx)(const CUlaunchConfig *, CUfunction, void **, void **) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t) = nullptr;
} drv;

template <class F>
int load_entry(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr)
        return fail(std::string("driver entry point '") + name + "' unavailable: " +
                    cudaGetErrorString(e));
    fn = reinterpret_cast<F>(p);
    return 0;
}

int load_driver() {
    if (drv.ready) return 0;
    if (load_entry("cuGetErrorString", drv.GetErrorString)) return 1;
    if (load_entry("cuModuleLoadData", drv.ModuleLoadData)) return 1;
    if (load_entry("cuModuleUnload", drv.ModuleUnload)) return 1;
    if (load_entry("cuModuleGetFunction", drv.ModuleGetFunction)) return 1;
    if (load_entry("cuFuncGetAttribute", drv.FuncGetAttribute)) return 1;
    if (load_entry("cuFuncSetAttribute", drv.FuncSetAttribute)) return 1;
    if (load_entry("cuLaunchKernelEx", drv.LaunchKernelEx)) return 1;
    if (load_entry("cuOccupancyMaxActiveBlocksPerMultiprocessor",
                   drv.OccupancyMaxActiveBlocksPerMultiprocessor)) return 1;
    drv.ready = true;
    return 0;
}

std::string cu_text(CUresult r) {
    const char *s = nullptr;
    if (drv.GetErrorString) drv.GetErrorString(r, &s);
    return s ? std::string(s) : ("CUresult " + std::to_string((int)r));
}

#define XGB_CU(expr)                                                                     \
    do {                                                                                 \
        CUresult r_ = (expr);                                                            \
        if (r_ != CUDA_SUCCESS) return fail(std::string(#expr) + ": " + cu_text(r_));    \
    } while (0)

inline cudaStream_t as_stream(xgb_handle h) {
    return h == 0 ? g_stream0 : reinterpret_cast<cudaStream_t>(h);
}

int require_init() {
    if (g_device < 0) return fail("xgb_init has not been called");
    return 0;
}

}  // namespace

namespace xgb_internal {
int fail(const char *msg) { return ::fail(msg); }
int require_init() { return ::require_init(); }
cudaStream_t stream_of(xgb_handle h) { return ::as_stream(h); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace xgb_internal

namespace {

// ---- NVRTC, resolved with dlopen -------------------------------------------
struct Nvrtc {
    void *lib = nullptr;
    int (*CreateProgram)(void **, const char *, const char *, int, const char *const *,
                         const char *const *) = nullptr;
    int (*DestroyProgram)(void **) = nullptr;
    int (*CompileProgram)(void *, int, const char *const *) = nullptr;
    int (*GetProgramLogSize)(void *, size_t *) = nullptr;
    int (*GetProgramLog)(void *, char *) = nullptr;
    int (*GetCUBINSize)(void *, size_t *) = nullptr;
    int (*GetCUBIN)(void *, char *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
} rtc;

int load_nvrtc() {
    if (rtc.lib) return 0;
    std::vector<std::string> tries;
    if (const char *env = getenv("XGB_NVRTC")) tries.push_back(env);
    tries.push_back("libnvrtc.so.12");
    tries.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    tries.push_back("/usr/local/cuda/lib64/libnvrtc.so");
    std::string why;
    for (auto &t : tries) {
        rtc.lib = dlopen(t.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (rtc.lib) break;
        why += std::string(dlerror()) + "; ";
    }
    if (!rtc.lib) return fail("cannot load NVRTC: " + why);
#define RTC_SYM(field, sym)                                                              \
    rtc.field = reinterpret_cast<decltype(rtc.field)>(dlsym(rtc.lib, sym));              \
    if (!rtc.field) return fail(std::string("NVRTC symbol missing: ") + sym)
    RTC_SYM(CreateProgram, "nvrtcCreateProgram");
    RTC_SYM(DestroyProgram, "nvrtcDestroyProgram");
    RTC_SYM(CompileProgram, "nvrtcCompileProgram");
    RTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    RTC_SYM(GetProgramLog, "nvrtcGetProgramLog");
    RTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    RTC_SYM(GetCUBIN, "nvrtcGetCUBIN");
    RTC_SYM(GetErrorString, "nvrtcGetErrorString");
#undef RTC_SYM
    return 0;
}

// ---- NCCL, resolved with dlopen ---------------------------------------------
struct NcclId { char internal[128]; };
struct Nccl {
    void *lib = nullptr;
    void *comm = nullptr;
    int rank = 0, n_ranks = 1;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
} nccl;

#define XGB_NCCL(expr)                                                                   \
    do {                                                                                 \
        int r_ = (expr);                                                                 \
        if (r_ != 0)                                                                     \
            return fail(std::string(#expr) + ": " +                                      \
                        (nccl.GetErrorString ? nccl.GetErrorString(r_) : "nccl error")); \
    } while (0)

}  // namespace

// ---- staged copies: pageable host memory <-> device ---------------------------------
// cudaMemcpy from / to pageable memory is bounded by ONE host thread copying through the driver's
// staging buffer (measured on the B200 box: ~11 GB/s up, ~5 GB/s down into a fresh array, against
// ~33 GB/s for page-locked mirrors).  Here kStageLanes host threads each own two page-locked chunks
// and a stream: a lane copies its chunks pageable -> pinned (or back) while the copy engine moves the
// lane's other chunk.  Correct on hardware (tests/test_staged_copy_gpu.py); opt-in from Python
// (XGB_STAGED_COPY=1) until it has been timed against the driver's pageable path.
namespace {
constexpr int kStageLanesMax = 16;
// bytes per staging chunk: XGB_STAGE_CHUNK_MB, default 4
size_t stage_chunk() {
    static size_t n = [] {
        const char *e = getenv("XGB_STAGE_CHUNK_MB");
        long v = e ? atol(e) : 4;
        return size_t(v < 1 ? 1 : (v > 64 ? 64 : v)) << 20;
    }();
    return n;
}
#define kStageChunk (stage_chunk())
// lanes in use: XGB_STAGE_LANES, default 8 (host memcpy of one thread is ~10 GB/s; PCIe 5 x16 moves ~50)
int stage_lanes() {
    static int n = [] {
        const char *e = getenv("XGB_STAGE_LANES");
        int v = e ? atoi(e) : 8;
        unsigned hw = std::thread::hardware_concurrency();
        if (hw && v > (int)hw) v = (int)hw;
        return v < 1 ? 1 : (v > kStageLanesMax ? kStageLanesMax : v);
    }();
    return n;
}
struct StageLane {
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr;
    cudaStream_t stream = nullptr;
};
StageLane g_lanes[kStageLanesMax];
cudaEvent_t g_stage_start = nullptr;
std::mutex g_stage_mu;  // one staged copy at a time: the lanes are shared

int stage_init() {
    if (!g_stage_start) XGB_CUDA(cudaEventCreateWithFlags(&g_stage_start, cudaEventDisableTiming));
    for (int i = 0; i < stage_lanes(); ++i) {
        StageLane &l = g_lanes[i];
        if (l.stream) continue;
        for (int b = 0; b < 2; ++b) {
            XGB_CUDA(cudaHostAlloc(&l.buf[b], kStageChunk, cudaHostAllocDefault));
            XGB_CUDA(cudaEventCreateWithFlags(&l.ev[b], cudaEventDisableTiming));
        }
        XGB_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        XGB_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    }
    return 0;
}

// Runs fn(lane index) on stage_lanes() threads bound to the runtime's device; returns the first CUDA error.
template <class F>
cudaError_t run_lanes(F fn) {
    std::atomic<int> err{(int)cudaSuccess};
    std::vector<std::thread> pool;
    const int kStageLanes = stage_lanes();
    for (int t = 0; t < kStageLanes; ++t)
        pool.emplace_back([&, t] {
            cudaError_t e = cudaSetDevice(g_device);
            if (e == cudaSuccess) e = fn(t);
            if (e != cudaSuccess) err.store((int)e);
        });
    for (auto &th : pool) th.join();
    return (cudaError_t)err.load();
}
}  // namespace

extern "C" {

int xgb_abi_version(void) { return XGB_ABI_VERSION; }

const char *xgb_last_error(void) { return g_err.c_str(); }

int xgb_init(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    int n = 0;
    XGB_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n)
        return fail("xgb_init: device " + std::to_string(device) + " out of range (" +
                    std::to_string(n) + " visible)");
    XGB_CUDA(cudaSetDevice(device));
    XGB_CUDA(cudaFree(nullptr));
    if (load_driver()) return 1;
    if (g_device != device || g_stream0 == nullptr) {
        g_device = device;
        XGB_CUDA(cudaStreamCreateWithFlags(&g_stream0, cudaStreamNonBlocking));
    }
    return 0;
}

int xgb_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_device < 0) return 0;
    cudaDeviceSynchronize();
    if (g_stream0) cudaStreamDestroy(g_stream0);
    g_stream0 = nullptr;
    g_device = -1;
    return 0;
}

int xgb_device_count(int *count) {
    XGB_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int xgb_get_device_info(xgb_device_info *out) {
    if (require_init()) return 1;
    cudaDeviceProp p;
    XGB_CUDA(cudaGetDeviceProperties(&p, g_device));
    memset(out, 0, sizeof(*out));
    out->ordinal = g_device;
    out->sm_count = p.multiProcessorCount;
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->max_smem_per_block_optin = (int32_t)p.sharedMemPerBlockOptin;
    out->l2_bytes = p.l2CacheSize;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_device);
    out->clock_khz = khz;
    out->total_mem = p.totalGlobalMem;
    strncpy(out->name, p.name, sizeof(out->name) - 1);
    return 0;
}

int xgb_device_sync(void) {
    if (require_init()) return 1;
    XGB_CUDA(cudaDeviceSynchronize());
    return 0;
}

// ---- memory -------------------------------------------------------------------
int xgb_alloc(size_t bytes, void **dptr) {
    if (require_init()) return 1;
    if (bytes == 0) bytes = 256;
    XGB_CUDA(cudaMalloc(dptr, bytes));
    XGB_CUDA(cudaMemsetAsync(*dptr, 0, bytes, g_stream0));
    return 0;
}

int xgb_free(void *dptr) {
    if (g_device < 0) return 0;  // after shutdown: the context owns it
    XGB_CUDA(cudaFree(dptr));
    return 0;
}

int xgb_memset(void *dptr, int byte, size_t bytes, xgb_handle stream) {
    if (require_init()) return 1;
    XGB_CUDA(cudaMemsetAsync(dptr, byte, bytes, as_stream(stream)));
    return 0;
}

int xgb_h2d(void *dst, const void *src, size_t bytes, xgb_handle stream) {
    if (require_init()) return 1;
    XGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return 0;
}

int xgb_d2h(void *dst, const void *src, size_t bytes, xgb_handle stream) {
    if (require_init()) return 1;
    XGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return 0;
}

int xgb_d2d(void *dst, const void *src, size_t bytes, xgb_handle stream) {
    if (require_init()) return 1;
    XGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return 0;
}

int xgb_host_alloc(size_t bytes, void **hptr) {
    if (require_init()) return 1;
    XGB_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

int xgb_host_free(void *hptr) {
    if (g_device < 0) return 0;
    XGB_CUDA(cudaFreeHost(hptr));
    return 0;
}

int xgb_host_register(void *hptr, size_t bytes) {
    if (require_init()) return 1;
    XGB_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
    return 0;
}

int xgb_host_unregister(void *hptr) {
    if (g_device < 0) return 0;
    XGB_CUDA(cudaHostUnregister(hptr));
    return 0;
}

int xgb_mem_info(uint64_t *free_bytes, uint64_t *total_bytes) {
    if (require_init()) return 1;
    size_t f = 0, t = 0;
    XGB_CUDA(cudaMemGetInfo(&f, &t));
    *free_bytes = f;
    *total_bytes = t;
    return 0;
}

// ---- staged copies: pageable host memory <-> device (helpers above extern "C") ----------
int xgb_h2d_staged(void *dst_dev, const void *src_host, size_t bytes, xgb_handle stream) {
    if (require_init()) return 1;
    std::lock_guard<std::mutex> lock(g_stage_mu);
    if (stage_init()) return 1;
    cudaStream_t s = as_stream(stream);
    XGB_CUDA(cudaEventRecord(g_stage_start, s));  // the lanes start after what `s` already holds
    const int kStageLanes = stage_lanes();
    for (int i = 0; i < kStageLanes; ++i) XGB_CUDA(cudaStreamWaitEvent(g_lanes[i].stream, g_stage_start, 0));
    const size_t chunks = (bytes + kStageChunk - 1) / kStageChunk;
    char *dst = static_cast<char *>(dst_dev);
    const char *src = static_cast<const char *>(src_host);
    cudaError_t e = run_lanes([&](int t) -> cudaError_t {
        StageLane &l = g_lanes[t];
        int b = 0;
        for (size_t c = (size_t)t; c < chunks; c += kStageLanes, b ^= 1) {
            const size_t off = c * kStageChunk, n = (off + kStageChunk <= bytes) ? kStageChunk : bytes - off;
            cudaError_t r = cudaEventSynchronize(l.ev[b]);  // the chunk's previous transfer has left it
            if (r != cudaSuccess) return r;
            memcpy(l.buf[b], src + off, n);
            r = cudaMemcpyAsync(dst + off, l.buf[b], n, cudaMemcpyHostToDevice, l.stream);
            if (r != cudaSuccess) return r;
            r = cudaEventRecord(l.ev[b], l.stream);
            if (r != cudaSuccess) return r;
        }
        return cudaEventRecord(l.done, l.stream);
    });
    if (e != cudaSuccess) return fail(std::string("xgb_h2d_staged: ") + cudaGetErrorString(e));
    // the source has been read completely; `s` continues once every lane's last chunk has landed
    for (int i = 0; i < kStageLanes; ++i) XGB_CUDA(cudaStreamWaitEvent(s, g_lanes[i].done, 0));
    return 0;
}

int xgb_d2h_staged(void *dst_host, const void *src_dev, size_t bytes, xgb_handle stream) {
    if (require_init()) return 1;
    std::lock_guard<std::mutex> lock(g_stage_mu);
    if (stage_init()) return 1;
    cudaStream_t s = as_stream(stream);
    XGB_CUDA(cudaEventRecord(g_stage_start, s));
    const int kStageLanes = stage_lanes();
    for (int i = 0; i < kStageLanes; ++i) XGB_CUDA(cudaStreamWaitEvent(g_lanes[i].stream, g_stage_start, 0));
    const size_t chunks = (bytes + kStageChunk - 1) / kStageChunk;
    char *dst = static_cast<char *>(dst_host);
    const char *src = static_cast<const char *>(src_dev);
    cudaError_t e = run_lanes([&](int t) -> cudaError_t {
        StageLane &l = g_lanes[t];
        auto span = [&](size_t c, size_t &off, size_t &n) {
            off = c * kStageChunk;
            n = (off + kStageChunk <= bytes) ? kStageChunk : bytes - off;
        };
        auto issue = [&](size_t c, int b) -> cudaError_t {
            size_t off, n;
            span(c, off, n);
            cudaError_t r = cudaMemcpyAsync(l.buf[b], src + off, n, cudaMemcpyDeviceToHost, l.stream);
            return r != cudaSuccess ? r : cudaEventRecord(l.ev[b], l.stream);
        };
        int b = 0;
        size_t c = (size_t)t;
        if (c < chunks) {
            cudaError_t r = issue(c, b);
            if (r != cudaSuccess) return r;
        }
        for (; c < chunks; c += kStageLanes, b ^= 1) {
            if (c + kStageLanes < chunks) {  // next chunk into the other buffer while this one is unpacked
                cudaError_t r = issue(c + kStageLanes, b ^ 1);
                if (r != cudaSuccess) return r;
            }
            cudaError_t r = cudaEventSynchronize(l.ev[b]);
            if (r != cudaSuccess) return r;
            size_t off, n;
            span(c, off, n);
            memcpy(dst + off, l.buf[b], n);
        }
        return cudaSuccess;
    });
    if (e != cudaSuccess) return fail(std::string("xgb_d2h_staged: ") + cudaGetErrorString(e));
    return 0;  // synchronous: dst_host is complete
}

// ---- host half of the mask upload: int32 boundary -> packed bytes + chunk flags + histogram ----
int xgb_mask_pack(const int32_t *src, size_t n, size_t n_padded, uint8_t *dst, uint8_t *flags,
                  uint64_t *hist_256, int *bad, int threads) {
    if (n_padded % 128 || n > n_padded) return fail("xgb_mask_pack: n_padded must be a multiple of 128 and >= n");
    const size_t chunks = n_padded / 128;
    if (threads <= 0) {
        unsigned hw = std::thread::hardware_concurrency();
        threads = hw ? (int)(hw > 16 ? 16 : hw) : 4;
    }
    if (chunks < (size_t(1) << 13)) threads = 1;          // under a million points a thread start costs more
    if ((size_t)threads > chunks) threads = chunks ? (int)chunks : 1;
    std::vector<std::vector<uint64_t>> hists(threads, std::vector<uint64_t>(256, 0));
    std::atomic<int> any_bad{0};
    auto work = [&](int t) {
        uint64_t *h = hists[t].data();
        const size_t c0 = chunks * t / threads, c1 = chunks * (t + 1) / threads;
        for (size_t c = c0; c < c1; ++c) {
            const size_t i0 = c * 128;
            const size_t m = i0 >= n ? 0 : (n - i0 < 128 ? n - i0 : 128);
            const int32_t *s = src + i0;
            uint8_t *d = dst + i0;
            int32_t acc = 0;
            for (size_t k = 0; k < m; ++k) acc |= s[k];
            if (acc == 0) {                               // the common chunk: interior, all zero
                memset(d, 0, 128);
                flags[c] = 0;
                h[0] += m;
                continue;
            }
            // pack (branch-free, vectorisable), then count only the non-zero bytes, eight at a time
            alignas(8) uint8_t tmp[128];
            for (size_t k = 0; k < m; ++k) {
                const uint32_t v = (uint32_t)s[k];
                tmp[k] = (uint8_t)(v > 254u ? 255u : v);
            }
            if (m < 128) memset(tmp + m, 0, 128 - m);
            size_t nonzero = 0;
            for (size_t w = 0; w < 16; ++w) {
                uint64_t word;
                memcpy(&word, tmp + 8 * w, 8);
                if (!word) continue;
                for (size_t k = 8 * w; k < 8 * w + 8; ++k)
                    if (tmp[k]) { ++h[tmp[k]]; ++nonzero; }
            }
            if (h[255]) any_bad.store(1, std::memory_order_relaxed);
            h[0] += m - nonzero;
            memcpy(d, tmp, 128);
            flags[c] = 1;
        }
    };
    if (threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    for (int v = 0; v < 256; ++v) {
        uint64_t sum = 0;
        for (int t = 0; t < threads; ++t) sum += hists[t][v];
        hist_256[v] = sum;
    }
    *bad = any_bad.load();
    return 0;
}

// ---- streams / events ---------------------------------------------------------
int xgb_stream_create(xgb_handle *stream) {
    if (require_init()) return 1;
    cudaStream_t s;
    XGB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = reinterpret_cast<xgb_handle>(s);
    return 0;
}

int xgb_stream_create_ex(xgb_handle *stream, int high_priority) {
    if (require_init()) return 1;
    int lo = 0, hi = 0;
    XGB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s;
    XGB_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
    *stream = reinterpret_cast<xgb_handle>(s);
    return 0;
}

int xgb_stream_destroy(xgb_handle stream) {
    if (g_device < 0 || stream == 0) return 0;
    XGB_CUDA(cudaStreamDestroy(as_stream(stream)));
    return 0;
}

int xgb_stream_sync(xgb_handle stream) {
    if (require_init()) return 1;
    XGB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

int xgb_stream_raw(xgb_handle stream, void **cuda_stream) {
    if (require_init()) return 1;
    *cuda_stream = as_stream(stream);
    return 0;
}

int xgb_event_create(xgb_handle *event) {
    if (require_init()) return 1;
    cudaEvent_t e;
    XGB_CUDA(cudaEventCreate(&e));
    *event = reinterpret_cast<xgb_handle>(e);
    return 0;
}

int xgb_event_destroy(xgb_handle event) {
    if (g_device < 0) return 0;
    XGB_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
    return 0;
}

int xgb_event_record(xgb_handle event, xgb_handle stream) {
    if (require_init()) return 1;
    XGB_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), as_stream(stream)));
    return 0;
}

int xgb_event_sync(xgb_handle event) {
    if (require_init()) return 1;
    XGB_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
    return 0;
}

int xgb_event_elapsed_ms(xgb_handle start, xgb_handle stop, float *ms) {
    if (require_init()) return 1;
    XGB_CUDA(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start),
                                  reinterpret_cast<cudaEvent_t>(stop)));
    return 0;
}

int xgb_stream_wait_event(xgb_handle stream, xgb_handle event) {
    if (require_init()) return 1;
    XGB_CUDA(cudaStreamWaitEvent(as_stream(stream), reinterpret_cast<cudaEvent_t>(event), 0));
    return 0;
}

// ---- JIT ------------------------------------------------------------------------
int xgb_compile(const char *source, const char *name, const char *const *options, int n_options,
                const char *const *header_names, const char *const *header_sources, int n_headers,
                void **image, size_t *image_bytes, char **log) {
    if (log) *log = nullptr;
    *image = nullptr;
    *image_bytes = 0;
    if (load_nvrtc()) return 1;
    void *prog = nullptr;
    int r = rtc.CreateProgram(&prog, source, name ? name : "xgrid_kernel.cu", n_headers,
                              header_sources, header_names);
    if (r != 0) return fail(std::string("nvrtcCreateProgram: ") + rtc.GetErrorString(r));
    int rc = rtc.CompileProgram(prog, n_options, options);
    size_t log_size = 0;
    rtc.GetProgramLogSize(prog, &log_size);
    std::string text(log_size ? log_size : 1, '\0');
    if (log_size) rtc.GetProgramLog(prog, &text[0]);
    if (log) {
        *log = static_cast<char *>(malloc(text.size() + 1));
        memcpy(*log, text.c_str(), text.size());
        (*log)[text.size()] = '\0';
    }
    if (rc != 0) {
        rtc.DestroyProgram(&prog);
        return fail(std::string("nvrtcCompileProgram: ") + rtc.GetErrorString(rc) + "\n" + text.c_str());
    }
    size_t n = 0;
    r = rtc.GetCUBINSize(prog, &n);
    if (r != 0 || n == 0) {
        rtc.DestroyProgram(&prog);
        return fail("nvrtcGetCUBINSize failed (was a real sm_ architecture requested?)");
    }
    *image = malloc(n);
    r = rtc.GetCUBIN(prog, static_cast<char *>(*image));
    rtc.DestroyProgram(&prog);
    if (r != 0) {
        free(*image);
        *image = nullptr;
        return fail(std::string("nvrtcGetCUBIN: ") + rtc.GetErrorString(r));
    }
    *image_bytes = n;
    return 0;
}

int xgb_release(void *p) {
    free(p);
    return 0;
}

int xgb_module_load(const void *image, size_t image_bytes, xgb_handle *module) {
    (void)image_bytes;
    if (require_init()) return 1;
    CUmodule m;
    XGB_CU(drv.ModuleLoadData(&m, image));
    *module = reinterpret_cast<xgb_handle>(m);
    return 0;
}

int xgb_module_unload(xgb_handle module) {
    if (g_device < 0) return 0;
    XGB_CU(drv.ModuleUnload(reinterpret_cast<CUmodule>(module)));
    return 0;
}

int xgb_get_function(xgb_handle module, const char *name, xgb_handle *function) {
    if (require_init()) return 1;
    CUfunction f;
    XGB_CU(drv.ModuleGetFunction(&f, reinterpret_cast<CUmodule>(module), name));
    *function = reinterpret_cast<xgb_handle>(f);
    return 0;
}

int xgb_function_info(xgb_handle function, int *regs, int *static_smem, int *local_bytes,
                      int *max_threads) {
    if (require_init()) return 1;
    CUfunction f = reinterpret_cast<CUfunction>(function);
    XGB_CU(drv.FuncGetAttribute(regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f));
    XGB_CU(drv.FuncGetAttribute(static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, f));
    XGB_CU(drv.FuncGetAttribute(local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f));
    XGB_CU(drv.FuncGetAttribute(max_threads, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, f));
    return 0;
}

int xgb_function_set_dynamic_smem(xgb_handle function, int bytes) {
    if (require_init()) return 1;
    XGB_CU(drv.FuncSetAttribute(reinterpret_cast<CUfunction>(function),
                                CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, bytes));
    return 0;
}

int xgb_occupancy(xgb_handle function, int block_threads, int dynamic_smem, int *blocks_per_sm) {
    if (require_init()) return 1;
    XGB_CU(drv.OccupancyMaxActiveBlocksPerMultiprocessor(
        blocks_per_sm, reinterpret_cast<CUfunction>(function), block_threads, (size_t)dynamic_smem));
    return 0;
}

// ---- launch ---------------------------------------------------------------------
static int launch_impl(xgb_handle function, const uint32_t grid[3], const uint32_t block[3],
                       uint32_t dynamic_smem, xgb_handle stream, const void *params) {
    if (require_init()) return 1;
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDimX = grid[0];
    cfg.gridDimY = grid[1];
    cfg.gridDimZ = grid[2];
    cfg.blockDimX = block[0];
    cfg.blockDimY = block[1];
    cfg.blockDimZ = block[2];
    cfg.sharedMemBytes = dynamic_smem;
    cfg.hStream = reinterpret_cast<CUstream>(as_stream(stream));
    void *args[1] = {const_cast<void *>(params)};
    XGB_CU(drv.LaunchKernelEx(&cfg, reinterpret_cast<CUfunction>(function), args, nullptr));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int xgb_launch(xgb_handle function, const uint32_t grid[3], const uint32_t block[3],
               uint32_t dynamic_smem, xgb_handle stream, const void *params, size_t param_bytes) {
    (void)param_bytes;
    return launch_impl(function, grid, block, dynamic_smem, stream, params);
}

int xgb_launch_count(uint64_t *count) {
    *count = g_launches.load(std::memory_order_relaxed);
    return 0;
}

// ---- graphs ---------------------------------------------------------------------
struct GraphExec {
    cudaGraphExec_t exec;
    int kernel_nodes;
};

int xgb_graph_begin(xgb_handle stream) {
    if (require_init()) return 1;
    XGB_CUDA(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
    return 0;
}

int xgb_graph_end(xgb_handle stream, xgb_handle *graph_exec, int *kernel_nodes) {
    if (require_init()) return 1;
    cudaGraph_t graph = nullptr;
    XGB_CUDA(cudaStreamEndCapture(as_stream(stream), &graph));
    size_t n = 0;
    XGB_CUDA(cudaGraphGetNodes(graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    int kernels = 0;
    if (n) {
        XGB_CUDA(cudaGraphGetNodes(graph, nodes.data(), &n));
        for (size_t i = 0; i < n; ++i) {
            cudaGraphNodeType t;
            XGB_CUDA(cudaGraphNodeGetType(nodes[i], &t));
            if (t == cudaGraphNodeTypeKernel) ++kernels;
        }
    }
    cudaGraphExec_t exec = nullptr;
    // per-node priorities: a halo exchange captured from the high-priority communication stream must still overtake
    // the interior sweep it overlaps with when the recorded call is replayed
    cudaError_t e = cudaGraphInstantiate(&exec, graph, cudaGraphInstantiateFlagUseNodePriority);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess)
        return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    // the capture-time launches were recorded, not executed
    g_launches.fetch_sub((uint64_t)kernels, std::memory_order_relaxed);
    auto *g = new GraphExec{exec, kernels};
    *graph_exec = reinterpret_cast<xgb_handle>(g);
    if (kernel_nodes) *kernel_nodes = kernels;
    return 0;
}

int xgb_graph_launch(xgb_handle graph_exec, xgb_handle stream) {
    if (require_init()) return 1;
    auto *g = reinterpret_cast<GraphExec *>(graph_exec);
    XGB_CUDA(cudaGraphLaunch(g->exec, as_stream(stream)));
    g_launches.fetch_add((uint64_t)g->kernel_nodes, std::memory_order_relaxed);
    return 0;
}

int xgb_graph_destroy(xgb_handle graph_exec) {
    auto *g = reinterpret_cast<GraphExec *>(graph_exec);
    if (!g) return 0;
    if (g_device >= 0) cudaGraphExecDestroy(g->exec);
    delete g;
    return 0;
}

// ---- NCCL halo exchange -----------------------------------------------------------
int xgb_nccl_load(const char *path) {
    if (nccl.lib) return 0;
    std::vector<std::string> tries;
    if (path && *path) tries.push_back(path);
    tries.push_back("libnccl.so.2");
    std::string why;
    for (auto &t : tries) {
        nccl.lib = dlopen(t.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (nccl.lib) break;
        why += std::string(dlerror()) + "; ";
    }
    if (!nccl.lib) return fail("cannot load NCCL: " + why);
#define NCCL_SYM(field, sym)                                                             \
    nccl.field = reinterpret_cast<decltype(nccl.field)>(dlsym(nccl.lib, sym));           \
    if (!nccl.field) return fail(std::string("NCCL symbol missing: ") + sym)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(GroupStart, "ncclGroupStart");
    NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(Send, "ncclSend");
    NCCL_SYM(Recv, "ncclRecv");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    return 0;
}

int xgb_nccl_unique_id(void *id_128B) {
    if (!nccl.lib && xgb_nccl_load(nullptr)) return 1;
    NcclId id;
    XGB_NCCL(nccl.GetUniqueId(&id));
    memcpy(id_128B, &id, sizeof(id));
    return 0;
}

int xgb_nccl_init(const void *id_128B, int rank, int n_ranks) {
    if (require_init()) return 1;
    if (!nccl.lib && xgb_nccl_load(nullptr)) return 1;
    NcclId id;
    memcpy(&id, id_128B, sizeof(id));
    XGB_NCCL(nccl.CommInitRank(&nccl.comm, n_ranks, id, rank));
    nccl.rank = rank;
    nccl.n_ranks = n_ranks;
    return 0;
}

int xgb_nccl_shutdown(void) {
    if (nccl.comm) {
        nccl.CommDestroy(nccl.comm);
        nccl.comm = nullptr;
    }
    return 0;
}

int xgb_halo_exchange(const xgb_halo_desc *descs, int n, xgb_handle stream) {
    if (require_init()) return 1;
    if (!nccl.comm) return fail("xgb_halo_exchange: xgb_nccl_init has not been called");
    cudaStream_t s = as_stream(stream);
    const int kChar = 0;  // ncclInt8
    XGB_NCCL(nccl.GroupStart());
    // Sends first (lo, hi), then receives in the order (hi, lo): when both neighbours are the SAME
    // rank (a ring of two, overstep="wrap") NCCL pairs the k-th send to a peer with the peer's k-th
    // receive from us, and the peer's upper ghost must get our first rows, its lower ghost our last.
    int rc = 0;
    for (int i = 0; i < n && rc == 0; ++i) {
        const xgb_halo_desc &d = descs[i];
        if (rc == 0 && d.lo_rank >= 0 && d.send_lo) rc = nccl.Send(d.send_lo, d.bytes, kChar, d.lo_rank, nccl.comm, s);
        if (rc == 0 && d.hi_rank >= 0 && d.send_hi) rc = nccl.Send(d.send_hi, d.bytes, kChar, d.hi_rank, nccl.comm, s);
        if (rc == 0 && d.hi_rank >= 0 && d.recv_hi) rc = nccl.Recv(d.recv_hi, d.bytes, kChar, d.hi_rank, nccl.comm, s);
        if (rc == 0 && d.lo_rank >= 0 && d.recv_lo) rc = nccl.Recv(d.recv_lo, d.bytes, kChar, d.lo_rank, nccl.comm, s);
    }
    const int end_rc = nccl.GroupEnd();      // always close the group, also after a failed call
    if (rc != 0 || end_rc != 0)
        return fail(std::string("xgb_halo_exchange: ") +
                    (nccl.GetErrorString ? nccl.GetErrorString(rc != 0 ? rc : end_rc) : "nccl error"));
    return 0;
}

}  // extern "C"
