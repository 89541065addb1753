// xgb_peer.cu -- halo exchange over peer memory (NVLink / NVSwitch), one kernel per exchange.
//
// Every rank owns a MAILBOX in its own HBM that both neighbours map through CUDA IPC:
//
//     header   ready[2]   written by the lo / hi neighbour: "exchange number s of mine is in your slot"
//              credit[2]  written by the lo / hi neighbour: "I have copied exchange s of yours out of my slot"
//              seq        this rank's exchange counter (advanced by the kernel itself, so a recorded CUDA graph
//                         replays correctly), done[2] CTA counters, error
//     in[side][parity][slot_bytes]    rows arriving from the lo (side 0) / hi (side 1) neighbour
//
// One launch of `xgb_peer_exchange_kernel` is one exchange s = seq + 1 of up to MAX_ITEMS levels:
//   1. wait until the neighbours have emptied the slots of exchange s - 2 (credit), then PUSH my first rows into the lo
//      neighbour's in[1][s & 1] and my last rows into the hi neighbour's in[0][s & 1] with plain stores over NVLink;
//      the last CTA to finish publishes ready = s in both neighbours' headers (st.release.sys);
//   2. wait for ready >= s from both neighbours (ld.acquire.sys), copy my in[0] / in[1] slots into my ghost rows, and
//      the last CTA to finish returns the credit and advances seq.
// Nothing but the two flags crosses a GPU boundary besides the rows themselves; there is no rendezvous with the host,
// no proxy thread and no second kernel.  All exchanges of a process run on ONE stream (dist.py), in the order the
// host issued them -- the same order on every rank -- so exchange numbers pair up by construction.  A neighbour that
// never arrives trips a timeout (XGB_PEER_TIMEOUT_S, default 120 s): the kernel reports and traps instead of hanging.
//
// New work, no reference counterpart; replaces ncclSend/ncclRecv on the path (xgb_halo_exchange
// stays as the transport for GPUs without peer access).
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "xgb_internal.h"

namespace {

constexpr int MAX_ITEMS = 8;
constexpr size_t HEADER_BYTES = 256;
// Small CTAs with few registers: the exchange must fit BESIDE the resident CTAs of the sweep it overlaps (the 3-D
// bulk-copy kernel leaves 10240 registers per SM; a 256-thread, 56-register version of this kernel did not fit, waited
// for a sweep CTA to retire and then delayed that SM's next one -- +100 us per step, profiles/r2_experiments.md).
constexpr int THREADS = 128;
constexpr int MIN_CTAS_PER_SM = 12;          // => at most 40 registers per thread

struct Header {
    uint32_t ready[2];
    uint32_t credit[2];
    uint32_t seq;
    uint32_t done[2];
    uint32_t error;
};
static_assert(sizeof(Header) <= HEADER_BYTES, "mailbox header");

struct Item {
    const char *send_lo;
    char *recv_lo;
    const char *send_hi;
    char *recv_hi;
    uint64_t bytes;
    uint64_t offset;        // of this item inside a slot
};

struct Args {
    Item items[MAX_ITEMS];
    int n;
    char *mine, *lo, *hi;   // mailboxes: this rank's, the lo / hi neighbour's (mapped), nullptr = no neighbour
    uint64_t slot_bytes;
    uint64_t timeout_ns;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint64_t now_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag has reached `want` (wrap-safe); one thread per CTA calls this
__device__ void wait_for(const uint32_t *flag, uint32_t want, uint64_t timeout_ns, Header *mine, int what) {
    uint64_t t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        if ((int32_t)(ld_acquire_sys(flag) - want) >= 0) return;
        __nanosleep(64);
        if ((spins & 1023u) == 1023u) {
            const uint64_t t = now_ns();
            if (t0 == 0) t0 = t;
            if (t - t0 > timeout_ns) {
                mine->error = 1;
                printf("xgrid_b200 peer halo exchange: timed out waiting for %s %u (have %u)\n",
                       what == 0 ? "credit" : "ready", want, *flag);
                __trap();
            }
        }
    }
}

template <class V, bool FROM_MAILBOX>
__device__ __forceinline__ void copy_as(char *dst, const char *src, uint64_t bytes, uint64_t tid, uint64_t nthreads) {
    const uint64_t n = bytes / sizeof(V);
    V *d = reinterpret_cast<V *>(dst);
    const V *s = reinterpret_cast<const V *>(src);
    uint64_t i = tid;
    for (; i + 3 * nthreads < n; i += 4 * nthreads) {          // four loads in flight per thread
        V a, b, c, e;
        if (FROM_MAILBOX) {
            a = __ldcg(s + i); b = __ldcg(s + i + nthreads); c = __ldcg(s + i + 2 * nthreads); e = __ldcg(s + i + 3 * nthreads);
        } else {
            a = s[i]; b = s[i + nthreads]; c = s[i + 2 * nthreads]; e = s[i + 3 * nthreads];
        }
        d[i] = a; d[i + nthreads] = b; d[i + 2 * nthreads] = c; d[i + 3 * nthreads] = e;
    }
    for (; i < n; i += nthreads) d[i] = FROM_MAILBOX ? __ldcg(s + i) : s[i];
}

// rows written by a peer are read with ld.cg (L2 is where NVLink stores land; L1 may hold the slot's previous content)
template <bool FROM_MAILBOX>
__device__ void copy_bytes(char *dst, const char *src, uint64_t bytes, uint64_t tid, uint64_t nthreads) {
    const uint64_t mix = reinterpret_cast<uint64_t>(dst) | reinterpret_cast<uint64_t>(src) | bytes;
    if ((mix & 15) == 0) copy_as<uint4, FROM_MAILBOX>(dst, src, bytes, tid, nthreads);
    else if ((mix & 7) == 0) copy_as<unsigned long long, FROM_MAILBOX>(dst, src, bytes, tid, nthreads);
    else if ((mix & 3) == 0) copy_as<uint32_t, FROM_MAILBOX>(dst, src, bytes, tid, nthreads);
    else copy_as<unsigned char, FROM_MAILBOX>(dst, src, bytes, tid, nthreads);
}

__global__ void __launch_bounds__(THREADS, MIN_CTAS_PER_SM) xgb_peer_exchange_kernel(const Args a) {
    Header *mine = reinterpret_cast<Header *>(a.mine);
    __shared__ uint32_t s_seq;
    if (threadIdx.x == 0) s_seq = ld_acquire_sys(&mine->seq) + 1;
    __syncthreads();
    const uint32_t s = s_seq;
    const uint64_t par = (s & 1u) * a.slot_bytes;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;

    // ---- 1. push
    if (threadIdx.x == 0) {
        if (a.lo) wait_for(&mine->credit[0], s - 2, a.timeout_ns, mine, 0);
        if (a.hi) wait_for(&mine->credit[1], s - 2, a.timeout_ns, mine, 0);
    }
    __syncthreads();
    for (int k = 0; k < a.n; ++k) {
        const Item &it = a.items[k];
        // I am the HI neighbour of my lo neighbour: my first rows go to its in[1]; my last rows to the hi one's in[0]
        if (a.lo && it.send_lo) copy_bytes<false>(a.lo + HEADER_BYTES + 2 * a.slot_bytes + par + it.offset, it.send_lo, it.bytes, tid, nthreads);
        if (a.hi && it.send_hi) copy_bytes<false>(a.hi + HEADER_BYTES + par + it.offset, it.send_hi, it.bytes, tid, nthreads);
    }
    // the CTA's stores happen-before thread 0's fence through the barrier, and fences are cumulative: one system-scope
    // fence per CTA (not per thread) orders all of them before the counter, and the last CTA's before the flags
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&mine->done[0], 1u) == gridDim.x - 1) {
            __threadfence_system();
            if (a.lo) st_release_sys(&reinterpret_cast<Header *>(a.lo)->ready[1], s);
            if (a.hi) st_release_sys(&reinterpret_cast<Header *>(a.hi)->ready[0], s);
        }
    }

    // ---- 2. pull
    if (threadIdx.x == 0) {
        if (a.lo) wait_for(&mine->ready[0], s, a.timeout_ns, mine, 1);
        if (a.hi) wait_for(&mine->ready[1], s, a.timeout_ns, mine, 1);
    }
    __syncthreads();
    for (int k = 0; k < a.n; ++k) {
        const Item &it = a.items[k];
        if (a.lo && it.recv_lo) copy_bytes<true>(it.recv_lo, a.mine + HEADER_BYTES + par + it.offset, it.bytes, tid, nthreads);
        if (a.hi && it.recv_hi) copy_bytes<true>(it.recv_hi, a.mine + HEADER_BYTES + 2 * a.slot_bytes + par + it.offset, it.bytes, tid, nthreads);
    }
    __syncthreads();                         // every load of the slot has returned (its value has been stored)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&mine->done[1], 1u) == gridDim.x - 1) {
            // every CTA of this launch is past both phases: hand the slots back, reset the counters, advance
            mine->done[0] = 0;
            mine->done[1] = 0;
            __threadfence_system();
            if (a.lo) st_release_sys(&reinterpret_cast<Header *>(a.lo)->credit[1], s);
            if (a.hi) st_release_sys(&reinterpret_cast<Header *>(a.hi)->credit[0], s);
            st_release_sys(&mine->seq, s);
        }
    }
}

// ---- mailbox memory: CUDA virtual memory management, shared as a POSIX file descriptor ---------------------------------
// NOT cudaIpcGetMemHandle / cudaIpcOpenMemHandle: opening a legacy IPC handle enables device-wide peer access.  A
// cuMemCreate allocation is mapped into the neighbour with access granted for that one allocation only, which is what
// NCCL does for its own peer buffers.  (Measured on 2 x B200, the HBM-bound 3-D sweep ran ~3 % slower with the mailboxes
// mapped -- under either scheme, and whether or not a single byte was exchanged through them; not understood yet,
// profiles/r2_experiments.md "Halo exchange over peer memory".)
struct Vmm {
    bool ready = false;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
} vmm;

template <class F>
int vmm_entry(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr)
        return xgb_internal::fail((std::string("driver entry point '") + name + "' unavailable: " + cudaGetErrorString(e)).c_str());
    fn = reinterpret_cast<F>(p);
    return 0;
}

int load_vmm() {
    if (vmm.ready) return 0;
    if (vmm_entry("cuGetErrorString", vmm.GetErrorString)) return 1;
    if (vmm_entry("cuMemGetAllocationGranularity", vmm.MemGetAllocationGranularity)) return 1;
    if (vmm_entry("cuMemCreate", vmm.MemCreate)) return 1;
    if (vmm_entry("cuMemRelease", vmm.MemRelease)) return 1;
    if (vmm_entry("cuMemAddressReserve", vmm.MemAddressReserve)) return 1;
    if (vmm_entry("cuMemAddressFree", vmm.MemAddressFree)) return 1;
    if (vmm_entry("cuMemMap", vmm.MemMap)) return 1;
    if (vmm_entry("cuMemUnmap", vmm.MemUnmap)) return 1;
    if (vmm_entry("cuMemSetAccess", vmm.MemSetAccess)) return 1;
    if (vmm_entry("cuMemExportToShareableHandle", vmm.MemExportToShareableHandle)) return 1;
    if (vmm_entry("cuMemImportFromShareableHandle", vmm.MemImportFromShareableHandle)) return 1;
    vmm.ready = true;
    return 0;
}

#define PEER_CU(expr)                                                                               \
    do {                                                                                            \
        CUresult r_ = (expr);                                                                       \
        if (r_ != CUDA_SUCCESS) {                                                                   \
            const char *t_ = nullptr;                                                               \
            if (vmm.GetErrorString) vmm.GetErrorString(r_, &t_);                                    \
            return xgb_internal::fail((std::string(#expr) + ": " + (t_ ? t_ : "CUDA driver error")).c_str()); \
        }                                                                                           \
    } while (0)

#define PEER_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return xgb_internal::fail((std::string(#expr) + ": " + cudaGetErrorName(e_) + " (" +    \
                                       cudaGetErrorString(e_) + ")").c_str());                      \
    } while (0)

// what a rank publishes about its mailbox (64 bytes, opaque to the host side except for `fd`, which it replaces by the
// descriptor it received over a Unix socket before calling xgb_peer_open)
struct Ticket {
    uint32_t magic;
    int32_t fd;
    uint64_t bytes;
    int64_t pid;
    int32_t device;
    char pad[64 - 28];
};
static_assert(sizeof(Ticket) == 64, "mailbox ticket");
constexpr uint32_t TICKET_MAGIC = 0x58474250u;      // "XGBP"

struct Mailbox {
    char *base = nullptr;
    uint64_t slot_bytes = 0, bytes = 0;
    CUmemGenericAllocationHandle handle = 0;
    int fd = -1;
} box;

struct Mapping {
    char *base;
    uint64_t bytes;
    CUmemGenericAllocationHandle handle;
};
Mapping mappings[8];
int n_mappings = 0;

int current_device(int *dev) {
    PEER_CUDA(cudaGetDevice(dev));
    return 0;
}

int map_rw(CUdeviceptr *out, CUmemGenericAllocationHandle h, size_t bytes, size_t gran, int dev) {
    CUdeviceptr va = 0;
    PEER_CU(vmm.MemAddressReserve(&va, bytes, gran, 0, 0));
    PEER_CU(vmm.MemMap(va, bytes, 0, h, 0));
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    PEER_CU(vmm.MemSetAccess(va, bytes, &acc, 1));
    *out = va;
    return 0;
}

}  // namespace

extern "C" {

int xgb_peer_create(uint64_t slot_bytes, void *ticket_64B) {
    if (xgb_internal::require_init()) return 1;
    if (load_vmm()) return 1;
    if (box.base) return xgb_internal::fail("xgb_peer_create: a mailbox exists (xgb_peer_destroy first)");
    int dev = 0;
    if (current_device(&dev)) return 1;
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    PEER_CU(vmm.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    slot_bytes = (slot_bytes + 255) / 256 * 256;
    size_t total = HEADER_BYTES + 4 * slot_bytes;
    total = (total + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h = 0;
    PEER_CU(vmm.MemCreate(&h, total, &prop, 0));
    CUdeviceptr va = 0;
    if (map_rw(&va, h, total, gran, dev)) return 1;
    int fd = -1;
    PEER_CU(vmm.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    PEER_CUDA(cudaMemset(reinterpret_cast<void *>(va), 0, HEADER_BYTES));
    PEER_CUDA(cudaDeviceSynchronize());
    box.base = reinterpret_cast<char *>(va);
    box.slot_bytes = slot_bytes;
    box.bytes = total;
    box.handle = h;
    box.fd = fd;
    Ticket t;
    memset(&t, 0, sizeof(t));
    t.magic = TICKET_MAGIC;
    t.fd = fd;
    t.bytes = total;
    t.pid = (int64_t)getpid();
    t.device = dev;
    memcpy(ticket_64B, &t, 64);
    return 0;
}

/* The ticket of this very process maps to its own mailbox (a ring of one rank).  For another rank's ticket the caller has
 * replaced `fd` by the descriptor it received from that rank (SCM_RIGHTS); it stays the caller's to close. */
int xgb_peer_open(const void *ticket_64B, void **mailbox) {
    if (xgb_internal::require_init()) return 1;
    if (!box.base) return xgb_internal::fail("xgb_peer_open: xgb_peer_create has not been called");
    Ticket t;
    memcpy(&t, ticket_64B, 64);
    if (t.magic != TICKET_MAGIC) return xgb_internal::fail("xgb_peer_open: not a mailbox ticket");
    if (t.pid == (int64_t)getpid() && t.fd == box.fd) {
        *mailbox = box.base;
        return 0;
    }
    if (n_mappings == 8) return xgb_internal::fail("xgb_peer_open: too many mapped mailboxes");
    int dev = 0;
    if (current_device(&dev)) return 1;
    CUmemGenericAllocationHandle h = 0;
    PEER_CU(vmm.MemImportFromShareableHandle(&h, reinterpret_cast<void *>(static_cast<uintptr_t>(t.fd)),
                                             CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    size_t gran = 0;
    PEER_CU(vmm.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    CUdeviceptr va = 0;
    if (map_rw(&va, h, t.bytes, gran, dev)) return 1;
    mappings[n_mappings++] = Mapping{reinterpret_cast<char *>(va), t.bytes, h};
    *mailbox = reinterpret_cast<void *>(va);
    return 0;
}

int xgb_peer_close(void *mailbox) {
    if (!mailbox || mailbox == box.base) return 0;
    for (int i = 0; i < n_mappings; ++i) {
        if (mappings[i].base == mailbox) {
            PEER_CUDA(cudaDeviceSynchronize());
            PEER_CU(vmm.MemUnmap(reinterpret_cast<CUdeviceptr>(mappings[i].base), mappings[i].bytes));
            PEER_CU(vmm.MemAddressFree(reinterpret_cast<CUdeviceptr>(mappings[i].base), mappings[i].bytes));
            PEER_CU(vmm.MemRelease(mappings[i].handle));
            mappings[i] = mappings[--n_mappings];
            return 0;
        }
    }
    return xgb_internal::fail("xgb_peer_close: not a mapped mailbox");
}

int xgb_peer_destroy(void) {
    if (box.base) {
        PEER_CUDA(cudaDeviceSynchronize());
        PEER_CU(vmm.MemUnmap(reinterpret_cast<CUdeviceptr>(box.base), box.bytes));
        PEER_CU(vmm.MemAddressFree(reinterpret_cast<CUdeviceptr>(box.base), box.bytes));
        PEER_CU(vmm.MemRelease(box.handle));
        if (box.fd >= 0) close(box.fd);
        box = Mailbox();
    }
    return 0;
}

/* Zero the exchange counters and flags of this rank's mailbox (device drained first).  Collective in effect: the
 * host side calls it on every rank between two barriers when the neighbour relation changes (chain <-> ring), because
 * a rank that had no neighbour on one side has never received credits from that side. */
int xgb_peer_reset(void) {
    if (!box.base) return 0;
    PEER_CUDA(cudaDeviceSynchronize());
    PEER_CUDA(cudaMemset(box.base, 0, HEADER_BYTES));
    PEER_CUDA(cudaDeviceSynchronize());
    return 0;
}

int xgb_peer_slot_bytes(uint64_t *slot_bytes) {
    *slot_bytes = box.slot_bytes;
    return 0;
}

/* descs as for xgb_halo_exchange (the rank fields are ignored: a neighbour exists where its mailbox is given).
 * Levels are packed into the slot in order; a batch that does not fit is split over several launches. */
int xgb_peer_exchange(const xgb_halo_desc *descs, int n, void *lo_mailbox, void *hi_mailbox, xgb_handle stream) {
    if (xgb_internal::require_init()) return 1;
    if (!box.base) return xgb_internal::fail("xgb_peer_exchange: xgb_peer_create has not been called");
    if (!lo_mailbox && !hi_mailbox) return 0;
    static uint64_t timeout_ns = 0;
    if (timeout_ns == 0) {
        const char *env = getenv("XGB_PEER_TIMEOUT_S");
        const double sec = env ? atof(env) : 120.0;
        timeout_ns = (uint64_t)((sec > 0 ? sec : 120.0) * 1e9);
    }
    static int max_ctas = 0;
    if (max_ctas == 0) {
        const char *env = getenv("XGB_PEER_CTAS");
        max_ctas = env ? atoi(env) : 128;
        if (max_ctas < 1) max_ctas = 1;
    }
    static int greatest_priority = 1 << 30;
    if (greatest_priority == 1 << 30) {
        int least = 0;
        PEER_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest_priority));
    }
    cudaStream_t s = xgb_internal::stream_of(stream);
    int i = 0;
    while (i < n) {
        Args a;
        memset(&a, 0, sizeof(a));
        a.mine = box.base;
        a.lo = static_cast<char *>(lo_mailbox);
        a.hi = static_cast<char *>(hi_mailbox);
        a.slot_bytes = box.slot_bytes;
        a.timeout_ns = timeout_ns;
        uint64_t used = 0;
        while (i < n && a.n < MAX_ITEMS) {
            const xgb_halo_desc &d = descs[i];
            const uint64_t padded = (d.bytes + 15) / 16 * 16;
            if (padded > box.slot_bytes)
                return xgb_internal::fail("xgb_peer_exchange: a level's halo is larger than the mailbox slot "
                                          "(the host side reserves it before the call: dist.PeerTransport.reserve)");
            if (used + padded > box.slot_bytes) break;
            Item &it = a.items[a.n++];
            it.send_lo = static_cast<const char *>(d.send_lo);
            it.recv_lo = static_cast<char *>(d.recv_lo);
            it.send_hi = static_cast<const char *>(d.send_hi);
            it.recv_hi = static_cast<char *>(d.recv_hi);
            it.bytes = d.bytes;
            it.offset = used;
            used += padded;
            ++i;
        }
        const uint64_t per_cta = 16 * 1024;
        uint64_t ctas = (2 * used + per_cta - 1) / per_cta;
        if (ctas < 1) ctas = 1;
        if (ctas > (uint64_t)max_ctas) ctas = max_ctas;
        // explicit launch priority: a kernel node recorded into a CUDA graph keeps it (the graph is instantiated with
        // cudaGraphInstantiateFlagUseNodePriority), so the exchange overtakes the interior sweep's CTAs on replay too
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)ctas, 1, 1);
        cfg.blockDim = dim3(THREADS, 1, 1);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributePriority;
        attr[0].val.priority = greatest_priority;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PEER_CUDA(cudaLaunchKernelEx(&cfg, xgb_peer_exchange_kernel, a));
        xgb_internal::count_launch();
    }
    return 0;
}

}  // extern "C"
