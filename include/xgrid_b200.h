/*
 * xgrid_b200.h -- C ABI of the B200 execution backend for xgrid.
 *
 * This is the drop-in boundary.  In the reference the
 * "native" side of the hot path is one gcc-compiled symbol per kernel reached
 * through ctypes (xgrid/util/ffi.py:18-35) with NumPy-owned buffers passed by
 * pointer inside a by-value struct (xgrid/xgrid/__init__.py:60-68).  Here the
 * per-kernel symbol becomes a JIT-compiled sm_100a cubin and the stable ABI is
 * this runtime: it owns device memory, compiles CUDA C, loads modules and
 * launches / replays them.  Plain C types only; every call returns 0 on
 * success, a non-zero status otherwise, and xgb_last_error() then returns the
 * CUDA / NVRTC / NCCL text for the calling thread.  Nothing here aborts.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef XGRID_B200_H
#define XGRID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XGB_ABI_VERSION 1

typedef uint64_t xgb_handle; /* opaque: stream, event, module, function, graph */

typedef struct xgb_device_info {
    int32_t ordinal;
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int32_t max_smem_per_block_optin;
    int32_t l2_bytes;
    int32_t clock_khz;
    int32_t reserved;
    uint64_t total_mem;
    char name[64];
} xgb_device_info;

/* ---- lifecycle / diagnostics ------------------------------------------- */
int xgb_abi_version(void);
const char *xgb_last_error(void);
/* replaces: xgrid.init() picking a host compiler (xgrid/util/init.py:48-66,
 * xgrid/util/ffi.py:41-57) -- binds the calling process to one GPU. */
int xgb_init(int device);
int xgb_shutdown(void);
int xgb_device_count(int *count);
int xgb_get_device_info(xgb_device_info *out);
int xgb_device_sync(void);

/* ---- device / pinned memory -------------------------------------------- */
/* replaces: np.zeros time levels + int32 mask owned by NumPy
 * (xgrid/xgrid/__init__.py:38-47).  xgb_alloc returns zero-filled memory. */
int xgb_alloc(size_t bytes, void **dptr);
int xgb_free(void *dptr);
int xgb_memset(void *dptr, int byte, size_t bytes, xgb_handle stream);
int xgb_h2d(void *dst_dev, const void *src_host, size_t bytes, xgb_handle stream);
int xgb_d2h(void *dst_host, const void *src_dev, size_t bytes, xgb_handle stream);
int xgb_d2d(void *dst_dev, const void *src_dev, size_t bytes, xgb_handle stream);
int xgb_host_alloc(size_t bytes, void **hptr);   /* pinned */
int xgb_host_free(void *hptr);
int xgb_host_register(void *hptr, size_t bytes); /* pin an existing NumPy mirror */
int xgb_host_unregister(void *hptr);
int xgb_mem_info(uint64_t *free_bytes, uint64_t *total_bytes);
/* Copies between PAGEABLE host memory and the device through page-locked staging chunks filled /
 * drained by several host threads (replaces: nothing -- the reference's arrays never leave the host;
 * this is the upload of Grid.now written by the user and the download behind Grid.now).
 * xgb_h2d_staged returns once the source has been read (the transfer is ordered on `stream` like
 * xgb_h2d from a pinned buffer); xgb_d2h_staged returns when dst_host is complete. */
int xgb_h2d_staged(void *dst_dev, const void *src_host, size_t bytes, xgb_handle stream);
int xgb_d2h_staged(void *dst_host, const void *src_dev, size_t bytes, xgb_handle stream);
/* Host half of the mask upload.  replaces: the reference handing the int32 `Grid.boundary` array to every
 * sweep (xgrid/xgrid/__init__.py:41,66; tested per point, xgrid/lang/generator.py:295-298).  Packs
 * src[0:n] (int32) into dst[0:n_padded] (one byte per point; n_padded a multiple of 128, tail zero),
 * flags[0:n_padded/128] (1 iff any byte of the 128-point chunk is non-zero) and hist[256] (points per
 * value) with several host threads, so that one byte per point -- not four -- crosses PCIe.  Values
 * outside [0, 254] are stored as 255 and set *bad (255 = "outside the grid" on the device).  Pure host
 * work: needs no GPU and no xgb_init.  threads <= 0: pick from the machine. */
int xgb_mask_pack(const int32_t *src, size_t n, size_t n_padded, uint8_t *dst, uint8_t *flags,
                  uint64_t *hist_256, int *bad, int threads);

/* ---- streams / events --------------------------------------------------- */
/* stream 0 is the backend's default compute stream (created by xgb_init). */
int xgb_stream_create(xgb_handle *stream);
/* high_priority != 0: greatest stream priority (halo-exchange stream, so that NCCL kernels
 * are scheduled ahead of the interior sweep they overlap with) */
int xgb_stream_create_ex(xgb_handle *stream, int high_priority);
int xgb_stream_destroy(xgb_handle stream);
int xgb_stream_sync(xgb_handle stream);
int xgb_stream_raw(xgb_handle stream, void **cuda_stream); /* cudaStream_t for interop */
int xgb_event_create(xgb_handle *event);
int xgb_event_destroy(xgb_handle event);
int xgb_event_record(xgb_handle event, xgb_handle stream);
int xgb_event_sync(xgb_handle event);
int xgb_event_elapsed_ms(xgb_handle start, xgb_handle stop, float *ms);
int xgb_stream_wait_event(xgb_handle stream, xgb_handle event);

/* ---- JIT: CUDA C -> sm_100a cubin -> module -> function ------------------ */
/* replaces: Compiler.compile (gcc -shared, xgrid/util/ffi.py:59-94).  NVRTC
 * in-process; `headers` are (name, source) pairs resolvable by #include.
 * On success *image is malloc'ed (release with xgb_release).  *log (may be
 * NULL) receives the compiler log on success and failure alike. */
int xgb_compile(const char *source, const char *name,
                const char *const *options, int n_options,
                const char *const *header_names, const char *const *header_sources, int n_headers,
                void **image, size_t *image_bytes, char **log);
int xgb_release(void *p);
/* replaces: ctypes.cdll.LoadLibrary + getattr(lib, entry_point)
 * (xgrid/util/ffi.py:14-22). */
int xgb_module_load(const void *image, size_t image_bytes, xgb_handle *module);
int xgb_module_unload(xgb_handle module);
int xgb_get_function(xgb_handle module, const char *name, xgb_handle *function);
int xgb_function_info(xgb_handle function, int *regs, int *static_smem, int *local_bytes,
                      int *max_threads);
int xgb_function_set_dynamic_smem(xgb_handle function, int bytes);
int xgb_occupancy(xgb_handle function, int block_threads, int dynamic_smem, int *blocks_per_sm);

/* ---- launch -------------------------------------------------------------- */
/* replaces: handler(*serialized_args) (xgrid/util/ffi.py:34).  Every generated
 * kernel takes ONE by-value parameter struct; `params` points at its bytes. */
int xgb_launch(xgb_handle function, const uint32_t grid[3], const uint32_t block[3],
               uint32_t dynamic_smem, xgb_handle stream, const void *params, size_t param_bytes);
/* kernels launched through this ABI since xgb_init (graph replays count their
 * kernel nodes) */
int xgb_launch_count(uint64_t *count);

/* ---- CUDA graphs: one Operator.__call__ replayed as a unit --------------- */
int xgb_graph_begin(xgb_handle stream);
int xgb_graph_end(xgb_handle stream, xgb_handle *graph_exec, int *kernel_nodes);
int xgb_graph_launch(xgb_handle graph_exec, xgb_handle stream);
int xgb_graph_destroy(xgb_handle graph_exec);

/* ---- multi-GPU: slab halo exchange (one process per GPU) ------------------ */
/* new work, no reference counterpart. */
int xgb_nccl_load(const char *libnccl_path);
int xgb_nccl_unique_id(void *id_128B);
int xgb_nccl_init(const void *id_128B, int rank, int n_ranks);
int xgb_nccl_shutdown(void);
typedef struct xgb_halo_desc {
    void *send_lo;   /* first owned rows  -> rank-1 */
    void *recv_lo;   /* ghost rows below  <- rank-1 */
    void *send_hi;   /* last owned rows   -> rank+1 */
    void *recv_hi;   /* ghost rows above  <- rank+1 */
    uint64_t bytes;  /* per face */
    int32_t lo_rank; /* -1 = no neighbour */
    int32_t hi_rank;
} xgb_halo_desc;
int xgb_halo_exchange(const xgb_halo_desc *descs, int n, xgb_handle stream);

/* ---- the same exchange over peer memory (NVLink / NVSwitch), one kernel per exchange: csrc/xgb_peer.cu ----
 * New work like the block above: the reference is single-address-space OpenMP.
 * Every rank creates a mailbox in its own HBM (a cuMemCreate allocation exported as a POSIX file descriptor) and
 * publishes a 64-byte ticket {magic, int32 fd at byte 4, size, pid, device}.  A neighbour receives the descriptor
 * itself over a Unix socket (SCM_RIGHTS; the host side does that), writes its own copy of the descriptor number into
 * the ticket and calls xgb_peer_open, which maps the mailbox with access for that one allocation -- no device-wide
 * peer access.  xgb_peer_exchange pushes the edge rows into the neighbours' mailboxes, signals, waits for theirs and
 * copies them into the ghost rows -- all inside one kernel on `stream`; the rank fields of the descs are ignored (a
 * neighbour exists where its mailbox is given).  All exchanges of a process must be issued on ONE stream, in the
 * same order on every rank. */
int xgb_peer_create(uint64_t slot_bytes, void *ticket_64B);
int xgb_peer_open(const void *ticket_64B, void **mailbox);          /* the rank's own ticket maps to its own mailbox */
int xgb_peer_close(void *mailbox);
int xgb_peer_destroy(void);
int xgb_peer_reset(void);       /* zero the exchange counters (every rank, between two barriers: neighbours changed) */
int xgb_peer_slot_bytes(uint64_t *slot_bytes);
int xgb_peer_exchange(const xgb_halo_desc *descs, int n, void *lo_mailbox, void *hi_mailbox, xgb_handle stream);

#ifdef __cplusplus
}
#endif
#endif /* XGRID_B200_H */
