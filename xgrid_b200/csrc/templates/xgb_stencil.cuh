// xgb_stencil.cuh -- hand-written device templates for xgrid sweeps on sm_100a.
//
// The code generator (xgrid_b200/lang/cudagen.py) lowers one *sweep group* -- a
// run of stencil statements that the reference executes as separate mask-
// predicated full-grid loop nests (xgrid/lang/generator.py:285-364) -- to one
// __global__ function assembled from the primitives below.  The generator only
// emits the per-point expressions and the list of row windows they read; every
// performance-relevant decision (thread->point mapping, vector width, window
// loads, mask fetch, predicated stores, marching, shared-memory staging) lives
// here.
//
// Memory model (DESIGN.md "Data layout in HBM"):
//   * a time level is a C-order array padded with `ghost` zero rows on both
//     sides of axis 0 plus a small linear slack, so a relative tap is a single
//     signed linear offset and out-of-range taps never fault (the reference
//     reads whatever is adjacent);
//   * the boundary mask is uint8 per point; `flags` holds one byte per
//     XGB_CHUNK consecutive points that is non-zero iff any mask byte in the
//     chunk is non-zero, so interior threads never touch the mask.
#pragma once

typedef signed char int8_t;
typedef short int16_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;

#define XGB_CHUNK_SHIFT 7
#define XGB_CHUNK (1 << XGB_CHUNK_SHIFT)
#define XGB_DEV __device__ __forceinline__

namespace xgb {

// --------------------------------------------------------------------------- math
// gcc folds pow(x, 2.0) to x*x at the reference's default -O2;
// emitting the product keeps fp64 results bit-identical.
template <class T> XGB_DEV T sq(T x) { return x * x; }

// C semantics of the reference's int32 `/` and `%` are what CUDA C gives too.

// IEEE-exact floating division with a shortcut for zero numerators: (+-0) / y for any
// non-zero, non-NaN y is +-0 with sign = sign(x) xor sign(y).  div.rn.f64 sends every
// zero / subnormal numerator through its special-operand subroutine (~10x the
// instructions of the fast path); quiescent regions of a PDE field are exactly zero.
XGB_DEV double fdiv(double x, double y) {
    if (x == 0.0 && y != 0.0 && y == y)
        return __longlong_as_double((__double_as_longlong(x) ^ __double_as_longlong(y)) &
                                    (long long)0x8000000000000000ULL);
    return x / y;
}
XGB_DEV float fdiv(float x, float y) {
    if (x == 0.0f && y != 0.0f && y == y)
        return __int_as_float((__float_as_int(x) ^ __float_as_int(y)) & (int)0x80000000u);
    return x / y;
}
XGB_DEV double fdiv(double x, float y) { return fdiv(x, (double)y); }
XGB_DEV double fdiv(float x, double y) { return fdiv((double)x, y); }

// IEEE-exact fp64 division by a divisor that does not change from point to point (dx, 2*dy,
// dx^2+dy^2 ...): the reciprocal is rounded once per thread, every quotient then costs one
// multiply and four FMAs instead of div.rn.f64's reciprocal seed + Newton + correction + range
// check.  With r = RN(1/y):  q0 = RN(x r);  q1 = RN(q0 + (x - q0 y) r) is within 1/2 ulp + 2^-51 ulp
// of x/y, i.e. a faithful quotient; one more correction step from a faithful quotient with a
// correctly rounded reciprocal yields RN(x/y) (Markstein's theorem).  The residuals are exact
// only while nothing underflows or overflows, so the sequence is used for 2^-900 <= |x| < 2^900
// and 2^-100 <= |y| <= 2^100; zero numerators return the signed zero directly and everything else
// (subnormals, infinities, NaNs, extreme exponents) takes the ordinary exact path, out of line.  tests/test_invdiv_gpu.py compares it with the host's
// IEEE division over exponent sweeps, special values and adversarial divisors.
__device__ __noinline__ double fdiv_outlined(double x, double y) { return fdiv(x, y); }   // rare path: keep call sites small
struct InvDiv {
    double y, r;
    bool ok;
    XGB_DEV explicit InvDiv(double y_) : y(y_), r(__drcp_rn(y_)) {
        const unsigned ey = ((unsigned)__double2hiint(y_) >> 20) & 0x7ffu;
        ok = (ey - 923u) <= 200u;                      // biased exponent 923 .. 1123  <=>  2^-100 <= |y| < 2^101
    }
    XGB_DEV double div(double x) const {
        const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
        if (ok && (ex - 123u) < 1800u) {               // 2^-900 <= |x| < 2^900
            double q = x * r;
            double e = fma(-q, y, x);
            q = fma(e, r, q);
            e = fma(-q, y, x);
            return fma(e, r, q);
        }
        if (ok && x == 0.0)                            // quiescent regions of a PDE field are exactly zero
            return __longlong_as_double((__double_as_longlong(x) ^ __double_as_longlong(y)) &
                                        (long long)0x8000000000000000ULL);
        return fdiv_outlined(x, y);
    }
};

// --------------------------------------------------------------------------- vector access
template <int BYTES> struct Pack;
template <> struct Pack<1>  { typedef uint8_t  type; };
template <> struct Pack<2>  { typedef uint16_t type; };
template <> struct Pack<4>  { typedef uint32_t type; };
template <> struct Pack<8>  { typedef uint2    type; };
template <> struct Pack<16> { typedef uint4    type; };

// load V consecutive elements starting at an address aligned to V*sizeof(T)
template <class T, int V>
XGB_DEV void ld_vec(const T *p, T (&out)[V]) {
    constexpr int B = V * (int)sizeof(T);
    if constexpr (B <= 16) {
        typedef typename Pack<B>::type P;
        union { P pk; T el[V]; } u;
        u.pk = *reinterpret_cast<const P *>(p);
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = u.el[i];
    } else {
        static_assert(B % 16 == 0, "vector must be a multiple of 16 bytes");
        constexpr int N = B / 16, E = 16 / (int)sizeof(T);
        union { uint4 pk[N]; T el[V]; } u;
#pragma unroll
        for (int i = 0; i < N; ++i) u.pk[i] = reinterpret_cast<const uint4 *>(p)[i];
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = u.el[i];
        (void)E;
    }
}

template <class T, int V>
XGB_DEV void st_vec(T *p, const T (&v)[V]) {
    constexpr int B = V * (int)sizeof(T);
    if constexpr (B <= 16) {
        typedef typename Pack<B>::type P;
        union { P pk; T el[V]; } u;
#pragma unroll
        for (int i = 0; i < V; ++i) u.el[i] = v[i];
        *reinterpret_cast<P *>(p) = u.pk;
    } else {
        constexpr int N = B / 16;
        union { uint4 pk[N]; T el[V]; } u;
#pragma unroll
        for (int i = 0; i < V; ++i) u.el[i] = v[i];
#pragma unroll
        for (int i = 0; i < N; ++i) reinterpret_cast<uint4 *>(p)[i] = u.pk[i];
    }
}

// Row window: w[k] = row[LO + k] for k in [0, V + HI - LO), where `row` points
// at the thread's first point (aligned to V elements) shifted by the tap's
// outer-axis offset.  The aligned body is one vector load; the LO/HI fringes
// are scalar loads that hit L1 (they are a neighbouring thread's body).
template <class T, int V, int LO, int HI>
XGB_DEV void ld_window(const T *row, T (&w)[V + HI - LO]) {
    static_assert(LO <= 0 && HI >= 0, "window must contain the centre");
    if constexpr (V > 1) {
        T body[V];
        ld_vec<T, V>(row, body);
#pragma unroll
        for (int i = 0; i < V; ++i) w[i - LO] = body[i];
    } else {
        w[-LO] = row[0];
    }
#pragma unroll
    for (int k = LO; k < 0; ++k) w[k - LO] = row[k];
#pragma unroll
    for (int k = 0; k < HI; ++k) w[V - LO + k] = row[V + k];
}

// Window assembled from a register-resident body: w[k] = row[LO + k].  The
// fringes come from the neighbouring lanes' bodies by warp shuffle; the lanes
// on a warp edge (or at the end of a grid row) read them from global memory
// instead.  Must be executed by all 32 lanes (no divergence around it); lanes
// own consecutive V-element bodies of the same row.
template <class T, int V, int LO, int HI>
XGB_DEV void window_from_body(const T *row, const T (&body)[V], bool edge_l, bool edge_r,
                              T (&w)[V + HI - LO]) {
    static_assert(LO <= 0 && HI >= 0 && -LO <= V && HI <= V, "fringe wider than the body");
#pragma unroll
    for (int i = 0; i < V; ++i) w[i - LO] = body[i];
#pragma unroll
    for (int k = LO; k < 0; ++k) {
        T x = __shfl_up_sync(0xffffffffu, body[V + k], 1);
        if (edge_l) x = row[k];
        w[k - LO] = x;
    }
#pragma unroll
    for (int k = 0; k < HI; ++k) {
        T x = __shfl_down_sync(0xffffffffu, body[k], 1);
        if (edge_r) x = row[V + k];
        w[V - LO + k] = x;
    }
}

// Store v[i] where bit i of `written` is set (points whose mask matched no
// statement keep the ring buffer's previous content).
template <class T, int V>
XGB_DEV void st_pred(T *p, const T (&v)[V], unsigned written) {
    constexpr unsigned ALL = (V >= 32) ? 0xffffffffu : ((1u << V) - 1u);
    if (written == ALL) {
        st_vec<T, V>(p, v);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (written & (1u << i)) p[i] = v[i];
    }
}

// --------------------------------------------------------------------------- mask
// m[i] = boundary value of point base+i (0 when the chunk flag says "all zero").
template <int V>
XGB_DEV void ld_mask(const uint8_t *mask, const uint8_t *flags, int64_t base, int (&m)[V]) {
#pragma unroll
    for (int i = 0; i < V; ++i) m[i] = 0;
    if (mask == nullptr) return;
    if (flags != nullptr && flags[base >> XGB_CHUNK_SHIFT] == 0) return;
    if constexpr (V == 1) {
        m[0] = mask[base];
    } else {
        uint8_t b[V];
        ld_vec<uint8_t, V>(mask + base, b);
#pragma unroll
        for (int i = 0; i < V; ++i) m[i] = b[i];
    }
}

// --------------------------------------------------------------------------- async tile pipeline
// sm_100a bulk-copy pipeline for the "tiled" sweep variant: one producer lane
// streams halo'd planes of every input level into a ring of shared-memory stages
// with cp.async.bulk (the TMA engine's linear mode -- SASS UBLKCP) completing on
// an mbarrier; consumer warps wait on the "full" barrier, compute one output plane
// from shared memory and release the oldest plane through the "empty" barrier.
// Bytes in flight live in shared memory, not in registers, so occupancy does not
// pay for memory-level parallelism.
namespace pipe {

XGB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

XGB_DEV void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
XGB_DEV void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
XGB_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
XGB_DEV void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
XGB_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "XGB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra XGB_DONE;\n"
        "bra XGB_WAIT;\n"
        "XGB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16-B aligned)
XGB_DEV void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace pipe

// window from a shared-memory row: w[k] = row[LO + k]; `row` points at the thread's
// body (16-B aligned), fringes are scalar LDS.
template <class T, int V, int LO, int HI>
XGB_DEV void lds_window(const T *row, T (&w)[V + HI - LO]) {
    T body[V];
    ld_vec<T, V>(row, body);
#pragma unroll
    for (int i = 0; i < V; ++i) w[i - LO] = body[i];
#pragma unroll
    for (int k = LO; k < 0; ++k) w[k - LO] = row[k];
#pragma unroll
    for (int k = 0; k < HI; ++k) w[V - LO + k] = row[V + k];
}

// Mask fetch split in two so that the (global) flag byte can be requested one
// plane ahead of its use: ld_flag() early, ld_mask_flagged() at the point of use.
XGB_DEV int ld_flag(const uint8_t *mask, const uint8_t *flags, int64_t base) {
    if (mask == nullptr) return 0;
    if (flags == nullptr) return 1;
    return flags[base >> XGB_CHUNK_SHIFT];
}
// Warp-uniform form for the pipeline kernels: "does any point of [lo, hi] (one warp's vector span of one row)
// lie in a chunk that holds a boundary point?"  A span is shorter than a chunk, so it touches at most two.
XGB_DEV int ld_flag_span(const uint8_t *mask, const uint8_t *flags, int64_t lo, int64_t hi) {
    if (mask == nullptr) return 0;
    if (flags == nullptr) return 1;
    return __ldg(flags + (lo >> XGB_CHUNK_SHIFT)) | __ldg(flags + (hi >> XGB_CHUNK_SHIFT));
}
template <int V>
XGB_DEV void ld_mask_flagged(const uint8_t *mask, int flag, int64_t base, int (&m)[V]) {
#pragma unroll
    for (int i = 0; i < V; ++i) m[i] = 0;
    if (flag == 0) return;
    if constexpr (V == 1) {
        m[0] = mask[base];
    } else {
        uint8_t b[V];
        ld_vec<uint8_t, V>(mask + base, b);
#pragma unroll
        for (int i = 0; i < V; ++i) m[i] = b[i];
    }
}

// --------------------------------------------------------------------------- geometry
// Dense sweeps view an N-d grid as rows x cols (cols = contiguous axis).
// blockDim = (TX, TY); a thread owns V consecutive columns of one row.
struct Tile {
    int64_t row;    // global row of this thread
    int64_t col;    // first column of this thread
    bool active;
};

template <int V>
XGB_DEV Tile dense_tile(int64_t rows, int64_t cols) {
    Tile t;
    const int64_t rb = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
    t.row = rb * blockDim.y + threadIdx.y;
    t.col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    t.active = (t.row < rows) && (t.col < cols);
    return t;
}

// overstep="limit" / "wrap" (xgrid/lang/generator.py:172-177) with per-axis
// extents paired correctly (the reference pairs them wrongly off-square, F1).
XGB_DEV int64_t clamp_idx(int64_t i, int64_t n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }
XGB_DEV int64_t wrap_idx(int64_t i, int64_t n) { i %= n; return i < 0 ? i + n : i; }
// Axis 0 of a slab: where the slab has a neighbour on that side (`open`), an index off the end is
// NOT clamped / wrapped locally -- it addresses the ghost rows, which hold the neighbour's rows
// (for "wrap" the ranks form a ring, so the rows of the far end of the global grid).
XGB_DEV int64_t clamp_idx0(int64_t i, int64_t n, int64_t open_lo, int64_t open_hi) {
    if (i < 0) return open_lo ? i : 0;
    if (i >= n) return open_hi ? i : n - 1;
    return i;
}
XGB_DEV int64_t wrap_idx0(int64_t i, int64_t n, int64_t open_lo, int64_t open_hi) {
    if (i < 0) return open_lo ? i : wrap_idx(i, n);
    if (i >= n) return open_hi ? i : wrap_idx(i, n);
    return i;
}

}  // namespace xgb
