// xgb_mask.cu -- device-side compilation of a grid's boundary mask.
//
// The reference keeps `Grid.boundary` as an int32 NumPy array and tests it per point in every
// sweep (xgrid/lang/generator.py:295-298).  The B200 backend compiles it once per change into
//   dst   : one uint8 per point (what the sweep kernels read),
//   flags : one byte per 128 consecutive points, non-zero iff any mask byte in the chunk is,
//   hist  : number of points per mask value (sparse-statement planning),
//   bad   : set when a value does not fit the uint8 encoding (255 is reserved for "outside").
// One fused pass over the int32 data; launched with 128 threads per block so that a block
// iteration covers exactly one flag chunk.  JIT-compiled with NVRTC like the sweep kernels.
typedef unsigned char uint8_t;
typedef int int32_t;
typedef long long int64_t;

struct xgb_mask_params {
    const int32_t *src;      // piece of the int32 mask resident on the device
    uint8_t *dst;            // uint8 mask, already offset to the piece
    uint8_t *flags;          // chunk flags, already offset to the piece
    unsigned long long *hist;
    int *bad;
    int64_t n;               // valid points in this piece
    int64_t n_padded;        // points to write (multiple of 128; tail is zero-filled)
};

extern "C" __global__ void __launch_bounds__(128) xgb_mask_compile(const __grid_constant__ xgb_mask_params p)
{
    __shared__ unsigned int bins[256];
    bins[threadIdx.x] = 0u;
    bins[threadIdx.x + 128] = 0u;
    __syncthreads();
    const int64_t chunks = p.n_padded >> 7;
    for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const int64_t i = (c << 7) + threadIdx.x;
        int v = 0;
        if (i < p.n) {
            v = p.src[i];
            if (v < 0 || v > 254) { atomicOr(p.bad, 1); v = 255; }
            atomicAdd(&bins[v], 1u);
        }
        p.dst[i] = (uint8_t)v;
        const int any = __syncthreads_or(v != 0);
        if (threadIdx.x == 0) p.flags[c] = (uint8_t)(any != 0);
    }
    __syncthreads();
    if (bins[threadIdx.x]) atomicAdd(&p.hist[threadIdx.x], (unsigned long long)bins[threadIdx.x]);
    if (bins[threadIdx.x + 128]) atomicAdd(&p.hist[threadIdx.x + 128], (unsigned long long)bins[threadIdx.x + 128]);
}
