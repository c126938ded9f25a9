// tanw_long.cu -- instantiations of the chained-stripe kernels (one whole-manuscript pair).
#include "tanw_launch.h"

namespace tanw {

// Column 0 of the matrices for rows r0+1 .. r0+nb as hand-over records (textSeqCompare.py:54-56:
// M[i][0] = Y[i][0] = gap_extend * i; the Q slot carries the M tag), so that stripe 0 consumes its
// left boundary exactly like every other stripe.
__global__ void long_col0_kernel(int4 *rec, int nb, int r0, int bg, int epoch)
{
    const int r = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= nb) {
        const int v = bg * (r0 + r);
        rec[r] = make_int4(v | kTagM, epoch, v, epoch);
    }
}

// Traceback of one band of a chained-stripe pair.  state = {x, y, st, k} persists between bands
// (x, y in matrix coordinates); `init` starts at (n, m), `final` flushes the remainders
// (textSeqCompare.py:154-164) and moves the op string to the start of its buffer.
// Warp 0 walks; the other warps of the block sleep at the barrier and only help with that last
// move: one warp shifting the 180 000 ops of config 5 a byte per lane and round trip took 0.5 ms
// (12 % of the traceback; 40 of the 270 us of a single page).
constexpr int kTraceThreads = 256;
constexpr int kShiftDepth = 8;                            // bytes per thread in flight
__global__ void __launch_bounds__(kTraceThreads)
trace_long_kernel(const uint8_t *ptr, const PairDesc *pd, int cfull, int r0, int nb, int init, int final,
                  int *state, uint8_t *ops_base, int *ops_len)
{
    __shared__ unsigned tile[kTileRows * kTileStride];
    __shared__ int s_len;
    const int lane = threadIdx.x & 31;
    const int n = pd->n, m = pd->m;
    uint8_t *ops = ops_base + pd->ops_off;
    if (threadIdx.x < 32) {
        uint8_t *ops_end = ops + (size_t)n + (size_t)m;
        int x = init ? n : state[0], y = init ? m : state[1], st = init ? -1 : state[2], k = init ? 0 : state[3];
        __syncwarp();
        int xl = x - r0;                                       // row inside this band's pointer block
        if (xl > 0 && y > 0) traceback_core(ptr, nb, m, cfull, ops_end, tile, lane, xl, y, st, k);
        x = xl + r0;
        if (lane == 0) {
            if (!final) {
                state[0] = x; state[1] = y; state[2] = st; state[3] = k;
            } else {
                while (y > 0) { ++k; *(ops_end - k) = 2; --y; }      // :154-158
                while (x > 0) { ++k; *(ops_end - k) = 1; --x; }      // :160-164
                *ops_len = k;
                s_len = k;
            }
        }
    }
    if (!final) return;
    __syncthreads();                    // warp 0's op bytes and s_len are read by every thread next
    const int L = s_len;
    const int shift = n + m - L;
    if (shift > 0) {
        // forward move to lower addresses: a group is read completely before any of it is written
        for (int base = 0; base < L; base += kTraceThreads * kShiftDepth) {
            uint8_t v[kShiftDepth];
#pragma unroll
            for (int u = 0; u < kShiftDepth; ++u) {
                const int q = base + u * kTraceThreads + (int)threadIdx.x;
                v[u] = 0;
                if (q < L) v[u] = __ldcg(ops + shift + q);
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kShiftDepth; ++u) {
                const int q = base + u * kTraceThreads + (int)threadIdx.x;
                if (q < L) ops[q] = v[u];
            }
            __syncthreads();
        }
    }
}

const void *long_kernel(int var, bool subst, int sym_bytes)
{
    if (sym_bytes == 2)
        return subst ? (const void *)align_long_kernel<1, 0, uint16_t> : (const void *)align_long_kernel<0, 0, uint16_t>;
    if (subst) return (const void *)align_long_kernel<1, 0, uint8_t>;
    // variants 3 / 4: the recurrences of 1 / 2 with the shorter dependent chain along a row
    // (config 5: fill 16.34 -> 16.11 ms)
    return var == 2 ? (const void *)align_long_kernel<0, 4, uint8_t>
         : var == 1 ? (const void *)align_long_kernel<0, 3, uint8_t>
                    : (const void *)align_long_kernel<0, 0, uint8_t>;
}

int long_blocks_per_sm()
{
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_long_kernel<1, 0, uint8_t>, 32, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return occ;
}

cudaError_t launch_long_col0(int4 *rec, int nb, int r0, int bg, int epoch, cudaStream_t stream)
{
    long_col0_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(rec, nb, r0, bg, epoch);
    return cudaGetLastError();
}

cudaError_t launch_long_trace(const uint8_t *ptr, const PairDesc *pd, int cfull, int r0, int nb, int init, int final,
                              int *state, uint8_t *ops_base, int *ops_len, cudaStream_t stream)
{
    trace_long_kernel<<<1, kTraceThreads, 0, stream>>>(ptr, pd, cfull, r0, nb, init, final, state, ops_base, ops_len);
    return cudaGetLastError();
}

}  // namespace tanw
