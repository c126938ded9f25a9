// tanw_lines.cu -- instantiations of the four-pairs-per-warp line kernel (align_lines_kernel).
#include "tanw_launch.h"

#include <algorithm>
#include "tanw_lines16.cuh"

namespace tanw {

template <int SUBST, int VAR>
static cudaError_t go(const LineArgs &a, const KParams &kp, int grid, cudaStream_t stream)
{
    align_lines_kernel<SUBST, VAR><<<grid, kWarpsPerBlock * 32, 0, stream>>>(a, kp);
    return cudaGetLastError();
}

cudaError_t launch_lines(const LineArgs &a, const KParams &kp, int var, bool subst, int grid, cudaStream_t stream)
{
    if (subst) {
        switch (var) {
        case 2:  return go<1, 2>(a, kp, grid, stream);
        case 1:  return go<1, 1>(a, kp, grid, stream);
        default: return go<1, 0>(a, kp, grid, stream);
        }
    }
    switch (var) {
    case 2:  return go<0, 2>(a, kp, grid, stream);
    case 1:  return go<0, 1>(a, kp, grid, stream);
    default: return go<0, 0>(a, kp, grid, stream);
    }
}

cudaError_t launch_lines16(const LineArgs &a, const KParams &kp, int var, int grid, cudaStream_t stream)
{
    if (var == 2) align_lines16_kernel<2><<<grid, kL16Warps * 32, 0, stream>>>(a, kp);
    else          align_lines16_kernel<1><<<grid, kL16Warps * 32, 0, stream>>>(a, kp);
    return cudaGetLastError();
}

int lines16_blocks_per_sm()
{
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_lines16_kernel<1>, kL16Warps * 32, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return occ;
}

int lines_blocks_per_sm()
{
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_lines_kernel<1, 0>, kWarpsPerBlock * 32, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return occ;
}

}  // namespace tanw
