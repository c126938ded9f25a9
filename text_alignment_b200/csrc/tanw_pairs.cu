// tanw_pairs.cu -- instantiations of the batched page kernel (align_pairs_kernel).
#include "tanw_launch.h"

namespace tanw {

template <int SUBST, int VAR, typename SYM, bool MULTI>
static cudaError_t go(const BatchArgs &a, const KParams &kp, int grid, cudaStream_t stream)
{
    align_pairs_kernel<SUBST, VAR, SYM, MULTI><<<grid, kWarpsPerBlock * 32, 0, stream>>>(a, kp);
    return cudaGetLastError();
}

// The query-profile form of a tabulated scorer: dynamic shared memory, K * 512 bytes per warp.
template <int VAR>
static cudaError_t go_profile(const BatchArgs &a, const KParams &kp, int grid, cudaStream_t stream)
{
    const int smem = kWarpsPerBlock * profile_bytes_per_warp(kp.subst_k);
    static bool configured[256] = {};                       // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 255]) {
        cudaError_t e = cudaFuncSetAttribute(align_pairs_kernel<2, VAR, uint8_t, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kWarpsPerBlock * profile_bytes_per_warp(kProfileMaxK));
        if (e != cudaSuccess) return e;
        configured[dev & 255] = true;
    }
    align_pairs_kernel<2, VAR, uint8_t, false><<<grid, kWarpsPerBlock * 32, smem, stream>>>(a, kp);
    return cudaGetLastError();
}

cudaError_t launch_pairs(const BatchArgs &a, const KParams &kp, int var, int subst, int sym_bytes, bool multi,
                         int grid, cudaStream_t stream)
{
    if (sym_bytes == 2)          // more than 256 distinct elements in a pair: rare, general recurrences only
        return subst ? go<1, 0, uint16_t, false>(a, kp, grid, stream) : go<0, 0, uint16_t, false>(a, kp, grid, stream);
    if (multi)                   // per-pair scoring systems: VAR is the most general any of them needs
        return var >= 1 ? go<0, 1, uint8_t, true>(a, kp, grid, stream) : go<0, 0, uint8_t, true>(a, kp, grid, stream);
    if (subst == 2) {
        switch (var) {
        case 2:  return go_profile<2>(a, kp, grid, stream);
        case 1:  return go_profile<1>(a, kp, grid, stream);
        default: return go_profile<0>(a, kp, grid, stream);
        }
    }
    if (subst) {
        switch (var) {
        case 2:  return go<1, 2, uint8_t, false>(a, kp, grid, stream);
        case 1:  return go<1, 1, uint8_t, false>(a, kp, grid, stream);
        default: return go<1, 0, uint8_t, false>(a, kp, grid, stream);
        }
    }
    switch (var) {
    case 2:  return go<0, 2, uint8_t, false>(a, kp, grid, stream);
    case 1:  return go<0, 1, uint8_t, false>(a, kp, grid, stream);
    default: return go<0, 0, uint8_t, false>(a, kp, grid, stream);
    }
}

int pairs_blocks_per_sm(bool subst)
{
    int occ = 0;
    cudaError_t e = subst
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_pairs_kernel<1, 0, uint8_t, false>, kWarpsPerBlock * 32, 0)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_pairs_kernel<0, 2, uint8_t, false>, kWarpsPerBlock * 32, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return occ;
}

int pairs_blocks_per_sm_profile(int subst_k)
{
    int occ = 0;
    cudaFuncSetAttribute(align_pairs_kernel<2, 0, uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kWarpsPerBlock * profile_bytes_per_warp(kProfileMaxK));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_pairs_kernel<2, 0, uint8_t, false>, kWarpsPerBlock * 32,
                                                      (size_t)kWarpsPerBlock * profile_bytes_per_warp(subst_k)) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return occ;
}

}  // namespace tanw
