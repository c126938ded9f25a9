// tanw_pairs.cu -- instantiations of the batched page kernel (align_pairs_kernel).
#include "tanw_launch.h"

namespace tanw {

template <bool SUBST, int VAR, typename SYM, bool MULTI>
static cudaError_t go(const BatchArgs &a, const KParams &kp, int grid, cudaStream_t stream)
{
    align_pairs_kernel<SUBST, VAR, SYM, MULTI><<<grid, kWarpsPerBlock * 32, 0, stream>>>(a, kp);
    return cudaGetLastError();
}

cudaError_t launch_pairs(const BatchArgs &a, const KParams &kp, int var, bool subst, int sym_bytes, bool multi,
                         int grid, cudaStream_t stream)
{
    if (sym_bytes == 2)          // more than 256 distinct elements in a pair: rare, general recurrences only
        return subst ? go<true, 0, uint16_t, false>(a, kp, grid, stream) : go<false, 0, uint16_t, false>(a, kp, grid, stream);
    if (multi)                   // per-pair scoring systems: VAR is the most general any of them needs
        return var >= 1 ? go<false, 1, uint8_t, true>(a, kp, grid, stream) : go<false, 0, uint8_t, true>(a, kp, grid, stream);
    if (subst) {
        switch (var) {
        case 2:  return go<true, 2, uint8_t, false>(a, kp, grid, stream);
        case 1:  return go<true, 1, uint8_t, false>(a, kp, grid, stream);
        default: return go<true, 0, uint8_t, false>(a, kp, grid, stream);
        }
    }
    switch (var) {
    case 2:  return go<false, 2, uint8_t, false>(a, kp, grid, stream);
    case 1:  return go<false, 1, uint8_t, false>(a, kp, grid, stream);
    default: return go<false, 0, uint8_t, false>(a, kp, grid, stream);
    }
}

int pairs_blocks_per_sm(bool subst)
{
    int occ = 0;
    cudaError_t e = subst
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_pairs_kernel<true, 0, uint8_t, false>, kWarpsPerBlock * 32, 0)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_pairs_kernel<false, 2, uint8_t, false>, kWarpsPerBlock * 32, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return occ;
}

}  // namespace tanw
