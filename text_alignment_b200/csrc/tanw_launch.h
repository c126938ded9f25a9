// tanw_launch.h -- launchers of the three kernel families.  Each family is instantiated in a
// translation unit of its own (tanw_pairs.cu, tanw_lines.cu, tanw_long.cu) so that the library
// builds in parallel; tanw.cu (the C ABI) only sees these functions.
//
// `var` is the recurrence variant (tanw_kernels.cuh: 0 general, 1 gap opens <= 0, 2 gap opens <= 0
// and gap_extend_y == 0), `subst` a tabulated scorer (page kernel: 1 table lookups, 2 query profile), `sym_bytes` 1 or 2 (16-bit symbol codes run
// the general variant only), `multi` per-pair scoring systems (uint8 symbols, equality scorer).
#pragma once
#include "tanw_kernels.cuh"

namespace tanw {

cudaError_t launch_pairs(const BatchArgs &a, const KParams &kp, int var, int subst, int sym_bytes, bool multi,
                         int grid, cudaStream_t stream);
int pairs_blocks_per_sm(bool subst);
int pairs_blocks_per_sm_profile(int subst_k);    // subst == 2: the query-profile form (K <= kProfileMaxK, |score| <= 127)

cudaError_t launch_lines(const LineArgs &a, const KParams &kp, int var, bool subst, int grid, cudaStream_t stream);
int lines_blocks_per_sm();
// the .u16x2 line kernel (tanw_lines16.cuh): var 1 or 2, equality scorer, match >= mismatch
cudaError_t launch_lines16(const LineArgs &a, const KParams &kp, int var, int grid, cudaStream_t stream);
int lines16_blocks_per_sm();

const void *long_kernel(int var, bool subst, int sym_bytes);
int long_blocks_per_sm();
cudaError_t launch_long_col0(int4 *rec, int nb, int r0, int bg, int epoch, cudaStream_t stream);
cudaError_t launch_long_trace(const uint8_t *ptr, const PairDesc *pd, int cfull, int r0, int nb, int init, int final,
                              int *state, uint8_t *ops_base, int *ops_len, cudaStream_t stream);

}  // namespace tanw
