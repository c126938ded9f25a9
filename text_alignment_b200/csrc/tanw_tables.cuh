// tanw_tables.cuh -- the batch tables, built on the device.
//
// The reference aligns one page per call (alignToOCR.py:273); a batch call brings 10^4..10^6
// pairs as four plain arrays (t_off, n, o_off, m; include/tanw.h).  Everything the align kernels
// need besides the symbols is derived from those arrays HERE, on the device, so that the host
// does no per-pair work on the way in (round 1 walked the pair table and counting-sorted it on
// one host core: 1.6 ms per 125 000 line pairs, more than their 0.95 ms of alignment):
//
//   survey_kernel   per-pair validation and routing (line / page / chained-stripe kernel), sizes
//                   of the scratch the batch needs, per-tile sums of n+m, for up to 64 slices of the
//                   batch (the host merges slices into the chunks it pipelines); its 1 KB result
//                   is the only thing the host waits for before it launches;
//   build_kernel    exclusive prefix sums of n+m (the canonical op-buffer layout), the pair
//                   descriptors, histograms of the two work-order keys;
//   bins_kernel     prefix sums over the histogram bins, strip-width classes of the line kernel;
//   scatter_kernel  counting-sort scatter: page pairs largest first (greedy longest-processing-
//                   time order for the persistent warps), line pairs by (strip width, height) so
//                   that the four pairs a warp aligns together are alike.
//
// The order of pairs with equal keys depends on atomics; it only moves work between warps, every
// pair's result is a function of the pair alone.
#pragma once
#include "tanw_kernels.cuh"

namespace tanw {

constexpr int kTile = 2048;                        // pairs per block of survey_kernel / build_kernel
constexpr int kTileThreads = 256;
constexpr int kTilePer = kTile / kTileThreads;     // consecutive pairs per thread
constexpr int kMaxChunks = 16;                     // a batch call is pipelined in up to this many chunks ...
constexpr int kMaxSlices = 64;                     // ... each made of slices the survey reports on
constexpr int kLineNKeys = 1024;                   // line sort keys: heights beyond this share the last key
constexpr int kLineKeys = 4 * kLineNKeys;          // ... x 4 strip-width classes
constexpr int kPageKeys = 4096;                    // page sort keys: n*m quantised to 12 bits
constexpr int kMaxLongList = 4096;                 // chained-stripe pairs listed for the host per batch

enum Route : int { kRoutePage = 0, kRouteLine = 1, kRouteLong = 2, kRouteLine16 = 3 };

// What the host learns about one chunk of the batch (pairs [first, first + count)).
struct ChunkSurvey {
    long long cells;            // sum n*m
    long long page_cells;       // ... of its page pairs alone
    long long cap;              // sum n+m: bytes of the chunk in the canonical op layout
    long long sym_end;          // largest symbol offset any of its pairs reads, exclusive
    long long max_slot;         // largest ptr_bytes() among its page pairs
    long long max_line_slot;    // largest line_ptr_bytes() among its line pairs
    long long max_page_cells;   // largest n*m among its page pairs
    long long max_line16_slot;  // largest line_ptr_bytes() among its pairs of the 16-bit line kernel
    int max_n_page;             // tallest page pair (boundary array rows)
    int n_page, n_line, n_long, n_line16;
    int line_class[4];          // line pairs per strip-width class (cell-less pairs count as class 0)
    int line16_class[4];        // the same for the pairs of the 16-bit line kernel
    int max_nm_line16;          // longest possible op string (n + m) among them
    int max_n_line16;           // tallest of them
};

struct Survey {
    unsigned long long bad;     // 0: every pair is valid; else ULLONG_MAX - (smallest invalid pair index)
    int max_nm;                 // largest n+m (range check of the fixed-point scores)
    int n_long;                 // entries in long_list (may exceed kMaxLongList: then the batch is refused)
    ChunkSurvey chunk[kMaxSlices];
    int long_list[kMaxLongList];
};

struct TableArgs {
    const long long *t_off, *o_off;
    const int *n, *m;
    long long n_pairs, symbols_len;
    long long chunk_pairs;      // pairs per chunk, a multiple of kTile
    // routing (mirrors tanw_batch_prepare of round 1)
    long long long_cells;       // n*m >= this: chained stripes
    long long slot_limit;       // a page pair whose pointer bytes * warps per block exceed this: chained stripes
    int use_lines, wide, tiny_batch, can_long;
    int line16_max_n;           // line pairs up to this height run two per register (tanw_lines16.cuh); 0: none
    // outputs
    Survey *survey;
    long long *tile_sums;       // per tile: sum of n+m
    PairDesc *pairs;
    unsigned char *route;
    int *order;                 // page pairs of chunk c at order[c * chunk_pairs ...], largest first
    int *line_sorted;           // line pairs of chunk c at line_sorted[c * chunk_pairs ...]
    int *hist;                  // per chunk: kLineKeys bins of the 16-bit line kernel, kLineKeys line bins, kPageKeys page bins
    LineClasses *classes;       // per chunk: [2c] 16-bit line kernel (octets), [2c + 1] line kernel (quads)
};

constexpr int kHistStride = 2 * kLineKeys + kPageKeys;

__host__ __device__ __forceinline__ int route_of(const TableArgs &a, long long np, long long mp)
{
    const bool oversize = np > 0 && mp > 0 &&
                          (ptr_bytes((int)np, (int)mp) + 255) / 256 * 256 * kWarpsPerBlock > a.slot_limit &&
                          !(!a.wide && a.use_lines && mp <= kLineMaxM && np <= kLineMaxN);
    const bool tiny = a.tiny_batch && mp > kLineMaxM && np * mp >= (1ll << 16);
    if ((np * mp >= a.long_cells || tiny || oversize) && a.can_long) return kRouteLong;
    if (!a.wide && a.use_lines && mp <= kLineMaxM && np <= kLineMaxN) return (a.line16_max_n > 0 && np <= a.line16_max_n) ? kRouteLine16 : kRouteLine;
    return kRoutePage;
}

__device__ __forceinline__ int line_key(int np, int mp)
{
    const bool act = np > 0 && mp > 0;
    return kLineKeys - 1 - ((act ? line_c(mp) / 4 - 1 : 0) * kLineNKeys + (act ? min(np, kLineNKeys - 1) : 0));
}

__device__ __forceinline__ long long block_sum(long long v, long long *sh)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(kFull, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < kTileThreads / 32; ++w) t += sh[w];
    return t;
}

__device__ __forceinline__ void atomic_max_ll(long long *p, long long v)
{
    if (v > 0) atomicMax(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
}

__global__ void __launch_bounds__(kTileThreads) survey_kernel(const TableArgs a)
{
    __shared__ long long sh[kTileThreads / 32];
    const long long base = (long long)blockIdx.x * kTile + (long long)threadIdx.x * kTilePer;
    const int chunk = (int)(((long long)blockIdx.x * kTile) / a.chunk_pairs);
    long long cap = 0, cells = 0, page_cells = 0, sym_end = 0, max_slot = 0, max_line = 0, max_line16 = 0, max_pc = 0, bad = -1;
    int max_n = 0, max_nm = 0, max_nm16 = 0, max_n16 = 0, n_page = 0, n_line = 0, n_long = 0, n_line16 = 0, cls[4] = {0, 0, 0, 0}, cls16[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < kTilePer; ++k) {
        const long long p = base + k;
        if (p >= a.n_pairs) break;
        const long long np = a.n[p], mp = a.m[p], to = a.t_off[p], oo = a.o_off[p];
        if (np < 0 || mp < 0 || to < 0 || oo < 0 || to + np > a.symbols_len || oo + mp > a.symbols_len) {
            if (bad < 0) bad = p;          // the host looks at the pair again to say what is wrong with it
            continue;
        }
        cap += np + mp;
        cells += np * mp;
        sym_end = max(sym_end, max(to + np, oo + mp));
        max_nm = max(max_nm, (int)min(np + mp, 0x7fffffffll));
        const int r = route_of(a, np, mp);
        if (r == kRouteLong) {
            ++n_long;
            const int at = atomicAdd(&a.survey->n_long, 1);
            if (at < kMaxLongList) a.survey->long_list[at] = (int)p;
        } else if (r == kRouteLine) {
            ++n_line;
            max_line = max(max_line, line_ptr_bytes((int)np, (int)mp));
            ++cls[(np > 0 && mp > 0) ? line_c((int)mp) / 4 - 1 : 0];
        } else if (r == kRouteLine16) {
            ++n_line16;
            max_nm16 = max(max_nm16, (int)(np + mp));
            max_n16 = max(max_n16, (int)np);
            max_line16 = max(max_line16, line_ptr_bytes((int)np, (int)mp));
            ++cls16[(np > 0 && mp > 0) ? line_c((int)mp) / 4 - 1 : 0];
        } else {
            ++n_page;
            page_cells += np * mp;
            max_slot = max(max_slot, ptr_bytes((int)np, (int)mp));
            max_pc = max(max_pc, np * mp);
            max_n = max(max_n, (int)np);
        }
    }
    const long long tile_cap = block_sum(cap, sh);
    if (threadIdx.x == 0) a.tile_sums[blockIdx.x] = tile_cap;
    // per-thread results -> the chunk's record; warp-reduce first to keep the atomics few
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cells += __shfl_down_sync(kFull, cells, d);
        page_cells += __shfl_down_sync(kFull, page_cells, d);
        sym_end = max(sym_end, __shfl_down_sync(kFull, sym_end, d));
        max_slot = max(max_slot, __shfl_down_sync(kFull, max_slot, d));
        max_line = max(max_line, __shfl_down_sync(kFull, max_line, d));
        max_line16 = max(max_line16, __shfl_down_sync(kFull, max_line16, d));
        max_pc = max(max_pc, __shfl_down_sync(kFull, max_pc, d));
        max_n = max(max_n, __shfl_down_sync(kFull, max_n, d));
        max_nm = max(max_nm, __shfl_down_sync(kFull, max_nm, d));
        max_nm16 = max(max_nm16, __shfl_down_sync(kFull, max_nm16, d));
        max_n16 = max(max_n16, __shfl_down_sync(kFull, max_n16, d));
        n_page += __shfl_down_sync(kFull, n_page, d);
        n_line += __shfl_down_sync(kFull, n_line, d);
        n_long += __shfl_down_sync(kFull, n_long, d);
        n_line16 += __shfl_down_sync(kFull, n_line16, d);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cls[c] += __shfl_down_sync(kFull, cls[c], d);
            cls16[c] += __shfl_down_sync(kFull, cls16[c], d);
        }
    }
    if (bad >= 0) atomicMax(&a.survey->bad, 0xFFFFFFFFFFFFFFFFull - (unsigned long long)bad);
    if ((threadIdx.x & 31) == 0) {
        ChunkSurvey *cs = &a.survey->chunk[chunk];
        atomicAdd(reinterpret_cast<unsigned long long *>(&cs->cells), (unsigned long long)cells);
        if (page_cells) atomicAdd(reinterpret_cast<unsigned long long *>(&cs->page_cells), (unsigned long long)page_cells);
        atomic_max_ll(&cs->sym_end, sym_end);
        atomic_max_ll(&cs->max_slot, max_slot);
        atomic_max_ll(&cs->max_line_slot, max_line);
        atomic_max_ll(&cs->max_line16_slot, max_line16);
        atomic_max_ll(&cs->max_page_cells, max_pc);
        atomicMax(&cs->max_n_page, max_n);
        atomicMax(&a.survey->max_nm, max_nm);
        atomicMax(&cs->max_nm_line16, max_nm16);
        atomicMax(&cs->max_n_line16, max_n16);
        if (n_page) atomicAdd(&cs->n_page, n_page);
        if (n_line) atomicAdd(&cs->n_line, n_line);
        if (n_long) atomicAdd(&cs->n_long, n_long);
        if (n_line16) atomicAdd(&cs->n_line16, n_line16);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (cls[c]) atomicAdd(&cs->line_class[c], cls[c]);
            if (cls16[c]) atomicAdd(&cs->line16_class[c], cls16[c]);
        }
    }
    if (threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long *>(&a.survey->chunk[chunk].cap), (unsigned long long)tile_cap);
}

// Pair descriptors with their canonical op offsets, routes, and the key histograms of a chunk.
__global__ void __launch_bounds__(kTileThreads) build_kernel(const TableArgs a, int chunk, int shift)
{
    __shared__ long long sh[kTileThreads / 32];
    __shared__ long long warp_base[kTileThreads / 32];
    const long long tile0 = (long long)chunk * (a.chunk_pairs / kTile);
    const long long tile = tile0 + blockIdx.x;
    // canonical offset of the tile's first pair: everything before it in the whole batch
    long long part = 0;
    for (long long t = threadIdx.x; t < tile; t += kTileThreads) part += a.tile_sums[t];
    const long long tile_base = block_sum(part, sh);

    const long long base = tile * kTile + (long long)threadIdx.x * kTilePer;
    int np[kTilePer], mp[kTilePer];
    long long mine = 0;
#pragma unroll
    for (int k = 0; k < kTilePer; ++k) {
        const long long p = base + k;
        np[k] = (p < a.n_pairs) ? a.n[p] : 0;
        mp[k] = (p < a.n_pairs) ? a.m[p] : 0;
        mine += (long long)np[k] + mp[k];
    }
    // exclusive scan over the block's threads
    long long incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long v = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += v;
    }
    __syncthreads();
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int w = 0; w < kTileThreads / 32; ++w) { warp_base[w] = run; run += sh[w]; }
    }
    __syncthreads();
    long long off = tile_base + warp_base[warp] + incl - mine;

    int *hist = a.hist + (size_t)chunk * kHistStride;
#pragma unroll
    for (int k = 0; k < kTilePer; ++k) {
        const long long p = base + k;
        if (p >= a.n_pairs) break;
        PairDesc pd;
        pd.t_off = a.t_off[p]; pd.o_off = a.o_off[p]; pd.ops_off = off; pd.n = np[k]; pd.m = mp[k];
        a.pairs[p] = pd;
        off += (long long)np[k] + mp[k];
        const int r = route_of(a, np[k], mp[k]);
        a.route[p] = (unsigned char)r;
        if (r == kRouteLine16) atomicAdd(hist + line_key(np[k], mp[k]), 1);
        else if (r == kRouteLine) atomicAdd(hist + kLineKeys + line_key(np[k], mp[k]), 1);
        else if (r == kRoutePage)
            atomicAdd(hist + 2 * kLineKeys + (kPageKeys - 1 - (int)(((long long)np[k] * mp[k]) >> shift)), 1);
    }
}

// Exclusive prefix sums over the line bins (16-bit kernel first, then the int32 line kernel: one
// sorted list) and over the page bins of a chunk (one block), and the strip-width classes of its
// line pairs.  Afterwards hist[key] / hist[kLineKeys + key] is the first slot of `key` in the
// sorted line list, hist[2 * kLineKeys + key] among the page pairs.
__global__ void __launch_bounds__(1024) bins_kernel(const TableArgs a, int chunk, int4 line16_class, int4 line_class)
{
    __shared__ int sh[32];
    __shared__ int carry;
    int *hist = a.hist + (size_t)chunk * kHistStride;
    for (int part = 0; part < 2; ++part) {
        int *h = hist + (part == 0 ? 0 : 2 * kLineKeys);
        const int nb = part == 0 ? 2 * kLineKeys : kPageKeys;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int b0 = 0; b0 < nb; b0 += 1024 * 4) {
            int v[4], s = 0;
            const int at = b0 + threadIdx.x * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) { v[k] = (at + k < nb) ? h[at + k] : 0; s += v[k]; }
            int incl = s;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) sh[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                int w = sh[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(kFull, w, d);
                    if (lane >= d) w += t;
                }
                sh[lane] = w;
            }
            __syncthreads();
            int run = carry + (warp ? sh[warp - 1] : 0) + incl - s;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (at + k < nb) h[at + k] = run;
                run += v[k];
            }
            __syncthreads();
            if (threadIdx.x == 1023) carry = run;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        // line pairs are sorted by descending (class, n): class 3 first
        int at = 0;
        for (int kind = 0; kind < 2; ++kind) {           // 0: octets of the 16-bit kernel, 1: quads
            const int4 cl = kind == 0 ? line16_class : line_class;
            const int count[4] = { cl.x, cl.y, cl.z, cl.w };
            const int per = kind == 0 ? 8 : 4;
            LineClasses lc;
            int units = 0;
            for (int c = 3; c >= 0; --c) {
                lc.start[c] = at;
                lc.count[c] = count[c];
                lc.quad0[c] = units;
                at += count[c];
                units += (count[c] + per - 1) / per;
            }
            lc.n_quads = units;
            a.classes[2 * chunk + kind] = lc;
        }
    }
}

__global__ void __launch_bounds__(kTileThreads) scatter_kernel(const TableArgs a, int chunk, int shift)
{
    const long long first = (long long)chunk * a.chunk_pairs;
    const long long p = first + (long long)blockIdx.x * kTileThreads + threadIdx.x;
    const long long end = min(a.n_pairs, first + a.chunk_pairs);
    if (p >= end) return;
    int *hist = a.hist + (size_t)chunk * kHistStride;
    const int r = a.route[p];
    const int np = a.n[p], mp = a.m[p];
    if (r == kRouteLine || r == kRouteLine16) {
        const int at = atomicAdd(hist + (r == kRouteLine ? kLineKeys : 0) + line_key(np, mp), 1);
        a.line_sorted[first + at] = (int)p;
    } else if (r == kRoutePage) {
        const int at = atomicAdd(hist + 2 * kLineKeys + (kPageKeys - 1 - (int)(((long long)np * mp) >> shift)), 1);
        a.order[first + at] = (int)p;
    }
}

// Packed op strings (tanw_set_packed_ops): an op is 0, 1 or 2, so four of them fit a byte.  Pair p's
// packed string starts at byte (ops_off[p] >> 2) + p of the packed buffer -- a fresh byte for every
// pair, no prefix sum needed (ceil((n+m)/4) <= (ops_off[p+1] >> 2) + 1 - (ops_off[p] >> 2)) -- op q in
// bits 2*(q & 3) of byte q >> 2.  One warp per pair of the chunk whose route is in `routes` (bit r set
// = route r), right after the kernels that aligned them, on their stream: a quarter of the bytes go
// over PCIe (config 3 on 8 GPUs is bound by exactly that traffic).
__global__ void __launch_bounds__(256) pack_ops_kernel(const PairDesc *pairs, const unsigned char *route, const int *ops_len,
                                                       const uint8_t *ops, uint8_t *packed, long long first, long long count,
                                                       unsigned routes)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = warp; k < count; k += warps) {
        const long long p = first + k;
        if (!((routes >> route[p]) & 1u)) continue;
        const long long off = pairs[p].ops_off;
        const int L = ops_len[p];
        const uint8_t *src = ops + off;
        uint8_t *dst = packed + (off >> 2) + p;
        for (int j = lane; 4 * j < L; j += 32) {
            unsigned b = src[4 * j];
            if (4 * j + 1 < L) b |= (unsigned)src[4 * j + 1] << 2;
            if (4 * j + 2 < L) b |= (unsigned)src[4 * j + 2] << 4;
            if (4 * j + 3 < L) b |= (unsigned)src[4 * j + 3] << 6;
            dst[j] = (uint8_t)b;
        }
    }
}

// The largest symbol code of the batch (tabulated scorers index a K x K table with them).
template <typename SYM>
__global__ void max_symbol_kernel(const SYM *sym, long long count, int *out)
{
    int v = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        v = max(v, (int)sym[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_down_sync(kFull, v, d));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

}  // namespace tanw
