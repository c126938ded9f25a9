// tanw_kernels.cuh -- device code of libtanw.so (sm_100a only).
//
// Affine-gap Needleman-Wunsch in the exact form of DDMAL/text_alignment
// (/root/reference/textSeqCompare.py:53-88 fill, :96-170 traceback), re-designed for B200:
//
//   * one warp aligns one (transcript, OCR) pair at a time; warps are persistent and pull
//     pairs (largest first) off a global counter;
//   * the OCR axis (columns j) is cut into passes of 32*C columns; inside a pass lane l owns
//     the C-column strip [j0 + l*C, j0 + (l+1)*C) and walks down the transcript rows, one row
//     per step, skewed by one step per lane (anti-diagonal wavefront).  The strip's right-edge
//     values travel to lane l+1 with two __shfl_up_sync per step;
//   * all H/E/F state (here M/X/Y) lives in registers: per column W = max(M,Y), X and
//     D = max(M,X,Y) of the row above, per step the running Q = max(M,X) and Y of the column
//     to the left;
//   * scores are carried in int32 fixed point, value*64, and the low six bits hold the origin
//     tag of a value (M = 0b101010, X = 0b010101, Y = 0).  A plain integer max over tagged
//     candidates therefore returns the maximum AND, on equal values, the candidate that comes
//     first in the reference's list order (M, X, Y) -- list.index(max(list)),
//     textSeqCompare.py:72,:80,:88 -- without any compare/select for the argmax;
//   * the three 2-bit traceback pointers of a cell are cut out of the three raw max results
//     with two bit-selects and written as one byte per cell, step-major so that every warp
//     store is one contiguous 32*C byte segment;
//   * the gap-extension additions are folded into per-row / per-column offsets
//     (X^ = X - ex*i, Y^ = Y - ey*k) so each recurrence is a single VIADDMNMX;
//   * the traceback runs in the same kernel right after the pair's fill, on the GPU, and
//     emits the op string.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tanw {

constexpr int      kShift   = 6;                 // fixed point: value << 6
constexpr int      kTagM    = 0x2A;              // 10 10 10 : "came from M" in all three fields
constexpr int      kTagX    = 0x15;              // 01 01 01 : "came from X"
constexpr int      kTagMask = 0x3F;              //            "came from Y" is 0
constexpr int      kNeg     = -(1 << 30);        // stands in for -1e100 (textSeqCompare.py:55,:60)
constexpr int      kMaxC    = 32;                // widest strip (columns per lane)
constexpr int      kPassW   = 32 * kMaxC;        // columns of a full pass
constexpr int      kWarpsPerBlock = 4;
constexpr unsigned kFull    = 0xFFFFFFFFu;

struct KParams {
    int maT, miT;            // (match<<6)|kTagM, (mismatch<<6)|kTagM
    int ox, ex, oy, ey;      // (gox+gex)<<6, gex<<6, (goy+gey)<<6, gey<<6
    int bg;                  // boundary_gap<<6 (module-level gap_extend, textSeqCompare.py:9)
    int subst_k;
    const int *subst;        // device table, entry = (score<<6)|kTagM, or nullptr
    int cy[kMaxC];           // oy - ey*k
    int ye[kMaxC];           // ey*k
};

struct PairDesc {
    long long t_off, o_off;  // into the symbol buffer
    long long ops_off;       // into the device op buffer (capacity n+m)
    int n, m;
};

struct BatchArgs {
    const uint8_t  *sym;
    const PairDesc *pairs;
    const int      *order;       // pair indices, largest first
    unsigned       *counter;     // work counter
    int             n_pairs;
    uint8_t        *ptr_arena;   // slot_bytes per warp slot
    long long       slot_bytes;
    int2           *bnd_arena;   // bnd_rows int2 per warp slot
    int             bnd_rows;
    uint8_t        *ops;         // out
    int            *ops_len;     // out
    int            *scores;      // out, 3 per pair (may be null)
};

// Width of the remainder pass: smallest multiple of 4 columns per lane covering r columns.
__host__ __device__ inline int remainder_c(int r) { return ((r + 127) / 128) * 4; }

// Bytes of traceback pointers of one pair (all passes, (n+32) step slots each).
__host__ __device__ inline long long ptr_bytes(int n, int m)
{
    if (n <= 0 || m <= 0) return 0;
    long long steps = (long long)n + 32;
    int nfull = m / kPassW, r = m % kPassW;
    return steps * 32 * ((long long)nfull * kMaxC + (r ? remainder_c(r) : 0));
}

template <int C>
struct Strip {
    int      W[C];        // max(M|tagM, Y) of the row above, real value
    int      Xh[C];       // X^ = (X|tagX) - ex*i of the row above
    int      D[C];        // max(M, X, Y) tagged, of the row above
    unsigned ow[C / 4];   // the strip's OCR symbols, four per register
};

// One row of one strip: C cells.  See the file header for the value encoding.
//   q_in, y_in : Q = max(M,X) tagged and Y (clean) of the cell left of the strip, same row
//   dul_in     : D of the cell up-left of the strip's first cell
// Returns the strip's right edge in q_out / y_out and the C pointer bytes in pw[C/4].
template <int C, bool FINAL, bool SUBST>
__device__ __forceinline__ void strip_row(Strip<C> &s, const KParams &kp, int tch, int i,
                                          int q_in, int y_in, int dul_in,
                                          int &q_out, int &y_out, unsigned (&pw)[C / 4],
                                          int kfin, int (&cap)[3])
{
    const int xe = kp.ex * i;                 // X = X^ + xe
    const int cx = kp.ox - xe;                // W + ox - ex*i
    const unsigned trep = (unsigned)tch * 0x01010101u;
    const int *srow = SUBST ? kp.subst + tch * kp.subst_k : nullptr;
    int q = q_in;
    int yh = y_in + kp.ey;                    // Y^ of column "-1" of the strip
    int dul = dul_in;
    unsigned bytes[4];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        int sc;
        if (SUBST) {
            unsigned oc = (s.ow[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            sc = __ldg(srow + oc);
        } else {
            unsigned x = s.ow[k >> 2] ^ trep;
            sc = ((x & (0xFFu << (8 * (k & 3)))) == 0u) ? kp.maT : kp.miT;   // :31-32
        }
        // M[i][j] = max(M,X,Y)[i-1][j-1] + score, tagged as an M value   (:70-72)
        const int m2 = (dul & ~kTagMask) + sc;
        // X[i][j] = max(M[i-1][j]+ox, X[i-1][j]+ex, Y[i-1][j]+ox)         (:83-88)
        const int xraw = __viaddmax_s32(s.W[k], cx, s.Xh[k]);
        const int xh = (xraw & ~kTagMask) | kTagX;
        // Y[i][j] = max(M[i][j-1]+oy, X[i][j-1]+oy, Y[i][j-1]+ey)         (:75-80)
        const int yraw = __viaddmax_s32(q, kp.cy[k], yh);
        const int yc = yraw & ~kTagMask;
        const int w  = __viaddmax_s32(yc, kp.ye[k], m2);     // max(M, Y)
        const int qn = __viaddmax_s32(xh, xe, m2);           // max(M, X)
        const int dn = __viaddmax_s32(xh, xe, w);            // max(M, X, Y)
        // pointer byte: bits 0-1 from D(up-left), 2-3 from xraw, 4-5 (and garbage 6-7) from yraw
        const unsigned t2 = ((unsigned)xraw & 0x0Cu) | ((unsigned)yraw & ~0x0Cu);
        bytes[k & 3] = ((unsigned)dul & 0x03u) | (t2 & ~0x03u);
        if (FINAL) {
            if (k == kfin) { cap[0] = m2; cap[1] = xh + xe; cap[2] = yc + kp.ye[k]; }
        }
        dul = s.D[k];
        s.W[k] = w; s.Xh[k] = xh; s.D[k] = dn;
        q = qn; yh = yc;
        if ((k & 3) == 3) {
            const unsigned lo = __byte_perm(bytes[0], bytes[1], 0x0040);
            const unsigned hi = __byte_perm(bytes[2], bytes[3], 0x0040);
            pw[k >> 2] = __byte_perm(lo, hi, 0x5410);
        }
    }
    q_out = q;
    y_out = yh + kp.ye[C - 1];
}

template <int C>
__device__ __forceinline__ void store_ptr_words(uint8_t *dst, const unsigned (&pw)[C / 4])
{
    if (C % 16 == 0) {
#pragma unroll
        for (int v = 0; v < C / 16; ++v)
            __stcs(reinterpret_cast<uint4 *>(dst) + v,
                   make_uint4(pw[4 * v], pw[4 * v + 1], pw[4 * v + 2], pw[4 * v + 3]));
    } else if (C % 8 == 0) {
#pragma unroll
        for (int v = 0; v < C / 8; ++v)
            __stcs(reinterpret_cast<uint2 *>(dst) + v, make_uint2(pw[2 * v], pw[2 * v + 1]));
    } else {
#pragma unroll
        for (int v = 0; v < C / 4; ++v)
            __stcs(reinterpret_cast<unsigned *>(dst) + v, pw[v]);
    }
}

// One pass: columns [j0, j0 + 32*C) of one pair, all n rows.
//   first    : j0 == 0 (left boundary is column 0 of the matrices, textSeqCompare.py:53-56)
//   has_next : another pass follows; lane 31 leaves its right edge in bnd[1..n]
//   ptr      : base of this pass's pointer bytes, laid out [step t][lane][C]
//   fin_lane, fin_k : where column m lives in this pass (or fin_lane = -1)
template <int C, bool SUBST>
__device__ __forceinline__ void fill_pass(const KParams &kp, const uint8_t *__restrict__ T,
                                       const uint8_t *__restrict__ O, int n, int m, int j0,
                                       bool first, bool has_next, int2 *bnd,
                                       uint8_t *__restrict__ ptr, int fin_lane, int fin_k,
                                       int (&cap)[3])
{
    const int lane = threadIdx.x & 31;
    const int c0 = j0 + lane * C;                 // 0-based first column of the strip
    Strip<C> s;
#pragma unroll
    for (int w = 0; w < C / 4; ++w) {
        unsigned v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = c0 + 4 * w + b;
            const unsigned ch = (c < m) ? (unsigned)__ldg(O + c) : 0xFFu;
            v |= ch << (8 * b);
        }
        s.ow[w] = v;
    }
    // row 0: M[0][j] = X[0][j] = bg*j, Y[0][j] = -inf   (textSeqCompare.py:57-60)
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int base = kp.bg * (c0 + k + 1);
        s.W[k] = base | kTagM;
        s.Xh[k] = base | kTagX;
        s.D[k] = base | kTagM;
    }
    int q_out = (kp.bg * (c0 + C)) | kTagM;       // right edge of row 0
    int y_out = kNeg;
    int q_prev = (kp.bg * c0) | kTagM;            // left neighbour column, row 0
    int y_prev = kNeg;                            // (Y[0][j] = -inf, also at j = 0)
    int2 bnext = make_int2(0, 0);
    if (!first && lane == 0 && n >= 1) bnext = __ldcg(bnd + 1);
    int tnext = (lane == 0 && n >= 1) ? (int)__ldg(T) : 0;   // symbol of row i+1 for this lane

    const int last_step = n + 31;
    for (int t = 1; t <= last_step; ++t) {
        const int i = t - lane;                   // this lane's row in this step
        int q_in = __shfl_up_sync(kFull, q_out, 1);
        int y_in = __shfl_up_sync(kFull, y_out, 1);
        if (lane == 0) {
            if (first) {                          // column 0: M = Y = bg*i, X = -inf (:54-56)
                q_in = (kp.bg * i) | kTagM;
                y_in = kp.bg * i;
            } else {
                q_in = bnext.x;
                y_in = bnext.y;
                if (t + 1 <= n) bnext = __ldcg(bnd + t + 1);
            }
        }
        const int dul_in = max(q_prev, y_prev);   // D of (i-1, left neighbour column)
        const int tch = tnext;
        if (i >= 0 && i < n) tnext = (int)__ldg(T + i);       // row i+1 reads T[i]
        if (i >= 1 && i <= n) {
            unsigned pw[C / 4];
            if (t < n) {                          // no lane is on the last row yet
                strip_row<C, false, SUBST>(s, kp, tch, i, q_in, y_in, dul_in, q_out, y_out, pw, -1, cap);
            } else {
                const int kfin = (i == n && lane == fin_lane) ? fin_k : -1;
                strip_row<C, true, SUBST>(s, kp, tch, i, q_in, y_in, dul_in, q_out, y_out, pw, kfin, cap);
            }
            store_ptr_words<C>(ptr + ((size_t)t * 32 + lane) * C, pw);
            if (has_next && lane == 31) __stcg(bnd + i, make_int2(q_out, y_out));
        }
        q_prev = q_in;
        y_prev = y_in;
    }
}

template <bool SUBST>
__device__ __forceinline__ void dispatch_pass(int C, const KParams &kp, const uint8_t *T,
                                              const uint8_t *O, int n, int m, int j0, bool first,
                                              bool has_next, int2 *bnd, uint8_t *ptr,
                                              int fin_lane, int fin_k, int (&cap)[3])
{
    switch (C) {
    case 4:  fill_pass<4,  SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    case 8:  fill_pass<8,  SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    case 12: fill_pass<12, SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    case 16: fill_pass<16, SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    case 20: fill_pass<20, SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    case 24: fill_pass<24, SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    case 28: fill_pass<28, SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    default: fill_pass<32, SUBST>(kp, T, O, n, m, j0, first, has_next, bnd, ptr, fin_lane, fin_k, cap); break;
    }
}

// Address of the pointer byte of cell (i, j), 1-based, inside a pair's pointer block.
struct PtrMap {
    long long steps;     // n + 32
    int nfull, cr;       // full passes, strip width of the remainder pass
    __device__ __forceinline__ PtrMap(int n, int m)
    {
        steps = (long long)n + 32;
        nfull = m / kPassW;
        const int r = m % kPassW;
        cr = r ? remainder_c(r) : 0;
    }
    __device__ __forceinline__ long long offset(int i, int j) const
    {
        const int c = j - 1;
        int p = c / kPassW;
        int C = kMaxC;
        if (p >= nfull) { p = nfull; C = cr; }
        const int cc = c - p * kPassW;
        const int lane = cc / C, k = cc - lane * C;
        const long long base = (long long)p * steps * kPassW;
        return base + ((long long)(i + lane) * 32 + lane) * C + k;
    }
};

// Traceback of one pair by one lane (textSeqCompare.py:96-164), ops written back to front at
// the END of the pair's op buffer (capacity n+m); returns the number of columns.
__device__ __forceinline__ int traceback_lane(const uint8_t *ptr, int n, int m, uint8_t *ops_end)
{
    int x = n, y = m, k = 0;
    if (n > 0 && m > 0) {
        const PtrMap map(n, m);
        unsigned b = __ldcg(ptr + map.offset(x, y));
        int st = 2 - (int)(b & 3u);                      // mpt = mat_ptr[n][m]            (:102)
        while (true) {
            int op;
            if (st == 0)      { op = 0; st = 2 - (int)(b & 3u);        --x; --y; }   // :115-125
            else if (st == 1) { op = 1; st = 2 - (int)((b >> 2) & 3u); --x; }        // :128-135
            else              { op = 2; st = 2 - (int)((b >> 4) & 3u); --y; }        // :138-145
            ++k;
            *(ops_end - k) = (uint8_t)op;
            if (x <= 0 || y <= 0) break;
            b = __ldcg(ptr + map.offset(x, y));
        }
    }
    while (y > 0) { ++k; *(ops_end - k) = 2; --y; }      // OCR remainder first          (:154-158)
    while (x > 0) { ++k; *(ops_end - k) = 1; --x; }      // then transcript remainder    (:160-164)
    return k;
}

__device__ __forceinline__ int score_out(int v)
{
    return (v <= kNeg / 2) ? kNeg : (v >> kShift);
}

template <bool SUBST>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
align_pairs_kernel(const BatchArgs a, const __grid_constant__ KParams kp)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    uint8_t *const ptr = a.ptr_arena + (size_t)slot * (size_t)a.slot_bytes;
    int2 *const bnd = a.bnd_arena + (size_t)slot * (size_t)a.bnd_rows;

    for (;;) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(a.counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= (unsigned)a.n_pairs) break;
        const int p = a.order[idx];
        const PairDesc pd = a.pairs[p];
        const int n = pd.n, m = pd.m;
        const uint8_t *T = a.sym + pd.t_off;
        const uint8_t *O = a.sym + pd.o_off;
        int cap[3];
        // corner scores when no cell is filled (textSeqCompare.py:53-60)
        cap[0] = kp.bg * (n > 0 ? n : m);
        cap[1] = (n > 0) ? kNeg : kp.bg * m;
        cap[2] = (n > 0) ? kp.bg * n : kNeg;
        if (n == 0 && m == 0) { cap[0] = 0; cap[1] = 0; cap[2] = kNeg; }

        if (n > 0 && m > 0) {
            const int nfull = m / kPassW, r = m % kPassW;
            const int npass = nfull + (r ? 1 : 0);
            const long long pass_bytes = ((long long)n + 32) * kPassW;
            for (int ps = 0; ps < npass; ++ps) {
                const int C = (ps < nfull) ? kMaxC : remainder_c(r);
                const int j0 = ps * kPassW;
                const bool last = (ps == npass - 1);
                const int cc = m - 1 - j0;
                const int fin_lane = last ? cc / C : -1;
                const int fin_k = last ? cc % C : -1;
                dispatch_pass<SUBST>(C, kp, T, O, n, m, j0, ps == 0, !last, bnd,
                                     ptr + (size_t)ps * (size_t)pass_bytes, fin_lane, fin_k, cap);
                __syncwarp();
            }
            // the lane that owns column m holds the corner scores
            const int src = (m - 1 - (npass - 1) * kPassW) / (r ? remainder_c(r) : kMaxC);
            cap[0] = __shfl_sync(kFull, cap[0], src);
            cap[1] = __shfl_sync(kFull, cap[1], src);
            cap[2] = __shfl_sync(kFull, cap[2], src);
        }
        __syncwarp();
        uint8_t *ops = a.ops + pd.ops_off;
        int L = 0;
        if (lane == 0) {
            L = traceback_lane(ptr, n, m, ops + (size_t)n + (size_t)m);
            a.ops_len[p] = L;
            if (a.scores) {
                a.scores[3 * (size_t)p + 0] = score_out(cap[0]);
                a.scores[3 * (size_t)p + 1] = score_out(cap[1]);
                a.scores[3 * (size_t)p + 2] = score_out(cap[2]);
            }
        }
        L = __shfl_sync(kFull, L, 0);
        // move the op string from the end of the buffer to its start (left to right order)
        const int shift = n + m - L;
        if (shift > 0) {
            for (int base = 0; base < L; base += 32) {
                const int q = base + lane;
                uint8_t v = 0;
                if (q < L) v = __ldcg(ops + shift + q);
                __syncwarp();
                if (q < L) ops[q] = v;
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

// ---- int32 issue-rate micro-benchmark (roofline denominator) -----------------
// 16 independent chains per thread, operands loaded from memory so ptxas cannot fold them.
//   WHICH 0: two add.s32 per chain-iteration, which ptxas merges into ONE 3-input IADD3
//   WHICH 1: max.s32 + min.s32  -> two VIMNMX
//   WHICH 2: fused add+max, add+min -> two VIADDMNMX
// The host counts INSTRUCTIONS (1, 2, 2 per chain-iteration); see tools/int32_pipes.cu for
// the full table of pipes.
template <int WHICH>
__global__ void __launch_bounds__(256) int32_peak_kernel(int iters, const int *__restrict__ src, int *sink)
{
    int x[16], a[16], b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        x[j] = src[threadIdx.x + j * 7];
        a[j] = src[threadIdx.x + j * 5 + 1];
        b[j] = src[threadIdx.x + j * 3 + 2];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (WHICH == 0) {
                asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(a[j]));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(b[j]));
            } else if (WHICH == 1) {
                asm volatile("max.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(a[j]));
                asm volatile("min.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(b[j]));
            } else {
                x[j] = __viaddmax_s32(x[j], a[j], b[j]);
                x[j] = __viaddmin_s32(x[j], b[j], a[j]);
            }
        }
    }
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= x[j] ^ a[j] ^ b[j];
    if (acc == 0x7fffffff) sink[0] = acc;
}

}  // namespace tanw
