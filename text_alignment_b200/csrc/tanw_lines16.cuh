// tanw_lines16.cuh -- short pairs, two per register: the .u16x2 line kernel (BASELINE config 3).
//
// The line kernel (tanw_kernels.cuh, align_lines_kernel) is bound by the alu pipe like the page
// kernel.  Scores of a 40-120 character line are small: with default_sys |score| < 2 500, so a
// value and its 2-bit origin tag fit 16 bits, and sm_100a has full-rate packed forms of exactly
// the operations the recurrences use (VIADDMNMX.U16x2, VIMNMX3.U16x2; profiles/r1_int32_pipes.txt).
// So here every 32-bit register carries TWO PAIRS: pair A of an 8-lane group in the low half, pair B
// in the high half, both walking the same row of their matrices in the same step; a warp aligns
// eight pairs at once.  Per cell the alu pipe issues about half the instructions of the int32 kernel.
//
// Encoding of a half: 4 * (value + kBias16) + tag, an UNSIGNED 16-bit number (tags as in the
// int32 kernels: M = 2, X = 1, Y = 0, so a plain unsigned max of tagged candidates is still
// "maximum, first in the reference's list order on ties", textSeqCompare.py:72, :80, :88).
// Because every half is a non-negative number far from 0 and 65535 (the host routes a pair here
// only if (2n + 132) * max|param| <= kRange16), adding a packed constant ((c << 16) + c, computed
// in 32-bit arithmetic) with an ordinary 32-bit add never carries between the halves: those adds
// go to the fma pipe / either pipe as in the int32 kernels.  The packed min/max instructions
// take constants in "simd" form ((c & 0xffff) * 0x10001).
//
// Only the D-only recurrences (gap opens <= 0: variants 1 and 2 of tanw_kernels.cuh, where the
// -1e100 sentinel never enters a computed cell) with an equality scorer and match >= mismatch run
// here; everything else keeps the int32 line kernel.  Results are bit-identical by construction
// and by test (tests/test_gpu_parity.py::test_line16_*).
#pragma once
#include "tanw_kernels.cuh"

// tuning knobs of the traceback (tools/kernel_bench.py compares builds)
#ifndef TANW_L16_ROWS
#define TANW_L16_ROWS 8
#endif
#ifndef TANW_L16_STREAM
#define TANW_L16_STREAM 2
#endif
#ifndef TANW_L16_LDPOL
#define TANW_L16_LDPOL 1
#endif
#ifndef TANW_L16_WARPS
#define TANW_L16_WARPS 4
#endif
#ifndef TANW_L16_PREFETCH
#define TANW_L16_PREFETCH 1
#endif

namespace tanw {

// warps per block of the 16-bit line kernel.  Blocks of 1 / 2 warps (a slot frees as soon as ONE
// warp is done, which should suit the one-octet-per-warp chunks of a pipelined call) measured
// slower: 0.812 / 0.786 ms per config-3 launch against 0.783, 690 / 710 GCUPS end to end against 740.
constexpr int kL16Warps = TANW_L16_WARPS;
constexpr int kBias16 = 8192;                 // value + kBias16 in [0, 16383]
constexpr int kRange16 = 8000;                // |value| bound the host guarantees for every half
constexpr unsigned kClean16 = 0xFFFCFFFCu;
constexpr unsigned kTagM16 = 0x00020002u, kTagX16 = 0x00010001u;

struct K16 {
    unsigned mi_add;     // packed for 32-bit adds: 4 * mismatch + tag M
    unsigned dmul;       // 4 * (match - mismatch) >= 0
    unsigned oy_simd;    // simd form of 4 * (gap_open_y + gap_extend_y)
    unsigned ey_add;     // packed for 32-bit adds: 4 * gap_extend_y
    int ox, ex, bg;      // 4 * (gap_open_x + gap_extend_x), 4 * gap_extend_x, 4 * boundary gap
    unsigned tagx;       // kTagX16 in a register the compiler cannot see through: LOP3 takes one immediate only
};

__host__ __device__ inline unsigned pk_add(int c) { return (unsigned)c * 0x10001u; }          // (c << 16) + c
__host__ __device__ inline unsigned pk_simd(int c) { return ((unsigned)c & 0xFFFFu) * 0x10001u; }

__device__ __forceinline__ K16 make_k16(const KParams &kp)
{
    K16 k;
    asm volatile("mov.b32 %0, 0x00010001;" : "=r"(k.tagx));
    const int match = kp.maT >> kShift, mismatch = kp.miT >> kShift;       // the tag bits fall off
    k.mi_add = pk_add(4 * mismatch + kTagM);
    k.dmul = (unsigned)(4 * (match - mismatch));
    k.oy_simd = pk_simd(4 * (kp.oy >> kShift));
    k.ey_add = pk_add(4 * (kp.ey >> kShift));
    k.ox = 4 * (kp.ox >> kShift);
    k.ex = 4 * (kp.ex >> kShift);
    k.bg = 4 * (kp.bg >> kShift);
    return k;
}

template <int C>
struct Strip16 {
    unsigned Xh[C];       // X^ = (X | tag X) - ex * i of the row above, pairs A | B
    unsigned D[C];        // max(M, X, Y) tagged, of the row above
    unsigned oc[C];       // OCR symbols of the strip, A | B << 16
};

__device__ __forceinline__ unsigned bitsel(unsigned mask, unsigned a, unsigned b)      // (a & mask) | (b & ~mask)
{
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(mask), "r"(a), "r"(b));
    return r;
}

// One row of one strip for both pairs: 2 * C cells.  Same quantities as strip_row (variants 1 / 2).
template <int C, bool FINAL, bool EYZ>
__device__ __forceinline__ void strip_row16(Strip16<C> &s, const K16 &kp, unsigned tch2, unsigned xe_add,
                                            unsigned cx_simd, unsigned q_in, unsigned y_in, unsigned dul_in,
                                            unsigned &q_out, unsigned &y_out,
                                            unsigned (&pwA)[C / 4], unsigned (&pwB)[C / 4],
                                            int kfinA, int kfinB, unsigned (&cap)[3])
{
    unsigned q = q_in;
    unsigned ypl = EYZ ? y_in : y_in + kp.ey_add;
    unsigned dul = dul_in;
    unsigned r[4];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        // score: 1 per half where the symbols are equal (~(o ^ t) = -1 - (o ^ t) as a signed half)   (:31-32)
        unsigned e;
        asm("lop3.b32 %0, %1, %2, 0, 0xC3;" : "=r"(e) : "r"(s.oc[k]), "r"(tch2));                     // ~(a ^ b)
        const unsigned eq = __viaddmax_s16x2_relu(e, 0x00020002u, e);      // max(e + 2, e, 0): no zero operand to materialise
        // M[i][j] = max(M,X,Y)[i-1][j-1] + score, tagged M                                          (:70-72)
        const unsigned dc = dul & kClean16;
        const unsigned m2 = eq * kp.dmul + (dc + kp.mi_add);
        // X[i][j] = max(D[i-1][j] + ox, X[i-1][j] + ex)                                             (:83-88)
        const unsigned xraw = __viaddmax_u16x2(s.D[k], cx_simd, s.Xh[k]);
        unsigned xh;
        asm("lop3.b32 %0, %1, 0xFFFCFFFC, %2, 0xEA;" : "=r"(xh) : "r"(xraw), "r"(kp.tagx));       // (xraw & ~3) | tag X
        // Y[i][j] = max(D[i][j-1] + oy, Y[i][j-1] + ey)                                             (:75-80)
        const unsigned yraw = __viaddmax_u16x2(q, kp.oy_simd, ypl);
        const unsigned yc = yraw & kClean16;
        const unsigned xx = xh + xe_add;
        const unsigned dn = __vimax3_u16x2(m2, xx, yc);
        // pointer bits of both cells: tag of D[i-1][j-1] | tag of xraw << 2 | tag of yraw << 4 in the low
        // byte of each half (the multiplies spill a half's top bits into its neighbour's bits 0-3,
        // which the selects below never take)
        unsigned x4, y16;
        asm("mad.lo.u32 %0, %1, 4, 0;" : "=r"(x4) : "r"(xraw));
        asm("mad.lo.u32 %0, %1, 16, 0;" : "=r"(y16) : "r"(yraw));
        r[k & 3] = bitsel(0x00300030u, y16, bitsel(0x000C000Cu, x4, dul));
        if (FINAL) {
            // the corner scores of pair A / pair B, kept in their own half of cap[]
            if (k == kfinA) { cap[0] = bitsel(0xFFFFu, m2, cap[0]); cap[1] = bitsel(0xFFFFu, xx, cap[1]); cap[2] = bitsel(0xFFFFu, yc, cap[2]); }
            if (k == kfinB) { cap[0] = bitsel(0xFFFF0000u, m2, cap[0]); cap[1] = bitsel(0xFFFF0000u, xx, cap[1]); cap[2] = bitsel(0xFFFF0000u, yc, cap[2]); }
        }
        dul = s.D[k];
        s.Xh[k] = xh; s.D[k] = dn;
        q = dn; ypl = EYZ ? yc : yc + kp.ey_add;
        if ((k & 3) == 3) {
            // byte 0 of a register is pair A's cell, byte 2 pair B's
            const unsigned p01 = __byte_perm(r[0], r[1], 0x6240);      // A0 A1 B0 B1
            const unsigned p23 = __byte_perm(r[2], r[3], 0x6240);      // A2 A3 B2 B3
            pwA[k >> 2] = __byte_perm(p01, p23, 0x5410);
            pwB[k >> 2] = __byte_perm(p01, p23, 0x7632);
        }
    }
    q_out = q;
    y_out = EYZ ? ypl : ypl - kp.ey_add;
}

struct Line16State {
    unsigned q_out, y_out, q_prev, tnext, bq;
    int xe;                          // ex * i of the row this lane computes next (scaled by 4)
    const uint8_t *tpA, *tpB;
    uint8_t *pstA, *pstB;
};

// Pointer words of a row: TANW_L16_STREAM 0 = st.cg, 1 = st.cs (evict first), 2 = L2 evict-last hint.
template <int C>
__device__ __forceinline__ void l16_store_ptr(uint8_t *dst, const unsigned (&pw)[C / 4])
{
#if TANW_L16_STREAM == 2
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (C % 16 == 0) {
#pragma unroll
        for (int v = 0; v < C / 16; ++v)
            asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;"
                         :: "l"(dst + 16 * v), "r"(pw[4 * v]), "r"(pw[4 * v + 1]), "r"(pw[4 * v + 2]), "r"(pw[4 * v + 3]), "l"(pol) : "memory");
    } else if (C % 8 == 0) {
#pragma unroll
        for (int v = 0; v < C / 8; ++v)
            asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;"
                         :: "l"(dst + 8 * v), "r"(pw[2 * v]), "r"(pw[2 * v + 1]), "l"(pol) : "memory");
    } else {
#pragma unroll
        for (int v = 0; v < C / 4; ++v)
            asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(dst + 4 * v), "r"(pw[v]), "l"(pol) : "memory");
    }
#else
    store_ptr_words<C, TANW_L16_STREAM != 0>(dst, pw);
#endif
}

// Pointer word read back by the traceback: TANW_L16_LDPOL 0 = ld.cg, 1 = L2 evict-first hint (the
// line is dead once the walk has passed it), 2 = ld.lu.
__device__ __forceinline__ unsigned l16_ld(const unsigned *p)
{
#if TANW_L16_LDPOL == 1
    unsigned long long pol;
    unsigned v;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
#elif TANW_L16_LDPOL == 2
    return __ldlu(p);
#else
    return __ldcg(p);
#endif
}

template <int C, bool GUARDED, bool EYZ>
__device__ __forceinline__ void line_step16(Strip16<C> &s, Line16State &ls, const K16 &kp, int nA, int nB,
                                            bool actA, bool actB, int t, int gl,
                                            int finA_lane, int finA_k, int finB_lane, int finB_k, unsigned (&cap)[3])
{
    const int i = t - gl;
    unsigned q_in = __shfl_up_sync(kFull, ls.q_out, 1, kLineG);
    unsigned y_in = __shfl_up_sync(kFull, ls.y_out, 1, kLineG);
    if (gl == 0) { q_in = ls.bq | kTagM16; y_in = ls.bq; }          // column 0: M = Y = bg * i (:54-56)
    const unsigned dul_in = ls.q_prev;
    const unsigned tch2 = ls.tnext;
    if (!GUARDED) {
        ls.tnext = (unsigned)__ldg(ls.tpA) | ((unsigned)__ldg(ls.tpB) << 16);
    } else {
        // 0x200: matches no OCR symbol (codes < 0x100, padding 0x100)
        const unsigned a = (actA && i >= 0 && i < nA) ? (unsigned)__ldg(ls.tpA) : 0x200u;
        const unsigned b = (actB && i >= 0 && i < nB) ? (unsigned)__ldg(ls.tpB) : 0x200u;
        ls.tnext = a | (b << 16);
    }
    const bool rowA = actA && i >= 1 && i <= nA, rowB = actB && i >= 1 && i <= nB;
    if (!GUARDED || rowA || rowB) {
        unsigned pwA[C / 4], pwB[C / 4];
        const int kfinA = (GUARDED && rowA && i == nA && gl == finA_lane) ? finA_k : -1;
        const int kfinB = (GUARDED && rowB && i == nB && gl == finB_lane) ? finB_k : -1;
        strip_row16<C, GUARDED, EYZ>(s, kp, tch2, pk_add(ls.xe), pk_simd(kp.ox - ls.xe), q_in, y_in, dul_in,
                                     ls.q_out, ls.y_out, pwA, pwB, kfinA, kfinB, cap);
        if (!GUARDED || rowA) l16_store_ptr<C>(ls.pstA, pwA);
        if (!GUARDED || rowB) l16_store_ptr<C>(ls.pstB, pwB);
    }
    ls.q_prev = q_in;
    ls.xe += kp.ex;
    ls.bq += pk_add(kp.bg);
    ls.tpA += 1;
    ls.tpB += 1;
    ls.pstA += kLineG * C;
    ls.pstB += kLineG * C;
}

// Traceback of the eight pairs of an octet (textSeqCompare.py:96-164): lane 0 of a group walks pair
// A, lane 1 pair B, at the same time.  A tile is kL16Rows rows x 2 strips of decoded, guarded pointer
// bytes (see traceback_groups); the group's lanes load BOTH pairs' tiles before any of the words is
// looked at, so a round costs one memory round trip.
//
// Measured alternatives (tools/kernel_bench.py, config 3, 125 000 pairs; this form: 0.85 ms):
//   * 16-row tiles (half the rounds, twice the words per round): 0.93 ms;
//   * fill kernel + a thread-per-pair traceback kernel over all pairs of the launch: the fill alone
//     takes 0.61 ms, but 1.4 * 10^7 dependent single-sector reads of pointer bytes that have left
//     the L2 by then (1.25 GB per launch) take 0.36 ms -- DRAM-bound on random 32-byte reads;
//   * the same with a strip-major pointer layout (consecutive rows of a strip share sectors; the
//     walk then needs 0.28 ms) makes the fill's stores uncoalesced: fill 1.56 ms.
// The tile loads miss the L2 although a warp reads what it wrote tens of microseconds before: a
// config-3 octet holds 72 KB of pointer bytes and 2 368 resident warps keep ~100 MB of them alive,
// more than the L2 retains beside the rest of the traffic (ncu: 1.13 GB written, 0.93 GB read
// back from DRAM per launch whatever the store flavour).  Marking the stores evict-last and the
// read-backs evict-first (a line is dead once the walk has passed it) gains 2 % (0.794 -> 0.777 ms).
constexpr int kL16Rows = TANW_L16_ROWS;                // rows of a traceback tile: 8 or 16
constexpr int kL16PerLane = kL16Rows / kLineG;         // tile rows a lane loads
constexpr int kL16TileWords = (kL16Rows + 1) * kLineTile;

template <int C>
__device__ __forceinline__ void l16_load(const uint8_t *ptr, bool mine, int hx, int sidx, int gl,
                                         unsigned (&w)[kL16PerLane][2 * C / 4])
{
#pragma unroll
    for (int rr = 0; rr < kL16PerLane; ++rr) {
#pragma unroll
        for (int q = 0; q < 2 * C / 4; ++q) w[rr][q] = 0xFFFFFFFFu;     // outside the matrix
        const int row = hx - gl - kLineG * rr;
        if (mine && row >= 1) {
            const unsigned *hi = reinterpret_cast<const unsigned *>(ptr + ((size_t)(row + sidx) * kLineG + sidx) * C);
#pragma unroll
            for (int q = 0; q < C / 4; ++q) w[rr][C / 4 + q] = l16_ld(hi + q);
            if (sidx >= 1) {
                const unsigned *lo = reinterpret_cast<const unsigned *>(ptr + ((size_t)(row + sidx - 1) * kLineG + sidx - 1) * C);
#pragma unroll
                for (int q = 0; q < C / 4; ++q) w[rr][q] = l16_ld(lo + q);
            }
        }
    }
}

template <int C>
__device__ __forceinline__ void l16_store(unsigned *tile, int gl, const unsigned (&w)[kL16PerLane][2 * C / 4])
{
#pragma unroll
    for (int rr = 0; rr < kL16PerLane; ++rr) {
        unsigned *row = tile + (gl + kLineG * rr) * kLineTile;
        row[0] = kGuardWord;
#pragma unroll
        for (int q = 0; q < 2 * C / 4; ++q)
            row[1 + q] = (w[rr][q] == 0xFFFFFFFFu) ? kGuardWord : (0x2A2A2A2Au - (w[rr][q] & 0x3F3F3F3Fu)) << 1;
    }
}

template <int C>
__device__ __forceinline__ int traceback_groups2(const uint8_t *ptrA, const uint8_t *ptrB, int n, int m, bool act,
                                                 uint8_t *ops_end, unsigned *tiles, int gl)
{
    int x = n, y = m, k = 0, st = -1;
    unsigned *const tileA = tiles, *const tileB = tiles + kL16TileWords;
    for (int w = gl; w < kLineTile; w += kLineG) {                  // guard rows above the tiles
        tileA[kL16Rows * kLineTile + w] = kGuardWord;
        tileB[kL16Rows * kLineTile + w] = kGuardWord;
    }
    while (__any_sync(kFull, gl < 2 && act && x > 0 && y > 0)) {      // lanes 0 / 1 of a group hold the pairs' state
        const int xA = __shfl_sync(kFull, x, 0, kLineG), yA = __shfl_sync(kFull, y, 0, kLineG);
        const int xB = __shfl_sync(kFull, x, 1, kLineG), yB = __shfl_sync(kFull, y, 1, kLineG);
        const unsigned acts = __ballot_sync(kFull, act && x > 0 && y > 0) >> ((threadIdx.x & 31) & ~(kLineG - 1));
        const bool mineA = (acts & 1u) != 0, mineB = (acts & 2u) != 0;
        unsigned wa[kL16PerLane][2 * C / 4], wb[kL16PerLane][2 * C / 4];
        l16_load<C>(ptrA, mineA, xA, mineA ? (yA - 1) / C : 0, gl, wa);
        l16_load<C>(ptrB, mineB, xB, mineB ? (yB - 1) / C : 0, gl, wb);
        __syncwarp();
        l16_store<C>(tileA, gl, wa);
        l16_store<C>(tileB, gl, wb);
        __syncwarp();
        if (gl < 2 && act && x > 0 && y > 0) {
            const unsigned char *tb = reinterpret_cast<const unsigned char *>(gl ? tileB : tileA);
            const int sidx = (y - 1) / C;
            const int col0 = (sidx - 1) * C;                 // 0-based column of the first data byte
            int off = 4 + (y - 1) - col0;                    // row 0 of the window
            unsigned d = tb[off];
            int s = 2 * st;
            if (st < 0) s = (int)(d & 6u);                   // state from mat_ptr first          (:102)
            uint8_t *op = ops_end - k;
            const uint8_t *const op0 = op;
            while (!(d & 0x80u)) {                                                        // :115-145
                *--op = (uint8_t)(s >> 1);
                int delta;                                   // byte s of {35, 0, 36, 0, -1, -1}, byte s+1 its sign extension
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(delta)
                    : "r"(0x00240023), "r"(0x0000FFFF), "r"(s * 0x1111 + 0x1110));
                off += delta;
                s = (int)((d >> s) & 6u);
                d = tb[off];
            }
            k += (int)(op0 - op);
            const int r = off / (kLineTile * 4);
            x -= r;
            y = col0 + (off - r * (kLineTile * 4)) - 3;
            st = s >> 1;
        }
        __syncwarp();
    }
    if (gl < 2) {
        while (y > 0) { ++k; *(ops_end - k) = 2; --y; }      // :154-158
        while (x > 0) { ++k; *(ops_end - k) = 1; --x; }      // :160-164
    }
    __syncwarp();                       // the walkers' op bytes are read by all lanes of the group next
    return k;                           // valid in lanes 0 (pair A) and 1 (pair B) of the group
}

__device__ __forceinline__ int half_score(unsigned v, int hi)
{
    return (int)(((hi ? (v >> 16) : (v & 0xFFFFu)) >> 2)) - kBias16;
}

// The work item of a warp: octet o of a class = its entries 8o .. 8o+7; group g aligns entries
// 8o+2g (pair A) and 8o+2g+1 (B).  The descriptors of the NEXT octet are fetched while the current
// one is in its traceback, so that the chain work counter -> sorted list -> pair table -> symbols is
// not paid in the open between two octets.
struct OctetDesc {
    int cls;                 // strip-width class (C = 4 * (cls + 1))
    int pA, pB;              // pair indices or -1
    PairDesc dA, dB;
};

__device__ __forceinline__ unsigned next_work(unsigned *counter, int lane)
{
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(counter, 1u);
    return __shfl_sync(kFull, idx, 0);
}

__device__ __forceinline__ OctetDesc load_octet(const LineArgs &a, const LineClasses &lc, unsigned idx, int g)
{
    OctetDesc d;
    d.cls = 3; d.pA = -1; d.pB = -1;
    d.dA.t_off = 0; d.dA.o_off = 0; d.dA.ops_off = 0; d.dA.n = 0; d.dA.m = 0;
    d.dB = d.dA;
    if (idx >= (unsigned)a.n_quads) return d;
    if ((int)idx >= lc.quad0[2]) d.cls = 2;
    if ((int)idx >= lc.quad0[1]) d.cls = 1;
    if ((int)idx >= lc.quad0[0]) d.cls = 0;
    const int e = 8 * ((int)idx - lc.quad0[d.cls]) + 2 * g;
    if (e < lc.count[d.cls]) d.pA = a.sorted[lc.start[d.cls] + e];
    if (e + 1 < lc.count[d.cls]) d.pB = a.sorted[lc.start[d.cls] + e + 1];
    if (d.pA >= 0) d.dA = a.pairs[d.pA];
    if (d.pB >= 0) d.dB = a.pairs[d.pB];
    return d;
}

template <int C, bool EYZ>
__device__ __forceinline__ void line_octet(const LineArgs &a, const KParams &kp32, const K16 &kp, const LineClasses &lc,
                                           const OctetDesc &cur, uint8_t *ptrA, uint8_t *ptrB, unsigned *tiles, int lane,
                                           unsigned &next_idx, OctetDesc &next)
{
    const int gl = lane & (kLineG - 1);
    const int pA = cur.pA, pB = cur.pB;
    const int nA = cur.dA.n, mA = cur.dA.m, nB = cur.dB.n, mB = cur.dB.m;
    const long long offA = cur.dA.ops_off, offB = cur.dB.ops_off;
    const uint8_t *TA = a.sym + cur.dA.t_off, *OA = a.sym + cur.dA.o_off;
    const uint8_t *TB = a.sym + cur.dB.t_off, *OB = a.sym + cur.dB.o_off;
    const bool actA = nA > 0 && mA > 0, actB = nB > 0 && mB > 0;
    TANW_ASSERT(a.check, (!actA || (line_ptr_bytes(nA, mA) <= a.slot_bytes && mA <= kLineG * C)) &&
                (!actB || (line_ptr_bytes(nB, mB) <= a.slot_bytes && mB <= kLineG * C)), 6);
    // every half stays inside [0, 65535]: (2n + m + 4) * max|param| <= kRange16 (the host's routing rule)
    TANW_ASSERT(a.check, (2 * max(nA, nB) + kLineMaxM + 4) * max(max(abs(kp.ox), abs(kp.ex)), max(abs(kp.bg), (int)kp.dmul)) / 4 <= 2 * kRange16, 7);
    // tallest / shortest active pair of the octet
    int nmax = max(actA ? nA : 0, actB ? nB : 0), nmin = min(actA ? nA : 0x7fffffff, actB ? nB : 0x7fffffff);
    const bool all_act = __all_sync(kFull, actA && actB);
#pragma unroll
    for (int d = kLineG; d < 32; d <<= 1) {
        nmax = max(nmax, __shfl_xor_sync(kFull, nmax, d));
        nmin = min(nmin, __shfl_xor_sync(kFull, nmin, d));
    }
    if (!all_act) nmin = 0;                                  // an idle half: every step is guarded
    unsigned cap[3] = {0u, 0u, 0u};
    if (nmax > 0) {
        const int c0 = gl * C;
        Strip16<C> s;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int c = c0 + k;
            const unsigned oa = (actA && c < mA) ? (unsigned)__ldg(OA + c) : 0x100u;
            const unsigned ob = (actB && c < mB) ? (unsigned)__ldg(OB + c) : 0x100u;
            s.oc[k] = oa | (ob << 16);
            const unsigned base = pk_add(kp.bg * (c + 1) + 4 * kBias16);     // row 0 (:57-60)
            s.Xh[k] = base | kTagX16;
            s.D[k] = base | kTagM16;
        }
        Line16State ls;
        ls.q_out = pk_add(kp.bg * (c0 + C) + 4 * kBias16) | kTagM16;
        ls.y_out = 0u;                                       // Y[0][j] = -inf: never read by a computed cell
        ls.q_prev = pk_add(kp.bg * c0 + 4 * kBias16) | kTagM16;
        ls.tnext = (gl == 0) ? ((actA ? (unsigned)__ldg(TA) : 0x200u) | ((actB ? (unsigned)__ldg(TB) : 0x200u) << 16))
                             : 0x02000200u;
        ls.tpA = TA + (1 - gl);
        ls.tpB = TB + (1 - gl);
        ls.xe = kp.ex * (1 - gl);
        ls.bq = pk_add(kp.bg * (1 - gl) + 4 * kBias16);
        ls.pstA = ptrA + ((size_t)kLineG + gl) * C;          // step t = 1
        ls.pstB = ptrB + ((size_t)kLineG + gl) * C;
        const int finA_lane = actA ? (mA - 1) / C : -1, finA_k = actA ? (mA - 1) % C : -1;
        const int finB_lane = actB ? (mB - 1) / C : -1, finB_k = actB ? (mB - 1) % C : -1;
        const int last_step = nmax + kLineG - 1;
        int t = 1;
        for (; t <= min(kLineG - 1, last_step); ++t)         // ramp-up
            line_step16<C, true, EYZ>(s, ls, kp, nA, nB, actA, actB, t, gl, finA_lane, finA_k, finB_lane, finB_k, cap);
        // every lane of every group on a row in [1, n-1] of both its pairs; two steps per iteration
#pragma unroll 1
        for (; t + 1 <= nmin - 1; t += 2) {
            line_step16<C, false, EYZ>(s, ls, kp, nA, nB, actA, actB, t, gl, finA_lane, finA_k, finB_lane, finB_k, cap);
            line_step16<C, false, EYZ>(s, ls, kp, nA, nB, actA, actB, t + 1, gl, finA_lane, finA_k, finB_lane, finB_k, cap);
        }
#pragma unroll 1
        for (; t <= nmin - 1; ++t)
            line_step16<C, false, EYZ>(s, ls, kp, nA, nB, actA, actB, t, gl, finA_lane, finA_k, finB_lane, finB_k, cap);
        for (; t <= last_step; ++t)                          // ramp-down and the taller pairs' tails
            line_step16<C, true, EYZ>(s, ls, kp, nA, nB, actA, actB, t, gl, finA_lane, finA_k, finB_lane, finB_k, cap);
    }
    __syncwarp();
#if TANW_L16_PREFETCH
    // the next octet's descriptors travel while this one is traced back
    next_idx = next_work(a.counter, lane);
    next = load_octet(a, lc, next_idx, lane >> 3);
#endif
    // from here on a lane works for pair (gl & 1) of its group
    const int h = gl & 1;
    const int p = h ? pB : pA, n = h ? nB : nA, m = h ? mB : mA;
    const bool act = h ? actB : actA;
    uint8_t *ops = a.ops + (h ? offB : offA);
    const int kmine = traceback_groups2<C>(ptrA, ptrB, n, m, act, ops + (size_t)n + (size_t)m, tiles, gl);
    const int L = __shfl_sync(kFull, kmine, h, kLineG);
    {   // the lane that owns column m of pair h holds its corner scores in half h
        const int fin_lane = act ? (m - 1) / C : 0;
        const int src = (lane & ~(kLineG - 1)) + fin_lane;
        const unsigned v0 = __shfl_sync(kFull, cap[0], src), v1 = __shfl_sync(kFull, cap[1], src), v2 = __shfl_sync(kFull, cap[2], src);
        if (p >= 0 && gl < 2) {
            a.ops_len[p] = L;
            if (a.scores) {
                int s0, s1, s2;
                if (act) {
                    s0 = half_score(v0, h); s1 = half_score(v1, h); s2 = half_score(v2, h);
                } else {                                     // no cell: the boundary values (:53-60)
                    const int bg = kp32.bg >> kShift;
                    s0 = bg * (n > 0 ? n : m);
                    s1 = (n > 0) ? kNeg : bg * m;
                    s2 = (n > 0) ? bg * n : kNeg;
                    if (n == 0 && m == 0) { s0 = 0; s1 = 0; s2 = kNeg; }
                }
                a.scores[3 * (size_t)p + 0] = s0;
                a.scores[3 * (size_t)p + 1] = s1;
                a.scores[3 * (size_t)p + 2] = s2;
            }
        }
    }
#if !TANW_L16_PREFETCH
    next_idx = next_work(a.counter, lane);
    next = load_octet(a, lc, next_idx, lane >> 3);
#endif
    // move each op string to the start of its buffer: four lanes per pair
    const int shift = (p >= 0) ? n + m - L : 0;
    const int sub = gl >> 1;                                 // 0..3 within the pair's lanes
    int rounds = (shift > 0) ? (L + 3) / 4 : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) rounds = max(rounds, __shfl_xor_sync(kFull, rounds, d));
    for (int it = 0; it < rounds; ++it) {
        const int q = it * 4 + sub;
        uint8_t v = 0;
        const bool on = shift > 0 && q < L;
        if (on) v = __ldcg(ops + shift + q);
        __syncwarp();
        if (on) ops[q] = v;
        __syncwarp();
    }
}

template <int VAR>
__global__ void __launch_bounds__(kL16Warps * 32, TANW_MINB * kWarpsPerBlock / kL16Warps)
align_lines16_kernel(const LineArgs a, const __grid_constant__ KParams kp32)
{
    __shared__ unsigned tiles[kL16Warps][4 * 2 * kL16TileWords];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane >> 3;
    const long long slot = (((long long)blockIdx.x * kL16Warps + warp) * 4 + g) * 2;
    uint8_t *const ptrA = a.ptr_arena + (size_t)slot * (size_t)a.slot_bytes;
    uint8_t *const ptrB = ptrA + (size_t)a.slot_bytes;
    unsigned *const tile = tiles[warp] + g * (2 * kL16TileWords);
    const LineClasses lc = *a.classes;
    const K16 kp = make_k16(kp32);
    unsigned idx = next_work(a.counter, lane);
    OctetDesc cur = load_octet(a, lc, idx, g);
    while (idx < (unsigned)a.n_quads) {
        unsigned nidx = 0;
        OctetDesc nxt;
        switch (cur.cls) {
        case 0:  line_octet<4,  VAR == 2>(a, kp32, kp, lc, cur, ptrA, ptrB, tile, lane, nidx, nxt); break;
        case 1:  line_octet<8,  VAR == 2>(a, kp32, kp, lc, cur, ptrA, ptrB, tile, lane, nidx, nxt); break;
        case 2:  line_octet<12, VAR == 2>(a, kp32, kp, lc, cur, ptrA, ptrB, tile, lane, nidx, nxt); break;
        default: line_octet<16, VAR == 2>(a, kp32, kp, lc, cur, ptrA, ptrB, tile, lane, nidx, nxt); break;
        }
        __syncwarp();
        idx = nidx;
        cur = nxt;
    }
}

}  // namespace tanw
