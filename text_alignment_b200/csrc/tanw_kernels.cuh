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
//   * all H/E/F state (here M/X/Y) lives in registers: per column X and D = max(M,X,Y) of the
//     row above (plus W = max(M,Y) in the general variant), per step the running D (or
//     Q = max(M,X)) and Y of the column to the left;
//   * scores are carried in int32 fixed point, value*64, and the low two bits hold the origin
//     tag of a value (M = 2, X = 1, Y = 0).  A plain integer max over tagged candidates
//     therefore returns the maximum AND, on equal values, the candidate that comes first in
//     the reference's list order (M, X, Y) -- list.index(max(list)),
//     textSeqCompare.py:72,:80,:88 -- without any compare/select for the argmax;
//   * B200 issues integer min/max/logic on the "alu" pipe and IMAD on the "fma" pipe, 64
//     lanes/clk/SM each (profiles/r1_int32_pipes.txt).  The recurrences need the alu pipe
//     (VIADDMNMX, LOP3), so everything else is phrased as IMAD work: scores are multiples of
//     64 with the tag in bits 0-1, so the low six bits of dul + 4*xraw + 16*yraw ARE the three
//     2-bit traceback pointers of the cell (two IMADs, no masking);
//   * pointers are written one byte per cell, step-major, so every warp store is one
//     contiguous 32*C byte segment;
//   * the traceback runs in the same kernel right after the pair's fill: the warp prefetches
//     a 64x64 tile of pointer bytes whose corner is the current cell into shared memory with
//     32 independent word loads per lane, one lane walks inside the tile (branch-free), repeat;
//   * short pairs (m <= 128) are aligned four per warp, 8 lanes each (align_lines_kernel); one
//     whole-manuscript pair is spread over one warp per column stripe, all resident at once,
//     handing their edges over through flag-stamped records (align_long_kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tanw {

constexpr int      kShift   = 6;                 // fixed point: value << 6 (bits 2-5 stay zero)
constexpr int      kTagM    = 2;                 // "came from M"
constexpr int      kTagX    = 1;                 // "came from X"; "came from Y" is 0
constexpr int      kTagMask = 3;
constexpr int      kNeg     = -(1 << 30);        // stands in for -1e100 (textSeqCompare.py:55,:60)
#ifndef TANW_MAXC
#define TANW_MAXC 16
#endif
constexpr int      kMaxC    = TANW_MAXC;         // widest strip (columns per lane), multiple of 4
constexpr int      kPassW   = 32 * kMaxC;        // columns of a full pass
constexpr int      kWarpsPerBlock = 4;
constexpr unsigned kFull    = 0xFFFFFFFFu;

struct KParams {
    int maT, miT;            // (match<<6)|kTagM, (mismatch<<6)|kTagM
    int ox, ex, oy, ey;      // (gox+gex)<<6, gex<<6, (goy+gey)<<6, gey<<6
    int bg;                  // boundary_gap<<6 (module-level gap_extend, textSeqCompare.py:9)
    int subst_k;
    const int *subst;        // device table, entry = (score<<6)|kTagM, or nullptr
};

// Query profile of a tabulated scorer (SUBST == 2, page kernel): before a pass, every lane
// tabulates, for each of the K transcript symbols, the scores against its own C OCR columns as
// C signed bytes -- 16 bytes per (symbol, lane) in shared memory, laid out [symbol][lane][16] so
// that a row's scores come with ONE conflict-free 128-bit load.  Per cell that leaves a byte
// extraction (PRMT, sign-extending) and one multiply-add  m2 = score * 64 + (dc | tag M): one alu
// instruction like the equality scorer's compare, where the table lookup per cell (SUBST == 1)
// needs address arithmetic and a shared-memory load whose bank conflicts saturate the LSU
// (measured on 2 000 config-2 pages, K = 24: 1.54 x the equality scorer's time).  Needs
// K <= kProfileMaxK and |score| <= 127 (the host checks; else SUBST == 1).
constexpr int kProfileMaxK = 32;
__host__ __device__ inline int profile_bytes_per_warp(int k)
{
    const int tile = ((64 + 1) * 17 * 4 + 255) / 256 * 256;      // the traceback tile shares the space
    const int prof = k * 32 * 16;
    return prof > tile ? prof : tile;
}

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
    const KParams  *kparams;     // MULTI kernels: the scoring systems of the batch ...
    const int      *sidx;        // ... and which of them pair p uses (evaluate_text_alignment.py:181-194)
    int            *check;       // TANW_CHECKED builds: first failed device assertion (0 = none)
};

// Device assertions of the TANW_CHECKED build (tools/stress_gpu.py runs under it): the first failure
// leaves its code in BatchArgs::check / LongArgs::check and the host turns it into an error.  The
// shipping build compiles them out.
#ifdef TANW_CHECKED
#define TANW_ASSERT(word, cond, code) do { if ((word) && !(cond)) atomicCAS((word), 0, (code)); } while (0)
#else
#define TANW_ASSERT(word, cond, code) do { } while (0)
#endif

// Width of the remainder pass: smallest multiple of 4 columns per lane covering r columns.
__host__ __device__ inline int remainder_c(int r) { return ((r + 127) / 128) * 4; }

// Bytes of traceback pointers of one pair (all passes, (n+32) step slots each).
__host__ __device__ inline long long ptr_bytes(int n, int m, int cfull = kMaxC)
{
    if (n <= 0 || m <= 0) return 0;
    long long steps = (long long)n + 32;
    int nfull = m / (32 * cfull), r = m % (32 * cfull);
    return steps * 32 * ((long long)nfull * cfull + (r ? remainder_c(r) : 0));
}

// Bitwise cleaning of the tag bits as opaque asm: the pointer extraction below computes
// (raw - clean) on the IMAD pipe, and the compiler must not fold that back into (raw & 3),
// which would put it on the already saturated alu pipe.
__device__ __forceinline__ int clean_tag(int v)
{
    int r;
    asm("and.b32 %0, %1, 0xFFFFFFFC;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int clean_tag_or(int v, int tag)
{
    int r;
    asm("lop3.b32 %0, %1, 0xFFFFFFFC, %2, 0xEA;" : "=r"(r) : "r"(v), "r"(tag));   // (v & ~3) | tag
    return r;
}

// Symbol codes index the K x K table of a tabulated scorer: a code >= K (which the host reports as
// an error after the batch) must not read outside the table.
template <int SUBST>
__device__ __forceinline__ int table_code(int v, const KParams &kp)
{
    return SUBST ? min(v, kp.subst_k - 1) : v;
}

// Kernel variants (template parameter VAR):
//   0  general: any gap parameters.  Keeps W = max(M,Y), Q = max(M,X) and D = max(M,X,Y).
//   1  gap_open_x <= 0 and gap_open_y <= 0 (every scoring system the reference ships or sweeps):
//      then max(M,Y)+ox vs X+ex and max(M,X)+oy vs Y+ey can both be taken from D = max(M,X,Y):
//      the extra candidate (X+ox resp. Y+oy) never beats X+ex resp. Y+ey and never changes the
//      first-index argmax (case analysis in DESIGN.md 4.1).  W and Q disappear: two maxes and
//      one register per column less, D is one VIMNMX3.
//   2  variant 1 with gap_extend_y == 0 (the reference's default_sys): no Y + ey add.
//   3, 4  variants 1, 2 for the chained-stripe kernel, where a stripe is ONE warp bound by the
//      latency of the dependent chain along a row, not by issue slots: X as in variant 1 (from D),
//      but Y from Q = max(M, X) as the reference writes it (max(M+oy, X+oy, Y+ey), :75-80), so that
//      the chain from a cell to its right neighbour is  Y -> max(Q+oy, Y+ey) -> clean  (two
//      dependent instructions) instead of  Y -> D = max3(M, X, Y) -> max(D+oy, Y+ey) -> clean.
//      Same instruction count (VIADDMNMX + VIMNMX for VIADD + VIMNMX3), one more of them on the
//      alu pipe -- which is why the batched kernels, bound by that pipe, keep variants 1 / 2.
template <int C>
struct Strip {
    int W[C];         // general variant only: max(M|tagM, Y) of the row above
    int Xh[C];        // X^ = (X|tagX) - ex*i of the row above
    int D[C];         // max(M, X, Y) tagged, of the row above
    int oc[C];        // the strip's OCR symbols
};

// One row of one strip: C cells.  See the file header for the value encoding.
//   q_in, y_in : Q = max(M,X) tagged and Y (clean) of the cell left of the strip, same row
//   dul_in     : D of the cell up-left of the strip's first cell
//   xe, cx     : ex*i and ox - ex*i for this lane's row i
// Returns the strip's right edge in q_out / y_out and the C pointer bytes in pw[C/4].
// EYZ: gap_extend_y == 0 (the reference's default_sys), which saves the Y + ey add.
template <int C, bool FINAL, int SUBST, int VAR>
__device__ __forceinline__ void strip_row(Strip<C> &s, const KParams &kp, int tch, int xe, int cx,
                                          int q_in, int y_in, int dul_in,
                                          int &q_out, int &y_out, unsigned (&pw)[C / 4],
                                          int kfin, int (&cap)[3], const uint4 *prof = nullptr)
{
    constexpr bool FAST = (VAR >= 1);                 // X from D: no W
    constexpr bool YQ = (VAR == 0 || VAR >= 3);       // Y from Q = max(M, X); the edge carries (Q, Y)
    constexpr bool EYZ = (VAR == 2 || VAR == 4);
    const int *srow = SUBST == 1 ? kp.subst + tch * kp.subst_k : nullptr;
    unsigned pr[4] = {0u, 0u, 0u, 0u};
    if (SUBST == 2) {                                       // this row's scores against the strip's columns
        const uint4 v = prof[tch * 32];
        pr[0] = v.x; pr[1] = v.y; pr[2] = v.z; pr[3] = v.w;
    }
    int q = q_in;                             // general: Q of the cell to the left; FAST: its D
    int ypl = EYZ ? y_in : y_in + kp.ey;      // Y of the column to the left, + ey
    int dul = dul_in;
    unsigned bytes[4];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        // M[i][j] = max(M,X,Y)[i-1][j-1] + score, tagged as an M value         (:70-72)
        const int dc = SUBST == 2 ? clean_tag_or(dul, kTagM) : clean_tag(dul);
        int m2;
        if (SUBST == 2) {
            int sx;                              // byte k & 3 of the profile word, sign-extended
            const int b = k & 3;                 // selector: byte b, then its sign three times (an immediate once unrolled)
            asm("prmt.b32 %0, %1, 0, %2;" : "=r"(sx) : "r"(pr[k >> 2]), "r"(b | ((b | 8) << 4) | ((b | 8) << 8) | ((b | 8) << 12)));
            asm("mad.lo.s32 %0, %1, 64, %2;" : "=r"(m2) : "r"(sx), "r"(dc));
        } else if (SUBST == 1) {
            m2 = dc + srow[s.oc[k]];            // generic load: the table is in shared memory when K <= kSubstSmemK
        } else {
            // compare + add + predicated add instead of compare + select + add: ptxas turns the
            // two adds into VIADD, which B200 issues on whichever of the alu / fma pipes is free
            // (profiles/r1_int32_pipes.txt), so the saturated alu pipe loses one op per cell
            // (measured +10 % on config 2)                                       (:31-32)
            asm("{ .reg .pred p;\n\t"
                "setp.eq.s32 p, %1, %2;\n\t"
                "add.s32 %0, %3, %4;\n\t"
                "@p add.s32 %0, %3, %5; }"
                : "=r"(m2) : "r"(s.oc[k]), "r"(tch), "r"(dc), "r"(kp.miT), "r"(kp.maT));
        }
        // X[i][j] = max(M[i-1][j]+ox, X[i-1][j]+ex, Y[i-1][j]+ox)               (:83-88)
        const int xraw = __viaddmax_s32(FAST ? s.D[k] : s.W[k], cx, s.Xh[k]);
        const int xh = clean_tag_or(xraw, kTagX);
        // Y[i][j] = max(M[i][j-1]+oy, X[i][j-1]+oy, Y[i][j-1]+ey)               (:75-80)
        const int yraw = __viaddmax_s32(q, kp.oy, ypl);
        const int yc = clean_tag(yraw);
        int dn, qn;
        if (!YQ) {
            dn = __vimax3_s32(m2, xh + xe, yc);              // max(M, X, Y)
            qn = dn;
        } else {
            qn = __viaddmax_s32(xh, xe, m2);                 // max(M, X)
            dn = max(qn, yc);                                // max(M, X, Y)
            if (!FAST) s.W[k] = max(yc, m2);                 // max(M, Y)
        }
        // pointer byte: scores are multiples of 64, so the low six bits of
        // dul + 4*xraw + 16*yraw are exactly tagM | tagX<<2 | tagY<<4 (two IMADs); bits 6-7 are
        // value garbage that the traceback masks off.
        int rb;
        asm("mad.lo.s32 %0, %1, 4, %2;" : "=r"(rb) : "r"(xraw), "r"(dul));
        asm("mad.lo.s32 %0, %1, 16, %0;" : "+r"(rb) : "r"(yraw));
        bytes[k & 3] = (unsigned)rb;
        if (FINAL) {
            if (k == kfin) { cap[0] = m2; cap[1] = xh + xe; cap[2] = yc; }
        }
        dul = s.D[k];
        s.Xh[k] = xh; s.D[k] = dn;
        q = qn; ypl = EYZ ? yc : yc + kp.ey;
        if ((k & 3) == 3) {
            const unsigned lo = __byte_perm(bytes[0], bytes[1], 0x0040);
            const unsigned hi = __byte_perm(bytes[2], bytes[3], 0x0040);
            pw[k >> 2] = __byte_perm(lo, hi, 0x5410);
        }
    }
    q_out = q;
    y_out = EYZ ? ypl : ypl - kp.ey;
}

// STREAM: evict-first stores (a page's pointers are 2-3 MB that nothing reads before the pair's
// traceback, 22 GB per launch of config 2); the line kernels read a pair's few KB back within
// microseconds and store with the default policy so that they are still in L2 then.
template <int C, bool STREAM = true>
__device__ __forceinline__ void store_ptr_words(uint8_t *dst, const unsigned (&pw)[C / 4])
{
    if (C % 16 == 0) {
#pragma unroll
        for (int v = 0; v < C / 16; ++v) {
            const uint4 q = make_uint4(pw[4 * v], pw[4 * v + 1], pw[4 * v + 2], pw[4 * v + 3]);
            if (STREAM) __stcs(reinterpret_cast<uint4 *>(dst) + v, q); else __stcg(reinterpret_cast<uint4 *>(dst) + v, q);
        }
    } else if (C % 8 == 0) {
#pragma unroll
        for (int v = 0; v < C / 8; ++v) {
            const uint2 q = make_uint2(pw[2 * v], pw[2 * v + 1]);
            if (STREAM) __stcs(reinterpret_cast<uint2 *>(dst) + v, q); else __stcg(reinterpret_cast<uint2 *>(dst) + v, q);
        }
    } else {
#pragma unroll
        for (int v = 0; v < C / 4; ++v) {
            if (STREAM) __stcs(reinterpret_cast<unsigned *>(dst) + v, pw[v]); else __stcg(reinterpret_cast<unsigned *>(dst) + v, pw[v]);
        }
    }
}

// Loop-carried state of a pass outside the strip registers.
struct PassState {
    int q_out, y_out;      // right edge of the row this lane computed last
    int q_prev, y_prev;    // what the left neighbour sent in the previous step (row i-1)
    int2 bnext;            // left boundary of the pass for the next step's row (used by lane 0)
    int tnext;             // transcript symbol of the row this lane computes next
    int tnext2;            // ... and of the row after that (chained stripes prefetch two steps ahead)
    int xe, cx;            // ex*i and ox - ex*i of the row this lane computes next
    const uint8_t *tp;     // &T[i] for the next step (row i+1 reads T[i]); byte address, symbols are SYM wide
    const int2 *bp;        // &bnd[t+2]: next boundary row to prefetch
    int2 *bw;              // &bnd[i+1]: where lane 31 leaves its right edge next
    uint8_t *pst;          // pointer bytes of this lane for the next step
    int c0;                // 0-based first column of the strip
    const uint4 *prof;     // SUBST == 2: this lane's column of the warp's query profile
    int2 blk_cur;          // chained passes: current 8-row block of the left boundary, one row per lane 0..7
};

// Chained passes (one huge pair spread over many warps): stripe w leaves its right edge in
// global memory for stripe w+1.  No fences and no progress counters: every row is one 16-byte
// record (Q, epoch, Y, epoch); a record is valid once both halves carry the epoch of the
// current launch (8-byte stores are single transactions, so a torn record shows a stale flag
// in one half and is simply polled again).  Measured: a __threadfence + st.release per 8 rows
// cost more than the stripe's arithmetic.
struct Chain {
    const int4 *in;        // records of the stripe to the left, indexed by row (unused if first)
    int4 *out;             // records this stripe produces
    int epoch;             // nonzero, unique per launch
    // Row bands (pairs whose pointer matrix does not fit the arena): a launch covers rows
    // r0+1 .. r0+n of the matrix, starting from the per-column state saved at row r0.
    int r0;                // rows above the band (0: start from the matrix's row 0)
    const int *ck_in;      // state at row r0: X (real, tagged) [m], D [m], W [m]; null when r0 == 0
    int *ck_out;           // where to leave the state of the band's last row, or null
    int m;                 // columns = stride of the checkpoint arrays
    bool store;            // write pointer bytes (false in the forward checkpointing sweep)
    int *check;            // TANW_CHECKED builds: first failed device assertion
    int4 *stage;           // shared memory: kChainBlock records, where the next block of `in` lands
};
__device__ __forceinline__ Chain no_chain()
{
    Chain c;
    c.in = nullptr; c.out = nullptr; c.epoch = 0;
    c.r0 = 0; c.ck_in = nullptr; c.ck_out = nullptr; c.m = 0; c.store = true; c.check = nullptr; c.stage = nullptr;
    return c;
}
// Hand-over records fetched per asynchronous copy (one per lane 0 .. kChainBlock-1).  A larger block
// adds rows of lag per stripe but halves the take / issue work per row; measured on config 5 /
// a single page: 2 -> 20.6 ms / 0.320 ms, 4 -> 17.2 / 0.272, 8 -> 16.15 / 0.260, 16 -> 15.98 / 0.253,
// 32 -> 16.0 / 0.253.
#ifndef TANW_CHAIN_BLOCK
#define TANW_CHAIN_BLOCK 16
#endif
constexpr int kChainBlock = TANW_CHAIN_BLOCK;

__device__ __forceinline__ int4 ld_volatile_v4(const int4 *p)
{
    int4 v;
    asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_v4(int4 *p, int4 v)
{
    asm volatile("st.volatile.global.v4.s32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Rows base .. base+kChainBlock-1 of the left stripe's edge, one per lane.  chain_issue only starts the
// copy; chain_take, called a block (kChainBlock steps) later, checks the stamps and spins only if the
// producer has not got there yet -- so the L2 round trip overlaps the steps in between
// (ncu: with a blocking fetch a stripe spent a third of its time in this load).  The copy is an
// asynchronous one into shared memory (LDGSTS) rather than a load into registers: ptxas puts every
// global load of the loop on one scoreboard, so the per-step transcript byte -- an L1 hit -- waited
// for the far-L2 round trip of a block fetch issued just before it (18 % of a stripe's time).
__device__ __forceinline__ void chain_issue(const Chain &ch, int base, int n, int lane)
{
    if (lane < kChainBlock && base + lane <= n) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(ch.stage + lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(ch.in + base + lane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ int2 chain_take(const Chain &ch, int base, int n, int lane)
{
    const bool mine = lane < kChainBlock && base + lane <= n;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    int4 v = make_int4(0, ch.epoch, 0, ch.epoch);
    if (mine) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(ch.stage + lane);
        asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src) : "memory");
    }
#ifdef TANW_CHECKED
    const long long t0 = clock64();
#endif
    while (!__all_sync(kFull, v.y == ch.epoch && v.w == ch.epoch)) {
        if (mine) v = ld_volatile_v4(ch.in + base + lane);
#ifdef TANW_CHECKED
        // a record is either untouched by this launch (both stamps stale) or complete, or torn for
        // an instant; a stripe that waits for seconds has lost its producer
        if (clock64() - t0 > (1ll << 33)) { TANW_ASSERT(ch.check, false, 3); break; }
#endif
    }
    TANW_ASSERT(ch.check, !mine || (v.y == ch.epoch && v.w == ch.epoch), 4);
    return make_int2(v.x, v.z);
}

// One wavefront step of one pass.  GUARDED steps check whether the lane's row is inside
// [1, n] (ramp-up / ramp-down) and capture the corner scores; steady steps do neither.
template <int C, bool GUARDED, int SUBST, int VAR, bool CHAINED, typename SYM = uint8_t>
__device__ __forceinline__ void pass_step(Strip<C> &s, PassState &ps, const KParams &kp,
                                          int n, int t, int lane, bool has_next,
                                          int fin_lane, int fin_k, int (&cap)[3], const Chain &ch)
{
    const int i = t - lane;                       // this lane's row in this step
    int q_in = __shfl_up_sync(kFull, ps.q_out, 1);
    int y_in = __shfl_up_sync(kFull, ps.y_out, 1);
    if (lane == 0) { q_in = ps.bnext.x; y_in = ps.bnext.y; }
    if (!GUARDED || t + 1 <= n) {
        if (CHAINED) {
            // A chained stripe has an SM sub-partition almost to itself, so a load issued one
            // step ahead does not hide L2 latency.  The boundary arrives in blocks of kChainBlock rows
            // (one coalesced load per block, fetched a block ahead) and reaches lane 0 by shuffle.
            // Stripe 0 reads column 0 of the matrices the same way (records written by
            // long_col0_kernel), so that the loop has no per-stripe special case.
            const int r = t + 1;                          // boundary row needed by the next step
            const int j = (r - 1) & (kChainBlock - 1);
            if (j == 0) {                         // this block was requested a block ago
                ps.blk_cur = chain_take(ch, r, n, lane);
                if (r + kChainBlock <= n) chain_issue(ch, r + kChainBlock, n, lane);
            }
            ps.bnext.x = __shfl_sync(kFull, ps.blk_cur.x, j);
            ps.bnext.y = __shfl_sync(kFull, ps.blk_cur.y, j);
        } else {
            ps.bnext = __ldcg(ps.bp);             // same address in every lane
        }
    }
    const int dul_in = (VAR == 1 || VAR == 2) ? ps.q_prev : max(ps.q_prev, ps.y_prev);   // D of (i-1, left neighbour column)
    const int tch = ps.tnext;
    if (CHAINED) {
        // a stripe has its scheduler to itself: one step does not cover the load latency
        ps.tnext = ps.tnext2;
        if (!GUARDED || (i + 1 >= 0 && i + 1 < n))
            ps.tnext2 = table_code<SUBST>((int)__ldg(reinterpret_cast<const SYM *>(ps.tp) + 1), kp);   // row i+2 reads T[i+1]
    } else {
        if (!GUARDED || (i >= 0 && i < n))
            ps.tnext = table_code<SUBST>((int)__ldg(reinterpret_cast<const SYM *>(ps.tp)), kp);     // row i+1 reads T[i]
    }
    if (!GUARDED || (i >= 1 && i <= n)) {
        unsigned pw[C / 4];
        const int kfin = (GUARDED && i == n && lane == fin_lane) ? fin_k : -1;
        strip_row<C, GUARDED, SUBST, VAR>(s, kp, tch, ps.xe, ps.cx, q_in, y_in, dul_in,
                                          ps.q_out, ps.y_out, pw, kfin, cap, ps.prof);
        if (!CHAINED || ch.store) store_ptr_words<C>(ps.pst, pw);
        if (has_next && lane == 31) {
            if (CHAINED) st_volatile_v4(ch.out + i, make_int4(ps.q_out, ch.epoch, ps.y_out, ch.epoch));
            else         __stcg(ps.bw, make_int2(ps.q_out, ps.y_out));
        }
        if (CHAINED && GUARDED && ch.ck_out != nullptr && i == n) {
            // last row of a band: leave the per-column state for the launch that continues below
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const int col = ps.c0 + k;
                if (col < ch.m) {
                    ch.ck_out[col] = s.Xh[k] + ps.xe;             // X^ back to the real (tagged) X
                    ch.ck_out[ch.m + col] = s.D[k];
                    if (VAR == 0) ch.ck_out[2 * ch.m + col] = s.W[k];
                }
            }
        }
    }
    ps.q_prev = q_in;
    ps.y_prev = y_in;
    ps.xe += kp.ex;
    ps.cx -= kp.ex;
    ps.tp += sizeof(SYM);
    ps.bp += 1;
    ps.bw += 1;
    ps.pst += 32 * C;
}

#ifndef TANW_STEADY_UNROLL2
#define TANW_STEADY_UNROLL2 1
#endif
#ifndef TANW_CHAINED_UNROLL2
#define TANW_CHAINED_UNROLL2 1
#endif
// One pass: columns [j0, j0 + 32*C) of one pair, all n rows.
//   bnd      : bnd[i], i = 1..n: (Q, Y) of the column left of the pass, row i -- column 0 of
//              the matrices for the first pass (written by the caller, textSeqCompare.py:53-56),
//              the previous pass's right edge otherwise.  If has_next, lane 31 overwrites
//              bnd[i] with this pass's right edge 31 steps after lane 0 consumed it.
//   ptr      : base of this pass's pointer bytes, laid out [step t][lane][C]
//   fin_lane, fin_k : where column m lives in this pass (or fin_lane = -1)
template <int C, int SUBST, int VAR, bool CHAINED, typename SYM = uint8_t>
__device__ __forceinline__ void fill_pass(const KParams &kp, const SYM *__restrict__ T,
                                          const SYM *__restrict__ O, int n, int m, int j0,
                                          bool has_next, const int2 *bnd, int2 *bnd_out,
                                          uint8_t *__restrict__ ptr, int fin_lane, int fin_k,
                                          int (&cap)[3], const Chain &ch, uint4 *prof = nullptr)
{
    const int lane = threadIdx.x & 31;
    const int c0 = j0 + lane * C;                 // 0-based first column of the strip
    Strip<C> s;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int c = c0 + k;
        s.oc[k] = (c < m) ? table_code<SUBST>((int)__ldg(O + c), kp) : (1 << (8 * sizeof(SYM)));   // never equals a symbol
        if (SUBST && c >= m) s.oc[k] = 0;
        // row 0: M[0][j] = X[0][j] = bg*j, Y[0][j] = -inf   (textSeqCompare.py:57-60)
        const int base = kp.bg * (c + 1);
        s.W[k] = base | kTagM;
        s.Xh[k] = base | kTagX;
        s.D[k] = base | kTagM;
        if (CHAINED && ch.ck_in != nullptr && c < m) {        // a band that starts below row 0
            s.Xh[k] = ch.ck_in[c];
            s.D[k] = ch.ck_in[ch.m + c];
            if (VAR == 0) s.W[k] = ch.ck_in[2 * ch.m + c];
        }
    }
    PassState ps;
    ps.prof = nullptr;
    if (SUBST == 2) {
        // the query profile of this pass: scores of every transcript symbol against this lane's columns
        for (int sym = 0; sym < kp.subst_k; ++sym) {
            const int *row = kp.subst + sym * kp.subst_k;
            unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < C; ++k)
                w[k >> 2] |= ((unsigned)(__ldg(row + s.oc[k]) >> kShift) & 0xFFu) << (8 * (k & 3));
            prof[sym * 32 + lane] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
        ps.prof = prof + lane;
    }
    ps.c0 = c0;
    ps.q_out = (kp.bg * (c0 + C)) | kTagM;        // right edge of row 0
    ps.y_out = kNeg;
    ps.q_prev = (kp.bg * c0) | kTagM;             // left neighbour column, row 0
    ps.y_prev = kNeg;                             // (Y[0][j] = -inf, also at j = 0)
    if (CHAINED && ch.ck_in != nullptr) {         // the same two edges at row r0 (D >= Q, Y not needed)
        ps.q_out = (c0 + C - 1 < m) ? ch.ck_in[ch.m + c0 + C - 1] : 0;
        ps.q_prev = (c0 == 0) ? ((kp.bg * ch.r0) | kTagM) : ((c0 - 1 < m) ? ch.ck_in[ch.m + c0 - 1] : 0);
    }
    ps.blk_cur = make_int2(0, 0);
    if (CHAINED) {
        chain_issue(ch, 1, n, lane);
        ps.blk_cur = chain_take(ch, 1, n, lane);                                        // rows 1..8
        if (1 + kChainBlock <= n) chain_issue(ch, 1 + kChainBlock, n, lane);
        ps.bnext.x = __shfl_sync(kFull, ps.blk_cur.x, 0);
        ps.bnext.y = __shfl_sync(kFull, ps.blk_cur.y, 0);
    } else {
        ps.bnext = __ldcg(bnd + 1);
    }
    ps.bp = bnd + 2;
    ps.bw = bnd_out + (1 - lane);
    ps.tnext = (lane == 0) ? table_code<SUBST>((int)__ldg(T), kp) : 0;
    ps.tnext2 = 0;
    if (CHAINED) {
        if (lane == 0 && n > 1) ps.tnext2 = table_code<SUBST>((int)__ldg(T + 1), kp);
        if (lane == 1) ps.tnext2 = table_code<SUBST>((int)__ldg(T), kp);
    }
    ps.tp = reinterpret_cast<const uint8_t *>(T + (1 - lane));
    ps.xe = kp.ex * (1 - lane);                   // row i = t - lane at t = 1
    ps.cx = kp.ox - ps.xe;
    ps.pst = ptr + ((size_t)32 + lane) * C;       // step t = 1

    const int last_step = n + 31;
    const int ramp_end = min(31, last_step);
    int t = 1;
    for (; t <= ramp_end; ++t)                    // ramp-up: lanes join one per step
        pass_step<C, true, SUBST, VAR, CHAINED, SYM>(s, ps, kp, n, t, lane, has_next, fin_lane, fin_k, cap, ch);
    // steady state: every lane on a row in [1, n-1].  Two steps per iteration in the batched
    // kernel: the strip state (D, X^, W per column) otherwise ends every step in other registers
    // than the next one expects, and 27 of the 252 instructions of a C = 16 step were moves.
    if ((!CHAINED || TANW_CHAINED_UNROLL2) && TANW_STEADY_UNROLL2) {
#pragma unroll 1
        for (; t + 1 <= n - 1; t += 2) {
            pass_step<C, false, SUBST, VAR, CHAINED, SYM>(s, ps, kp, n, t, lane, has_next, fin_lane, fin_k, cap, ch);
            pass_step<C, false, SUBST, VAR, CHAINED, SYM>(s, ps, kp, n, t + 1, lane, has_next, fin_lane, fin_k, cap, ch);
        }
    }
#pragma unroll 1
    for (; t <= n - 1; ++t)
        pass_step<C, false, SUBST, VAR, CHAINED, SYM>(s, ps, kp, n, t, lane, has_next, fin_lane, fin_k, cap, ch);
    for (; t <= last_step; ++t)                   // ramp-down: last row, lanes leave one per step
        pass_step<C, true, SUBST, VAR, CHAINED, SYM>(s, ps, kp, n, t, lane, has_next, fin_lane, fin_k, cap, ch);
}

template <int SUBST, int VAR, bool CHAINED, typename SYM = uint8_t>
__device__ __forceinline__ void dispatch_pass(int C, const KParams &kp, const SYM *T,
                                              const SYM *O, int n, int m, int j0,
                                              bool has_next, const int2 *bnd, int2 *bnd_out,
                                              uint8_t *ptr, int fin_lane, int fin_k, int (&cap)[3],
                                              const Chain &ch, uint4 *prof = nullptr)
{
#define TANW_CASE(CC)                                                                              \
    case CC:                                                                                       \
        if constexpr (CC <= kMaxC)                                                                 \
            fill_pass<CC, SUBST, VAR, CHAINED, SYM>(kp, T, O, n, m, j0, has_next, bnd, bnd_out, ptr, \
                                                    fin_lane, fin_k, cap, ch, prof);                \
        break;
    switch (C) {
        TANW_CASE(4) TANW_CASE(8) TANW_CASE(12) TANW_CASE(16)
        TANW_CASE(20) TANW_CASE(24) TANW_CASE(28) TANW_CASE(32)
    }
#undef TANW_CASE
}

// Where the pointer bytes of a pair live: passes of 32*cfull columns (the last one narrower),
// each laid out [step t][lane][C].  cfull is kMaxC for the batched kernel and the stripe width
// chosen by the host for a chained-pass (whole-manuscript) pair.
struct PtrMap {
    long long steps;     // n + 32
    int cfull, passw;    // strip width and columns of a full pass
    int nfull, cr;       // number of full passes, strip width of the remainder pass
    __device__ __forceinline__ PtrMap(int n, int m, int cfull_)
    {
        steps = (long long)n + 32;
        cfull = cfull_;
        passw = 32 * cfull_;
        nfull = m / passw;
        const int r = m % passw;
        cr = r ? remainder_c(r) : 0;
    }
};

// Cursor over the pointer words (4 cells each) of one matrix row, moving towards column 1.
struct PtrCursor {
    long long rowless;   // offset of the word for row 0 (add row * 32 * C)
    int p, ls, k4, C;    // pass, lane of the strip, word inside the strip, strip width
    __device__ __forceinline__ void seek(const PtrMap &map, int word_col)     // word_col = (j-1) >> 2
    {
        const int c = word_col * 4;
        p = c / map.passw;
        C = map.cfull;
        if (p >= map.nfull) { p = map.nfull; C = map.cr; }
        const int cc = c - p * map.passw;
        ls = cc / C;
        k4 = (cc - ls * C) >> 2;
        rebase(map);
    }
    __device__ __forceinline__ void rebase(const PtrMap &map)
    {
        rowless = (long long)p * map.steps * map.passw + (long long)ls * 33 * C + 4 * k4;
    }
    __device__ __forceinline__ long long at_row(int row) const { return rowless + (long long)row * 32 * C; }
    __device__ __forceinline__ void left(const PtrMap &map)                    // one word towards column 1
    {
        if (k4 > 0) { --k4; rowless -= 4; return; }
        if (ls > 0) { --ls; }
        else { --p; C = map.cfull; ls = 31; }      // only ever called while word_col > 0
        k4 = C / 4 - 1;
        rebase(map);
    }
};

constexpr unsigned kGuardWord = 0x80808080u;       // four guard bytes: end a guarded traceback walk
constexpr int kTileRows = 64;                      // rows of a traceback tile (two per lane)
constexpr int kTileWords = 16;                     // 64 columns of a traceback tile
constexpr int kTileStride = kTileWords + 1;        // words per tile row in shared memory (+ pad)

// Traceback of one whole-manuscript pair (textSeqCompare.py:96-164) by one warp (trace_long_kernel).
// Nothing else runs beside it, so both halves of a tile's cost are in the open:
//   * the load of the 64-row x 64-column tile whose bottom-right corner is the current cell: lane l
//     takes word column wq_hi - (l & 15) and rows x - 32*(l >> 4) - r, r = 0..31 -- one cursor seek
//     per lane, then 32 independent loads whose addresses differ by a constant (round 1 walked a
//     cursor along the row for every load: ~700 instructions per lane and tile, as long as the
//     DRAM round trip itself);
//   * the walk inside the tile.  A path is mostly runs: diagonal steps whose pointer says "from
//     M" again, or a long OCR insertion whose pointer says "extend" again.  While the state does
//     not change, the next cell is known, so lane l looks at the l-th cell ahead in the current
//     direction, a ballot finds the first lane whose pointer leaves the state, and the whole run
//     -- up to 32 steps -- is taken at once; its (identical) ops are written by the lanes in
//     parallel.  Round 1 walked one cell per ~70 cycles of dependent instructions.
// Measured and dropped (config 5, 19.9 ms with this form): asking the L2 for the rows above the tile
// (prefetch.global.L2 of rows x-40 .. x-127, 32 word columns) while the tile's own loads are in
// flight -- 21.4 ms: 2 800 scattered prefetches per tile cost more than the round trip they hide;
// the same from a second warp through a shared-memory mailbox -- the spinning warp slows the
// walker's own shared-memory loads.
// Ops are written back to front at the END of the pair's op buffer (capacity n+m); x, y, st, k are
// updated in every lane.  st = -1: take mat_ptr of the start cell first (:102).
__device__ __forceinline__ unsigned lds_u8(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// 32 consecutive rows of one pointer word, going up from the row at p.
template <int C>
__device__ __forceinline__ void tile_rows(const uint8_t *p, unsigned (&w)[32])
{
#pragma unroll
    for (int r = 0; r < 32; ++r) w[r] = __ldcg(reinterpret_cast<const unsigned *>(p - r * 32 * C));
}

__device__ __forceinline__ void traceback_core(const uint8_t *ptr, int n, int m, int cfull,
                                               uint8_t *ops_end, unsigned *tile, int lane,
                                               int &x, int &y, int &st, int &k)
{
    const PtrMap map(n, m, cfull);
    // the tile's byte address in the shared window, taken once: left to the compiler, the walk loop
    // re-derives it from the generic pointer in every iteration (S2UR + ULEA on its dependent chain)
    const unsigned tb = (unsigned)__cvta_generic_to_shared(tile);
    while (x > 0 && y > 0) {
        // ---- tile load: word columns wq_hi-15 .. wq_hi, rows x .. x-63 ---------------------------
        const int wq_hi = (y - 1) >> 2;
        const int q = 15 - (lane & 15);                      // this lane's word column of the tile
        const int wq = wq_hi - (lane & 15);
        const int row_hi = x - 32 * (lane >> 4);             // its first row (it goes up from there)
        unsigned w[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) w[r] = 0u;
        if (wq >= 0) {
            PtrCursor cur;
            cur.seek(map, wq);
            const uint8_t *p = ptr + cur.at_row(row_hi);           // the word of row_hi; row - 1 is 32 * C bytes before
            if (x >= kTileRows) {
                // (warp-uniform) all 64 rows are inside the matrix: the strip width as a compile-time
                // constant turns the 32 addresses into immediate offsets of one base register
                // (the generic loop below spends ten instructions per load on 64-bit multiplies)
                switch (cur.C) {
                case 4:  tile_rows<4>(p, w); break;
                case 8:  tile_rows<8>(p, w); break;
                case 12: tile_rows<12>(p, w); break;
                default: tile_rows<16>(p, w); break;
                }
            } else {
                const int step = 32 * cur.C;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    if (row_hi - r >= 1) w[r] = __ldcg(reinterpret_cast<const unsigned *>(p));
                    p -= step;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[(32 * (lane >> 4) + r) * kTileStride + q] = w[r];
        __syncwarp();
        // (an L2 prefetch of the tile the walk will most likely enter next, issued here, changes
        // nothing: 17.15 ms with and without on config 5 -- the tile phase is not DRAM latency)
        // ---- walk inside the tile, a run per iteration -----------------------------------------------
        // tile row R = matrix row x - R; tile byte column c = matrix column (col_lo + c + 1)
        const int col_lo = (wq_hi - (kTileWords - 1)) * 4;          // 0-based matrix column of tile byte 0
        const int rows_avail = min(x, kTileRows);
        const int c_min = max(0, -col_lo);                          // tile column of matrix column 1
        int R = 0, c = (y - 1) - col_lo;
        if (st < 0) st = 2 - (int)(lds_u8(tb + c) & 3u);                                           // :102
        for (;;) {
            const int dx = (st != 2), dy = (st != 1);
            const int Rl = R + lane * dx, cl = c - lane * dy;
            const bool valid = Rl < rows_avail && cl >= c_min;
            int nxt = -1;
            if (valid) nxt = 2 - (int)((lds_u8(tb + Rl * (kTileStride * 4) + cl) >> (2 * st)) & 3u);
            const unsigned same = __ballot_sync(kFull, valid && nxt == st);
            const unsigned val = __ballot_sync(kFull, valid);
            const int run = (same == kFull) ? 32 : __ffs(~same) - 1; // leading lanes that stay in the state
            const bool turns = run < 32 && ((val >> run) & 1u);      // the run ends on a cell that changes the state
            const int steps = turns ? run + 1 : run;
            if (steps == 0) break;                                   // the next cell is outside the tile
            if (lane < steps) *(ops_end - (k + 1 + lane)) = (uint8_t)st;                  // :115-145
            k += steps;
            R += steps * dx;
            c -= steps * dy;
            if (turns) st = __shfl_sync(kFull, nxt, run);
            if (R >= rows_avail || c < c_min) break;
        }
        x -= R;
        y = col_lo + c + 1;
    }
}

// The same walk for the batched page kernel, where the one-lane loop costs issue slots that the
// other warps of the scheduler could use (3.5 % of the kernel's instructions): guarded tile,
// decoded bytes and a permute-based move as in traceback_groups -- 13 instructions per path step
// instead of 25 (+0.9 % on config 2).  The lone warp of a chained pair keeps traceback_core: there
// the loop branch of this form has to wait for each step's own load (measured 5 % slower).
// `tile` needs (kTileRows + 1) * kTileStride words.
__device__ __forceinline__ void traceback_core_guarded(const uint8_t *ptr, int n, int m, int cfull,
                                               uint8_t *ops_end, unsigned *tile, int lane,
                                               int &x, int &y, int &st, int &k)
{
    const PtrMap map(n, m, cfull);
    if (lane < kTileStride) tile[kTileRows * kTileStride + lane] = kGuardWord;       // guard row above the tile
    while (x > 0 && y > 0) {
        // ---- tile load: word columns wq_hi-15 .. wq_hi, rows x-lane and x-32-lane ----------
        const int wq_hi = (y - 1) >> 2;
        const int row0 = x - lane, row1 = x - 32 - lane;
        unsigned w0[kTileWords], w1[kTileWords];
        PtrCursor cur;
        cur.seek(map, wq_hi);
#pragma unroll
        for (int q = kTileWords - 1; q >= 0; --q) {
            const int wq = wq_hi - (kTileWords - 1 - q);
            // outside the matrix: becomes a guard word (no pointer word is all ones: tags are <= 2).
            // Nothing here may consume a loaded value, or the 32 loads of a lane would serialise.
            w0[q] = 0xFFFFFFFFu; w1[q] = 0xFFFFFFFFu;
            if (wq >= 0) {
                if (row0 >= 1) w0[q] = __ldcg(reinterpret_cast<const unsigned *>(ptr + cur.at_row(row0)));
                if (row1 >= 1) w1[q] = __ldcg(reinterpret_cast<const unsigned *>(ptr + cur.at_row(row1)));
                if (wq > 0) cur.left(map);
            }
        }
        __syncwarp();
        // Tile row r = matrix row x - r: one guard word, then 16 words of DECODED bytes.
        //   pointer byte b = tagM | tagX << 2 | tagY << 4 (tags 2 / 1 / 0 = came from M / X / Y);
        //   walk state st = 2 - tag (0 diagonal, 1 x-gap, 2 y-gap; :110-145), kept as s = 2*st;
        //   decoded byte d = (0x2A - b) << 1 holds 2*(2 - tag) per field: next state s' = (d >> s) & 6;
        //   guard bytes are 0x80 (a decoded byte is at most 0x54): the walk stops on them.
        tile[lane * kTileStride] = kGuardWord;
        tile[(lane + 32) * kTileStride] = kGuardWord;
#pragma unroll
        for (int q = 0; q < kTileWords; ++q) {
            tile[lane * kTileStride + 1 + q] =
                (w0[q] == 0xFFFFFFFFu) ? kGuardWord : (0x2A2A2A2Au - (w0[q] & 0x3F3F3F3Fu)) << 1;
            tile[(lane + 32) * kTileStride + 1 + q] =
                (w1[q] == 0xFFFFFFFFu) ? kGuardWord : (0x2A2A2A2Au - (w1[q] & 0x3F3F3F3Fu)) << 1;
        }
        __syncwarp();
        // ---- walk inside the tile ---------------------------------------------------------
        // The move comes out of a byte permute on the state (+67 diagonal, +68 up, -1 left in a
        // tile of 68-byte rows) and the guards replace the row / column counters.
        const int col_lo = (wq_hi - (kTileWords - 1)) * 4;          // 0-based column of the first data byte
        int cnt = 0;
        if (lane == 0) {
            const unsigned char *tb = reinterpret_cast<const unsigned char *>(tile);
            int off = 4 + (y - 1) - col_lo;                         // row 0 of the tile
            unsigned d = tb[off];
            int s = 2 * st;
            if (st < 0) s = (int)(d & 6u);                          // state from mat_ptr first   (:102)
            uint8_t *op = ops_end - k;                              // ops go out back to front
            while (!(d & 0x80u)) {
                *--op = (uint8_t)(s >> 1);
                ++cnt;
                // byte s of {67, 0, 68, 0, -1, -1} with byte s+1 as its sign extension
                // (prmt directly: __byte_perm would first mask the selector, one more op on the chain)
                int delta;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(delta)
                    : "r"(0x00440043), "r"(0x0000FFFF), "r"(s * 0x1111 + 0x1110));
                off += delta;
                s = (int)((d >> s) & 6u);
                d = tb[off];
            }
            const int r = off / (kTileStride * 4);
            x -= r;
            y = col_lo + (off - r * (kTileStride * 4)) - 3;
            st = s >> 1;
        }
        x = __shfl_sync(kFull, x, 0);
        y = __shfl_sync(kFull, y, 0);
        st = __shfl_sync(kFull, st, 0);
        cnt = __shfl_sync(kFull, cnt, 0);
        k += cnt;
    }
}

// Traceback of one pair (textSeqCompare.py:96-164) by one warp.  The pointer chase is a chain
// of dependent loads, so the warp first pulls the 64-row x 64-column tile of pointer bytes
// whose bottom-right corner is the current cell into shared memory (lane r: rows x-r and
// x-32-r, 32 independent word loads), then lane 0 walks inside the tile.  Ops are written back
// to front at the END of the pair's op buffer (capacity n+m); returns the number of columns
// (valid in every lane).
__device__ __forceinline__ int traceback_warp(const uint8_t *ptr, int n, int m, int cfull,
                                              uint8_t *ops_end, unsigned *tile, int lane)
{
    int x = n, y = m, k = 0, st = -1;
    if (n > 0 && m > 0) traceback_core_guarded(ptr, n, m, cfull, ops_end, tile, lane, x, y, st, k);
    if (lane == 0) {
        while (y > 0) { ++k; *(ops_end - k) = 2; --y; }      // OCR remainder first       (:154-158)
        while (x > 0) { ++k; *(ops_end - k) = 1; --x; }      // then transcript remainder (:160-164)
    }
    __syncwarp();                       // lane 0's op bytes are read by every lane next
    return __shfl_sync(kFull, k, 0);
}

__device__ __forceinline__ int score_out(int v)
{
    return (v <= kNeg / 2) ? kNeg : (v >> kShift);
}

#ifndef TANW_MINB
#define TANW_MINB 4
#endif
constexpr int kSubstSmemK = 64;      // substitution tables up to this side are staged in shared memory (16 KB)

// A tabulated scorer's K x K table (textSeqCompare.py:27-29), staged in shared memory when it
// fits: every cell reads one entry, and the whole block shares the table.  Returns the KParams the
// block uses from here on (subst points at the staged copy).  All threads of the block call it.
__device__ __forceinline__ KParams stage_subst(const KParams &kp, int *stab)
{
    KParams out = kp;
    if (kp.subst != nullptr && kp.subst_k <= kSubstSmemK) {
        for (int i = threadIdx.x; i < kp.subst_k * kp.subst_k; i += blockDim.x) stab[i] = __ldg(kp.subst + i);
        __syncthreads();
        out.subst = stab;
    }
    return out;
}

// SYM: uint8_t symbol codes, or uint16_t for pairs with more than 256 distinct elements
// (tanw_set_symbol_bytes); offsets in PairDesc count symbols.
// MULTI: every pair names its own scoring system (BatchArgs::kparams / sidx) -- the reference's
// parameter sweep, 729 vectors over the same pages (evaluate_text_alignment.py:134-194), as ONE
// launch; VAR is then the most general variant any system of the batch needs.
template <int SUBST, int VAR, typename SYM = uint8_t, bool MULTI = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, TANW_MINB)
align_pairs_kernel(const BatchArgs a, const __grid_constant__ KParams kp_launch)
{
    // SUBST == 2: a warp's query profile and, after the fill, its traceback tile share one region
    // of dynamic shared memory (profile_bytes_per_warp); otherwise the tiles are static
    __shared__ unsigned tiles_static[SUBST == 2 ? 1 : kWarpsPerBlock][SUBST == 2 ? 1 : (kTileRows + 1) * kTileStride];
    __shared__ int stab[SUBST == 1 ? kSubstSmemK * kSubstSmemK : 1];
    extern __shared__ uint4 dyn_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * kWarpsPerBlock + warp;
    uint8_t *const ptr = a.ptr_arena + (size_t)slot * (size_t)a.slot_bytes;
    int2 *const bnd = a.bnd_arena + (size_t)slot * (size_t)a.bnd_rows;
    KParams kp = kp_launch;
    if (SUBST == 1) kp = stage_subst(kp_launch, stab);
    uint4 *const prof = SUBST == 2 ? dyn_smem + (size_t)warp * (profile_bytes_per_warp(kp_launch.subst_k) / 16) : nullptr;
    unsigned *const tile = SUBST == 2 ? reinterpret_cast<unsigned *>(prof) : tiles_static[SUBST == 2 ? 0 : warp];

    for (;;) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(a.counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= (unsigned)a.n_pairs) break;
        const int p = a.order[idx];
        const PairDesc pd = a.pairs[p];
        const int n = pd.n, m = pd.m;
        if (MULTI) kp = a.kparams[a.sidx[p]];
        TANW_ASSERT(a.check, ptr_bytes(n, m) <= a.slot_bytes && n + 2 <= a.bnd_rows, 1);
        const SYM *T = reinterpret_cast<const SYM *>(a.sym) + pd.t_off;
        const SYM *O = reinterpret_cast<const SYM *>(a.sym) + pd.o_off;
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
            // left boundary of the first pass = column 0: M = Y = bg*i, X = -inf   (:54-56)
            for (int i = 1 + lane; i <= n + 1; i += 32)
                __stcg(bnd + i, make_int2((kp.bg * i) | kTagM, kp.bg * i));
            __syncwarp();
            for (int ps = 0; ps < npass; ++ps) {
                const int C = (ps < nfull) ? kMaxC : remainder_c(r);
                const int j0 = ps * kPassW;
                const bool last = (ps == npass - 1);
                const int cc = m - 1 - j0;
                const int fin_lane = last ? cc / C : -1;
                const int fin_k = last ? cc % C : -1;
                dispatch_pass<SUBST, VAR, false, SYM>(C, kp, T, O, n, m, j0, !last, bnd, bnd,
                                                 ptr + (size_t)ps * (size_t)pass_bytes, fin_lane, fin_k, cap,
                                                 no_chain(), prof);
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
        const int L = traceback_warp(ptr, n, m, kMaxC, ops + (size_t)n + (size_t)m, tile, lane);
        TANW_ASSERT(a.check, L >= max(n, m) && L <= n + m, 2);
        if (lane == 0) {
            a.ops_len[p] = L;
            if (a.scores) {
                a.scores[3 * (size_t)p + 0] = score_out(cap[0]);
                a.scores[3 * (size_t)p + 1] = score_out(cap[1]);
                a.scores[3 * (size_t)p + 2] = score_out(cap[2]);
            }
        }
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

// ---- short pairs: four pairs per warp (BASELINE config 3, 40-120 character lines) -------------
// A 32-lane wavefront on an 80-column pair is mostly ramp.  Here a pair gets 8 lanes (one
// group), C = 4..16 columns per lane (m <= 128, a single pass, column 0 synthesised in
// registers), so the skew costs 7 steps instead of 31, and a warp works on a "quad" of four
// pairs of the same strip width and similar height at once.  Same strip_row inner loop, same
// pointer bytes ([step][group lane][C]); the traceback runs for the four pairs concurrently,
// each group pulling 8-row x 2-strip tiles into shared memory.
constexpr int kLineG = 8;
constexpr int kLineMaxC = 16;
constexpr int kLineMaxM = kLineG * kLineMaxC;      // 128 columns
constexpr int kLineMaxN = 4096;                    // bounds the per-group pointer slot
constexpr int kLineTile = 2 * kLineMaxC / 4 + 1;   // words per tile row (guard word + two strips)

__host__ __device__ inline int line_c(int m) { return m <= 0 ? 4 : ((m + 31) / 32) * 4; }
__host__ __device__ inline long long line_ptr_bytes(int n, int m)
{
    return (n <= 0 || m <= 0) ? 0 : ((long long)n + 8) * kLineG * line_c(m);
}

// Line pairs of a chunk, sorted by descending (strip-width class, n) on the device
// (tanw_tables.cuh): class c = pairs with strip width 4*(c+1) (cell-less pairs count as class 0).
// Quad q of a class = its entries 4q .. 4q+3; quads never mix classes.
struct LineClasses {
    int start[4], count[4];      // where a class begins in the sorted list, and its size
    int quad0[4];                // quads before the class (classes are laid out 3, 2, 1, 0)
    int n_quads;
};

struct LineArgs {
    const uint8_t  *sym;
    const PairDesc *pairs;
    const int      *sorted;      // the chunk's line pairs, sorted
    const LineClasses *classes;
    unsigned       *counter;
    int             n_quads;     // work units of this launch (quads; octets of the 16-bit kernel)
    uint8_t        *ptr_arena;   // slot_bytes per 8-lane group (16-bit kernel: per pair of a group)
    long long       slot_bytes;
    uint8_t        *ops;
    int            *ops_len;
    int            *scores;
    int            *check;       // TANW_CHECKED builds: first failed device assertion (0 = none)
};

struct LineState {
    int q_out, y_out, q_prev, y_prev, tnext, xe, cx, bq;
    const uint8_t *tp;
    uint8_t *pst;
};

template <int C, bool GUARDED, int SUBST, int VAR>
__device__ __forceinline__ void line_step(Strip<C> &s, LineState &ls, const KParams &kp, int n, bool act,
                                          int t, int gl, int fin_lane, int fin_k, int (&cap)[3])
{
    const int i = t - gl;
    int q_in = __shfl_up_sync(kFull, ls.q_out, 1, kLineG);
    int y_in = __shfl_up_sync(kFull, ls.y_out, 1, kLineG);
    if (gl == 0) { q_in = ls.bq | kTagM; y_in = ls.bq; }        // column 0: M = Y = bg*i (:54-56)
    const int dul_in = (VAR >= 1) ? ls.q_prev : max(ls.q_prev, ls.y_prev);
    const int tch = ls.tnext;
    if (!GUARDED || (i >= 0 && i < n)) ls.tnext = table_code<SUBST>((int)__ldg(ls.tp), kp);
    if (!GUARDED || (act && i >= 1 && i <= n)) {
        unsigned pw[C / 4];
        const int kfin = (GUARDED && i == n && gl == fin_lane) ? fin_k : -1;
        strip_row<C, GUARDED, SUBST, VAR>(s, kp, tch, ls.xe, ls.cx, q_in, y_in, dul_in,
                                          ls.q_out, ls.y_out, pw, kfin, cap);
        store_ptr_words<C>(ls.pst, pw);
    }
    ls.q_prev = q_in;
    ls.y_prev = y_in;
    ls.xe += kp.ex;
    ls.cx -= kp.ex;
    ls.bq += kp.bg;
    ls.tp += 1;
    ls.pst += kLineG * C;
}

// Traceback of the (up to) four pairs of a quad, one per 8-lane group, concurrently.
template <int C>
__device__ __forceinline__ int traceback_groups(const uint8_t *ptr, int n, int m, bool act,
                                                uint8_t *ops_end, unsigned *tile, int gl)
{
    // Tile of a group: 8 rows (row r = matrix row x - r) plus a guard row; every row is a guard
    // word followed by the DECODED pointer bytes of two strips.  The walk is 3.6 % of the page
    // kernel's instructions but 12 % of this kernel's (a tile is only eight rows high), so its
    // loop is kept to eleven instructions:
    //   pointer byte b = tagM | tagX << 2 | tagY << 4 (tags 2 / 1 / 0 = came from M / X / Y);
    //   walk state st = 2 - tag (0 diagonal, 1 x-gap, 2 y-gap; :110-145), kept as s = 2*st;
    //   decoded byte d = (0x2A - b) << 1 holds 2*(2 - tag) per field: next state s' = (d >> s) & 6;
    //   guard bytes are 0x80 (a decoded byte is at most 0x54) and end the walk instead of row /
    //   column counters; the move (+35 diagonal, +36 up, -1 left) is one byte permute on s.
    int x = n, y = m, k = 0, st = -1;
    tile[kLineG * kLineTile + gl] = kGuardWord;           // guard row: words 0..7 ...
    if (gl == 0) tile[kLineG * kLineTile + kLineG] = kGuardWord;                 // ... and 8
    while (__any_sync(kFull, act && x > 0 && y > 0)) {
        const bool mine = act && x > 0 && y > 0;
        const int sidx = mine ? (y - 1) / C : 0;             // strip that holds column y
        const int row = x - gl;
        unsigned w[2 * C / 4];
#pragma unroll
        for (int q = 0; q < 2 * C / 4; ++q) w[q] = 0xFFFFFFFFu;     // outside the matrix (no pointer word is all ones)
        if (mine && row >= 1) {
            const unsigned *hi = reinterpret_cast<const unsigned *>(ptr + ((size_t)(row + sidx) * kLineG + sidx) * C);
#pragma unroll
            for (int q = 0; q < C / 4; ++q) w[C / 4 + q] = __ldcg(hi + q);
            if (sidx >= 1) {
                const unsigned *lo = reinterpret_cast<const unsigned *>(ptr + ((size_t)(row + sidx - 1) * kLineG + sidx - 1) * C);
#pragma unroll
                for (int q = 0; q < C / 4; ++q) w[q] = __ldcg(lo + q);
            }
        }
        __syncwarp();
        tile[gl * kLineTile] = kGuardWord;
#pragma unroll
        for (int q = 0; q < 2 * C / 4; ++q)
            tile[gl * kLineTile + 1 + q] = (w[q] == 0xFFFFFFFFu) ? kGuardWord : (0x2A2A2A2Au - (w[q] & 0x3F3F3F3Fu)) << 1;
        __syncwarp();
        if (mine && gl == 0) {
            const unsigned char *tb = reinterpret_cast<const unsigned char *>(tile);
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
        x = __shfl_sync(kFull, x, 0, kLineG);
        y = __shfl_sync(kFull, y, 0, kLineG);
    }
    if (gl == 0) {
        while (y > 0) { ++k; *(ops_end - k) = 2; --y; }      // :154-158
        while (x > 0) { ++k; *(ops_end - k) = 1; --x; }      // :160-164
    }
    __syncwarp();                       // the group's op bytes are read by all its lanes next
    return __shfl_sync(kFull, k, 0, kLineG);
}

template <int C, int SUBST, int VAR>
__device__ __forceinline__ void line_quad(const LineArgs &a, const KParams &kp, int p, uint8_t *ptr,
                                          unsigned *tile, int lane)
{
    const int gl = lane & (kLineG - 1);
    int n = 0, m = 0;
    long long ops_off = 0;
    const uint8_t *T = a.sym, *O = a.sym;
    if (p >= 0) {
        const PairDesc pd = a.pairs[p];
        n = pd.n; m = pd.m; ops_off = pd.ops_off;
        T = a.sym + pd.t_off; O = a.sym + pd.o_off;
    }
    const bool act = (n > 0 && m > 0);
    if (!act) { T = a.sym; O = a.sym; }
    TANW_ASSERT(a.check, !act || (line_ptr_bytes(n, m) <= a.slot_bytes && m <= kLineG * C), 5);
    // tallest / shortest active pair of the quad (n is uniform inside a group)
    int nmax = act ? n : 0, nmin = act ? n : 0x7fffffff;
#pragma unroll
    for (int d = kLineG; d < 32; d <<= 1) {
        nmax = max(nmax, __shfl_xor_sync(kFull, nmax, d));
        nmin = min(nmin, __shfl_xor_sync(kFull, nmin, d));
    }
    int cap[3];
    cap[0] = kp.bg * (n > 0 ? n : m);                        // corner scores without any cell (:53-60)
    cap[1] = (n > 0) ? kNeg : kp.bg * m;
    cap[2] = (n > 0) ? kp.bg * n : kNeg;
    if (n == 0 && m == 0) { cap[0] = 0; cap[1] = 0; cap[2] = kNeg; }

    if (nmax > 0) {
        const int c0 = gl * C;
        Strip<C> s;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int c = c0 + k;
            s.oc[k] = (act && c < m) ? table_code<SUBST>((int)__ldg(O + c), kp) : (SUBST ? 0 : 0x100);
            const int base = kp.bg * (c + 1);                // row 0 (:57-60)
            s.W[k] = base | kTagM;
            s.Xh[k] = base | kTagX;
            s.D[k] = base | kTagM;
        }
        LineState ls;
        ls.q_out = (kp.bg * (c0 + C)) | kTagM;
        ls.y_out = kNeg;
        ls.q_prev = (kp.bg * c0) | kTagM;
        ls.y_prev = kNeg;
        ls.tnext = (gl == 0) ? table_code<SUBST>((int)__ldg(T), kp) : 0;
        ls.tp = T + (1 - gl);
        ls.xe = kp.ex * (1 - gl);
        ls.cx = kp.ox - ls.xe;
        ls.bq = kp.bg * (1 - gl);
        ls.pst = ptr + ((size_t)kLineG + gl) * C;            // step t = 1
        const int fin_lane = act ? (m - 1) / C : -1;
        const int fin_k = act ? (m - 1) % C : -1;
        const int last_step = nmax + kLineG - 1;
        int t = 1;
        for (; t <= min(kLineG - 1, last_step); ++t)         // ramp-up
            line_step<C, true, SUBST, VAR>(s, ls, kp, n, act, t, gl, fin_lane, fin_k, cap);
        // every lane of every group on a row in [1, n-1]; two steps per iteration (see fill_pass)
#pragma unroll 1
        for (; t + 1 <= nmin - 1; t += 2) {
            line_step<C, false, SUBST, VAR>(s, ls, kp, n, act, t, gl, fin_lane, fin_k, cap);
            line_step<C, false, SUBST, VAR>(s, ls, kp, n, act, t + 1, gl, fin_lane, fin_k, cap);
        }
#pragma unroll 1
        for (; t <= nmin - 1; ++t)
            line_step<C, false, SUBST, VAR>(s, ls, kp, n, act, t, gl, fin_lane, fin_k, cap);
        for (; t <= last_step; ++t)                          // ramp-down and the taller pairs' tails
            line_step<C, true, SUBST, VAR>(s, ls, kp, n, act, t, gl, fin_lane, fin_k, cap);
        {   // the lane of the group that owns column m holds the corner scores
            const int src = (lane & ~(kLineG - 1)) + (act ? fin_lane : 0);
            const int v0 = __shfl_sync(kFull, cap[0], src);
            const int v1 = __shfl_sync(kFull, cap[1], src);
            const int v2 = __shfl_sync(kFull, cap[2], src);
            if (act) { cap[0] = v0; cap[1] = v1; cap[2] = v2; }
        }
    }
    __syncwarp();
    uint8_t *ops = a.ops + ops_off;
    const int L = traceback_groups<C>(ptr, n, m, act, ops + (size_t)n + (size_t)m, tile, gl);
    if (p >= 0 && gl == 0) {
        a.ops_len[p] = L;
        if (a.scores) {
            a.scores[3 * (size_t)p + 0] = score_out(cap[0]);
            a.scores[3 * (size_t)p + 1] = score_out(cap[1]);
            a.scores[3 * (size_t)p + 2] = score_out(cap[2]);
        }
    }
    // move each op string to the start of its buffer
    const int shift = (p >= 0) ? n + m - L : 0;
    int rounds = (shift > 0) ? (L + kLineG - 1) / kLineG : 0;
#pragma unroll
    for (int d = kLineG; d < 32; d <<= 1) rounds = max(rounds, __shfl_xor_sync(kFull, rounds, d));
    for (int it = 0; it < rounds; ++it) {
        const int q = it * kLineG + gl;
        uint8_t v = 0;
        const bool on = shift > 0 && q < L;
        if (on) v = __ldcg(ops + shift + q);
        __syncwarp();
        if (on) ops[q] = v;
        __syncwarp();
    }
}

template <int SUBST, int VAR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, TANW_MINB)
align_lines_kernel(const LineArgs a, const __grid_constant__ KParams kp_launch)
{
    __shared__ unsigned tiles[kWarpsPerBlock][4 * (kLineG + 1) * kLineTile];
    __shared__ int stab[SUBST ? kSubstSmemK * kSubstSmemK : 1];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane >> 3;
    const long long slot = ((long long)blockIdx.x * kWarpsPerBlock + warp) * 4 + g;
    uint8_t *const ptr = a.ptr_arena + (size_t)slot * (size_t)a.slot_bytes;
    unsigned *const tile = tiles[warp] + g * ((kLineG + 1) * kLineTile);
    const LineClasses lc = *a.classes;
    KParams kp = kp_launch;
    if (SUBST) kp = stage_subst(kp_launch, stab);
    for (;;) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(a.counter, 1u);
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= (unsigned)a.n_quads) break;
        int cls = 3;                                         // the class this quad belongs to
        if ((int)idx >= lc.quad0[2]) cls = 2;
        if ((int)idx >= lc.quad0[1]) cls = 1;
        if ((int)idx >= lc.quad0[0]) cls = 0;
        const int e = 4 * ((int)idx - lc.quad0[cls]) + g;    // this group's entry of the class
        const int p = (e < lc.count[cls]) ? a.sorted[lc.start[cls] + e] : -1;
        const int C = 4 * (cls + 1);                         // uniform: a quad holds one strip width
        switch (C) {
        case 4:  line_quad<4,  SUBST, VAR>(a, kp, p, ptr, tile, lane); break;
        case 8:  line_quad<8,  SUBST, VAR>(a, kp, p, ptr, tile, lane); break;
        case 12: line_quad<12, SUBST, VAR>(a, kp, p, ptr, tile, lane); break;
        default: line_quad<16, SUBST, VAR>(a, kp, p, ptr, tile, lane); break;
        }
        __syncwarp();
    }
}

// ---- one huge pair: chained passes ---------------------------------------------------------------
// BASELINE config 5 (100k x 80k): a single pair must use the whole GPU.  Every pass (a stripe
// of 32*C columns) gets its own warp, all passes are resident at once (cooperative launch), and
// pass w consumes the right edge of pass w-1 through global memory a few rows
// behind it -- a wavefront over stripes.  Pointers for the whole matrix stay in HBM
// (1 byte/cell; 8 GB for config 5), and the ordinary tile-prefetch traceback runs afterwards.
struct LongArgs {
    const uint8_t *T, *O;  // byte addresses; symbols are SYM wide
    int n, m;              // the whole pair
    int r0, nb;            // this launch: rows r0+1 .. r0+nb (one band; r0 = 0, nb = n without banding)
    uint8_t *ptr;          // ptr_bytes(nb, m, cfull) when store != 0
    int4 *chain;           // npass arrays of chain_stride records; array w = right edge of stripe w
    long long chain_stride;
    int epoch;             // nonzero, unique per launch within the context
    int pass0;             // first stripe handled by this launch (waves when stripes > resident warps)
    int cfull;             // stripe strip width: a full stripe has 32*cfull columns
    const int *ck_in;      // per-column state at row r0 (3*m ints) or null when r0 == 0
    int *ck_out;           // where to leave the state at row r0+nb (3*m ints) or null
    int store;             // write pointer bytes
    int *scores;           // 3 ints, written by the launch that contains row n (or null)
    int *check;            // TANW_CHECKED builds: first failed device assertion
};

template <int SUBST, int VAR, typename SYM = uint8_t>
__global__ void __launch_bounds__(32, 8)
align_long_kernel(const LongArgs a, const __grid_constant__ KParams kp)
{
    const int w = a.pass0 + blockIdx.x;
    const int n = a.nb, m = a.m;
    const int passw = 32 * a.cfull;
    const int nfull = m / passw, r = m % passw;
    const int npass = nfull + (r ? 1 : 0);
    if (w >= npass) return;
    const int C = (w < nfull) ? a.cfull : remainder_c(r);
    const int j0 = w * passw;
    const bool last = (w == npass - 1);
    const int cc = m - 1 - j0;
    const int fin_lane = last ? cc / C : -1;
    const int fin_k = last ? cc % C : -1;
    int cap[3] = {0, 0, 0};
    const long long pass_bytes = ((long long)n + 32) * passw;
    Chain ch;
    ch.in = a.chain + (size_t)(w > 0 ? w - 1 : npass) * (size_t)a.chain_stride;   // array npass: column 0
    ch.out = a.chain + (size_t)w * (size_t)a.chain_stride;
    ch.epoch = a.epoch;
    ch.r0 = a.r0;
    ch.ck_in = a.ck_in;
    ch.ck_out = a.ck_out;
    ch.m = m;
    ch.store = a.store != 0;
    ch.check = a.check;
    __shared__ __align__(16) int4 chain_stage[kChainBlock];
    ch.stage = chain_stage;
    // Launch-constant operands of the steady loop, held in registers: left as kernel parameters
    // they are re-read from the constant bank (5 LDCU + 1 LDC at the top of every row) and the
    // row's dependent chain -- what bounds a stripe -- waits for them.  A round trip through shared
    // memory makes them opaque to ptxas (config 5: 19.89 -> 18.83 ms).
    KParams kq = kp;
    __shared__ unsigned long long regw[8];
    __shared__ int regi[12];
    const SYM *Tp = reinterpret_cast<const SYM *>(a.T) + a.r0;
    const SYM *Op = reinterpret_cast<const SYM *>(a.O);
    uint8_t *pp = a.ptr + (size_t)w * (size_t)pass_bytes;
    int nn = n, mm = m;
    {
        if ((threadIdx.x & 31) == 0) {
            regi[0] = kp.maT; regi[1] = kp.miT; regi[2] = kp.ox; regi[3] = kp.ex; regi[4] = kp.oy; regi[5] = kp.ey;
            regi[6] = ch.epoch; regi[7] = ch.store ? 1 : 0; regi[8] = n; regi[9] = m;
            regw[0] = (unsigned long long)ch.in; regw[1] = (unsigned long long)ch.out;
            regw[2] = (unsigned long long)Tp; regw[3] = (unsigned long long)Op; regw[4] = (unsigned long long)pp;
        }
        __syncwarp();
        const volatile int *vi = regi;
        const volatile unsigned long long *vw = regw;
        kq.maT = vi[0]; kq.miT = vi[1]; kq.ox = vi[2]; kq.ex = vi[3]; kq.oy = vi[4]; kq.ey = vi[5];
        ch.epoch = vi[6]; ch.store = vi[7] != 0;
        ch.in = (const int4 *)vw[0]; ch.out = (int4 *)vw[1];
        nn = vi[8]; mm = vi[9];
        Tp = (const SYM *)vw[2]; Op = (const SYM *)vw[3]; pp = (uint8_t *)vw[4];
    }
    dispatch_pass<SUBST, VAR, true, SYM>(C, kq, Tp, Op, nn, mm, j0, !last, nullptr, nullptr, pp, fin_lane, fin_k,
                                         cap, ch);
    if (last && (int)(threadIdx.x & 31) == fin_lane && a.scores && a.r0 + a.nb == a.n) {
        a.scores[0] = score_out(cap[0]);
        a.scores[1] = score_out(cap[1]);
        a.scores[2] = score_out(cap[2]);
    }
}

// ---- int32 issue-rate micro-benchmark (roofline denominator) -----------------
// 16 independent chains per thread, operands loaded from memory so ptxas cannot fold them.
//   WHICH 0: two add.s32 per chain-iteration, which ptxas merges into ONE 3-input IADD3
//   WHICH 1: max.s32 + min.s32  -> two VIMNMX
//   WHICH 2: fused add+max, add+min -> two VIADDMNMX
//   WHICH 3: add of a uniform (kernel-parameter) value + xor -> one VIADD, which either integer
//            pipe executes, + one LOP3 (alu pipe): both pipes busy, i.e. the warp-instruction
//            issue rate -- the ceiling of a kernel that mixes the pipes as strip_row does.
//            (VIMNMX + a three-register IMAD reaches only ~88 lanes/clk/SM: tools/int32_pipes.cu)
// The host counts INSTRUCTIONS (1, 2, 2, 2 per chain-iteration); see tools/int32_pipes.cu for
// the full table of pipes.
template <int WHICH>
__global__ void __launch_bounds__(256) int32_peak_kernel(int iters, const int *__restrict__ src, int *sink,
                                                         int c1, int c2)
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
            } else if (WHICH == 2) {
                x[j] = __viaddmax_s32(x[j], a[j], b[j]);
                x[j] = __viaddmin_s32(x[j], b[j], a[j]);
            } else {
                asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[j]) : "r"(c2));
            }
        }
    }
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= x[j] ^ a[j] ^ b[j];
    if (acc == 0x7fffffff) sink[0] = acc;
}

}  // namespace tanw
