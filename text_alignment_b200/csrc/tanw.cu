// tanw.cu -- host side of libtanw.so: the C ABI declared in include/tanw.h.
//
// Replaces the per-page call textSeqCompare.perform_alignment (textSeqCompare.py:13-177,
// call site alignToOCR.py:273-274) by a batched device implementation.  No CPU fallback:
// every entry either runs on an sm_100 device or returns an error.
//
// A batch moves through three streams of the context:
//   s_in   host -> device copies: the four pair arrays first, then the symbols in pieces;
//   s_k    the table kernels (tanw_tables.cuh: everything the align kernels need is derived from
//          the pair arrays on the device; the host only waits for a 1 KB survey of the batch to
//          size its scratch), then the align kernels, chunk after chunk;
//   s_k2   the page kernel of a chunk, beside the chunk's line kernels on s_k (disjoint parts of the
//          pointer arena): a batch of lines always has a few pairs that are pages, and their kernel
//          -- one warp per pair, ~50 us however few they are -- fills the tail of the line kernel
//          instead of following it.  A batch of pages is cut in two chunks whose kernels alternate
//          between s_k2 and s_k3 (two pointer arenas): the second kernel's blocks move in as the
//          first one's retire, while the first chunk's results travel and the second's inputs arrive;
//   s_out  device -> host copies of a chunk's op strings / lengths / scores while the next
//          chunk is being aligned.
// tanw_align_batch cuts a batch whose copies matter (10^5 short line pairs: 40 MB of copies for
// 1 ms of alignment) into chunks so that the three streams overlap; the three-phase form
// (prepare / run / fetch) treats the batch as one chunk.
#include "tanw.h"
#include "tanw_launch.h"
#include "tanw_lines16.cuh"
#include "tanw_tables.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace tanw;

namespace {

thread_local std::string g_last_error;   // for failures that have no context yet

struct HostTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    float ms() const
    {
        return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Every entry point runs on the context's device and leaves the caller's current device as it
// found it (a torch process keeps allocating on its own device afterwards).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;            // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int kPieces = 8;                          // the symbol upload is cut into this many copies

struct ChunkPlan {
    int64_t first = 0, count = 0;                   // pairs [first, first + count)
    int64_t ops_base = 0, cap = 0;                  // its bytes of the canonical op layout
    int64_t cells = 0;
    int n_page = 0, n_line = 0, n_quads = 0;
    int n_line16 = 0, n_octets = 0;                 // pairs / warps' work units of the 16-bit line kernel
    int line_class[4] = {0, 0, 0, 0}, line16_class[4] = {0, 0, 0, 0};
    int page_shift = 0;                             // quantisation of n*m for the page order keys
    int piece = 0;                                  // symbol piece that completes the chunk's inputs
    bool tables_built = false;
};

struct LongPair {
    int p;
    int sidx;                                       // its scoring system in a multi batch
    PairDesc pd;                                    // ops_off is read from the device table
    int cfull, rows;                                // stripe strip width, rows per band
};

}  // namespace

struct tanw_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_k2 = nullptr, s_k3 = nullptr, s_l0 = nullptr, s_l1 = nullptr, s_out = nullptr;
    cudaEvent_t ev_h2d0 = nullptr, ev_h2d1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr,
                ev_d2h0 = nullptr, ev_d2h1 = nullptr, ev_tab = nullptr, ev_survey = nullptr, ev_idle = nullptr, ev_small = nullptr;
#ifdef TANW_TUNING
    cudaEvent_t tl_ls[kMaxChunks] = {}, tl_le[kMaxChunks] = {}, tl_d[kMaxChunks] = {};   // timeline of a pipelined call (TANW_TIMELINE)
#endif
    cudaEvent_t ev_piece[kPieces] = {}, ev_chunk[kMaxChunks] = {}, ev_fork[kMaxChunks] = {}, ev_pages[kMaxChunks] = {},
                ev_lines[kMaxChunks] = {};
    std::string err;
    int64_t arena_limit = 0;
    int64_t max_nm = 0;                   // largest n+m of the prepared batch (range check on rescore)
    int64_t total_mem = 0;

    DevBuf d_sym, d_n, d_m, d_toff, d_ooff, d_pairs, d_route, d_order, d_lsorted, d_hist, d_classes, d_tilesums,
           d_survey, d_counter, d_arena, d_bnd, d_ops, d_len, d_scores, d_subst, d_chain, d_ck, d_kparams, d_sidx,
           d_misc, d_pack;
    Survey *h_survey = nullptr;           // pinned: the device's report on the batch
    int *h_misc = nullptr;                // pinned: [0] device assertion word, [1] largest symbol code
    uint8_t *h_small = nullptr;           // pinned: pair descriptors + routes of a handful of pairs built by the host
    uint8_t *h_io = nullptr;              // pinned: inputs / results of a small batch travel through here (kSmallIo bytes each way)
    int *h_subst = nullptr;               // pinned copy of the substitution table in kernel encoding
    size_t h_subst_cap = 0;
    KParams *h_kparams = nullptr;         // pinned: per-pair scoring systems of a multi batch
    size_t h_kparams_cap = 0;
    int line_mode = 1;                    // 0: no line kernels, 1: both, 2: the int32 line kernel only
    bool packed_ops = false;              // fetch delivers 2-bit packed op strings (tanw_set_packed_ops)
    bool batch_packed = false;            // ... as the prepared batch was laid out
    bool alternate = false;               // page kernels of successive chunks on s_k2 / s_k3, arenas of their own
    bool alt_lines = false;               // line kernels of successive chunks on s_l0 / s_l1, arenas of their own
    int64_t line_arena_bytes = 0;         // one line arena
    int64_t page_slots = 0;               // warp slots of one page arena
    int line16_max_n = 0;                 // tallest pair of the prepared batch on the 16-bit line kernel (0: none)
    int long_capacity = 0;                // resident warps for a cooperative launch
    int long_epoch = 0;                   // stamps the chain records of a launch
    int long_band_rows = 0;               // 0 = one band unless the pointer block exceeds the arena limit
    int sym_bytes = 1;                    // 1: uint8 symbol codes; 2: uint16
    int batch_sym_bytes = 1;              // width the prepared batch was uploaded with
    int64_t long_cells = int64_t(1) << 26;   // pairs with n*m >= this use the chained-pass path
    std::vector<LongPair> longs;
    std::vector<int32_t> h_n, h_m;        // three-phase form: the lengths, for the layout check in fetch
    std::vector<uint8_t> h_stage;         // used when the caller's op layout is not canonical

    // state of the prepared batch
    bool prepared = false, ran = false;
    int64_t n_pairs = 0, ops_total = 0;
    int n_chunks = 0;
    ChunkPlan chunk[kMaxChunks];
    KParams kp;
    bool use_subst = false, multi = false;
    int page_subst = 0;                   // how the page kernel scores: 0 equality, 1 table lookups, 2 query profile
    int var = 0;                          // recurrence variant of the batch (tanw_kernels.cuh)
    BatchArgs args;
    LineArgs largs;
    TableArgs ta;                         // the table kernels' arguments of the prepared batch
    int table_launches = 0;
    int64_t slot_bytes = 0, line_slot = 0, line16_slot = 0;
    int grid = 0, line_grid = 0, line16_grid = 0;
    int occ_plain = 0, occ_subst = 0, occ_line = 0, occ_line16 = 0;
    tanw_timing timing;
};

namespace {

int fail(tanw_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_last_error = buf;
    return code;
}

#define TANW_CUDA(ctx, call)                                                              \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(ctx, TANW_E_CUDA, "%s failed: %s (%s:%d)", #call,                 \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                      \
    } while (0)

#define TANW_ENTER(ctx)                                                                   \
    DeviceGuard guard_((ctx)->device);                                                    \
    if (guard_.err != cudaSuccess)                                                        \
        return fail(ctx, TANW_E_CUDA, "cudaSetDevice(%d): %s", (ctx)->device, cudaGetErrorString(guard_.err))

bool device_is_blackwell(int device, cudaDeviceProp *prop_out)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return false;
    if (prop_out) *prop_out = prop;
    return prop.major == 10;
}

// int32 fixed point: every finite intermediate must stay far away from kNeg = -2^30.
// |value| <= (n+m+2) * max|param| ; carried << 6 and offset by up to ex*n once more: 2^22 * 2 * 64 = 2^29 < 2^30.
int64_t scoring_pmax(const tanw_scoring *s)
{
    int64_t pmax = 1;
    auto upd = [&](int64_t v) { pmax = std::max<int64_t>(pmax, v < 0 ? -v : v); };
    upd(s->match); upd(s->mismatch); upd(s->boundary_gap);
    upd((int64_t)s->gap_open_x + s->gap_extend_x); upd(s->gap_extend_x);
    upd((int64_t)s->gap_open_y + s->gap_extend_y); upd(s->gap_extend_y);
    if (s->subst)
        for (int64_t i = 0; i < (int64_t)s->subst_k * s->subst_k; ++i) upd(s->subst[i]);
    return pmax;
}
bool in_range(int64_t pmax, int64_t max_n_plus_m) { return (max_n_plus_m + 2) * pmax < (int64_t(1) << 22); }

const char *kRangeMessage =
    "scores do not fit the int32 fixed-point representation: (n+m+2)*max|param| must be < 2^22";

KParams make_kparams(const tanw_scoring *sc)
{
    KParams kp;
    memset(&kp, 0, sizeof kp);
    kp.maT = (sc->match * (1 << kShift)) | kTagM;
    kp.miT = (sc->mismatch * (1 << kShift)) | kTagM;
    kp.ox = (sc->gap_open_x + sc->gap_extend_x) * (1 << kShift);
    kp.ex = sc->gap_extend_x * (1 << kShift);
    kp.oy = (sc->gap_open_y + sc->gap_extend_y) * (1 << kShift);
    kp.ey = sc->gap_extend_y * (1 << kShift);
    kp.bg = sc->boundary_gap * (1 << kShift);
    return kp;
}

// Which recurrence variant serves a scoring system: 0 general, 1 gap opens <= 0, 2 gap opens <= 0
// and gap_extend_y == 0.
int variant_of(const tanw_scoring *sc)
{
    if (!(sc->gap_open_x <= 0 && sc->gap_open_y <= 0)) return 0;
    return sc->gap_extend_y == 0 ? 2 : 1;
}

// Tallest line pair the 16-bit line kernel may take under a scoring system (tanw_lines16.cuh): the
// D-only recurrences (gap opens <= 0), an equality scorer with match >= mismatch, and every value a
// half can hold -- (2n + m + 4) * max|param| with m <= 128 -- within kRange16.  0: not eligible.
int line16_limit(const tanw_scoring *sc)
{
    if (sc->subst || variant_of(sc) < 1 || sc->match < sc->mismatch) return 0;
    const int64_t n = (kRange16 / scoring_pmax(sc) - (kLineMaxM + 4)) / 2;
    return (int)std::max<int64_t>(0, std::min<int64_t>(n, kLineMaxN));
}

// The table of a tabulated scorer in kernel encoding, uploaded from a pinned copy the context
// keeps (so that nothing waits for the copy).
int upload_subst(tanw_ctx *ctx, const tanw_scoring *sc, int64_t *h2d)
{
    const size_t kk = (size_t)sc->subst_k * (size_t)sc->subst_k;
    if (kk > ctx->h_subst_cap) {
        if (ctx->h_subst) cudaFreeHost(ctx->h_subst);
        ctx->h_subst = nullptr;
        ctx->h_subst_cap = 0;
        if (cudaMallocHost((void **)&ctx->h_subst, sizeof(int) * kk) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, TANW_E_NOMEM, "out of pinned host memory for the substitution table");
        }
        ctx->h_subst_cap = kk;
    }
    // the previous table may still be on its way to the device
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
    for (size_t i = 0; i < kk; ++i) ctx->h_subst[i] = (sc->subst[i] * (1 << kShift)) | kTagM;
    if (ctx->d_subst.reserve(sizeof(int) * kk) != cudaSuccess)
        return fail(ctx, TANW_E_NOMEM, "device allocation failed (substitution table)");
    TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_idle, 0));
    TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_subst.p, ctx->h_subst, sizeof(int) * kk, cudaMemcpyHostToDevice, ctx->s_in));
    ctx->kp.subst = (const int *)ctx->d_subst.p;
    ctx->kp.subst_k = sc->subst_k;
    if (h2d) *h2d += (int64_t)(sizeof(int) * kk);
    return TANW_OK;
}

// The chain records carry an epoch stamp; fresh memory must not contain a live one.
cudaError_t reserve_zeroed(tanw_ctx *ctx, DevBuf &buf, size_t bytes)
{
    if (bytes <= buf.cap) return cudaSuccess;
    cudaError_t e = buf.reserve(bytes);
    if (e == cudaSuccess && buf.cap) {
        e = cudaMemsetAsync(buf.p, 0, buf.cap, ctx->s_k);
        ctx->long_epoch = 0;
    }
    return e;
}

// Stripe width of a chained-pass pair.  A stripe is one warp that is bound by its own
// instruction latency, so more, narrower stripes mean more warps per SM sub-partition; but every
// stripe also adds ~46 steps of pipeline fill (31 rows of lane skew + the boundary look-ahead).
// Measured on config 5 (100k columns): C = 4 / 8 / 12 / 16 -> 37.0 / 35.5 / 37.3 / 39.9 ms.
int long_stripe_c(const tanw_ctx *ctx, int m)
{
#ifdef TANW_TUNING
    if (const char *e = getenv("TANW_LONG_C")) {          // tuning builds only
        const int c = atoi(e);
        if (c >= 4 && c <= kMaxC && c % 4 == 0 && (m + 32 * c - 1) / (32 * c) <= ctx->long_capacity) return c;
    }
#endif
    if ((m + 255) / 256 >= 2 * ctx->sm_count && (m + 255) / 256 <= ctx->long_capacity) return 8;
    for (int c = 4; c < kMaxC; c += 4)
        if ((m + 32 * c - 1) / (32 * c) <= ctx->long_capacity) return c;
    return kMaxC;
}

// Rows per band of a chained-pass pair: the whole pair when its pointer block fits `limit`
// bytes (and no band height is forced), otherwise the tallest band that does.  0 = not even
// kMinBandRows rows fit.
constexpr int kMaxWideSubstK = 2048;                // substitution table side with 16-bit symbols (16 MB)
constexpr int kMinBandRows = 32;
int long_band_rows(const tanw_ctx *ctx, int n, int m, int cf, int64_t limit)
{
    const int64_t row_bytes = (int64_t)((m + 32 * cf - 1) / (32 * cf)) * 32 * cf;   // ptr_bytes = row_bytes*(rows+32)
    int64_t rows = limit / std::max<int64_t>(row_bytes, 1) - 32;
    if (ctx->long_band_rows > 0) rows = std::min<int64_t>(rows, ctx->long_band_rows);
    if (rows >= n) return std::max(n, 1);
    return rows >= kMinBandRows || (ctx->long_band_rows > 0 && rows >= ctx->long_band_rows) ? (int)rows : 0;
}

// One whole-manuscript pair on the chained-pass path: one cooperative launch per wave of
// resident stripes, then the traceback.  When the pair is cut into row bands (its pointer block
// would not fit the arena) the fill runs top to bottom once without storing pointers, leaving
// the per-column state (X, D, W) at every band edge; then, bottom to top, each band is filled
// again from its checkpoint with pointers stored and the traceback continues through it.
int run_long_pair(tanw_ctx *ctx, const LongPair &lp, const KParams &kp_pair, int var, int *launches)
{
    const PairDesc &pd = lp.pd;
    const int sb = ctx->batch_sym_bytes;
    const int R = lp.rows;
    const int B = (pd.n + R - 1) / R;
    int *state = (int *)ctx->d_ck.p;
    int *ck = state + 4;
    LongArgs la;
    la.T = (const uint8_t *)ctx->d_sym.p + pd.t_off * sb;
    la.O = (const uint8_t *)ctx->d_sym.p + pd.o_off * sb;
    la.n = pd.n;
    la.m = pd.m;
    la.ptr = (uint8_t *)ctx->d_arena.p;
    la.chain = (int4 *)ctx->d_chain.p;
    la.chain_stride = (long long)std::min(R, pd.n) + 4;
    la.cfull = lp.cfull;
    la.scores = (int *)ctx->d_scores.p + 3 * (size_t)lp.p;
    la.check = (int *)ctx->d_misc.p;
    const int npass = (pd.m + 32 * la.cfull - 1) / (32 * la.cfull);
    KParams kp = kp_pair;
    const void *fn = long_kernel(var, ctx->use_subst, sb);
    auto fill_band = [&](int b, bool store) -> int {
        la.r0 = b * R;
        la.nb = std::min(R, pd.n - la.r0);
        la.ck_in = b > 0 ? ck + (size_t)(b - 1) * 3 * (size_t)pd.m : nullptr;
        la.ck_out = (!store && b < B - 1) ? ck + (size_t)b * 3 * (size_t)pd.m : nullptr;
        la.store = store ? 1 : 0;
        la.epoch = ++ctx->long_epoch;
        if (la.epoch == 0) la.epoch = ++ctx->long_epoch;
        TANW_CUDA(ctx, launch_long_col0(la.chain + (size_t)npass * (size_t)la.chain_stride, la.nb, la.r0, kp.bg,
                                        la.epoch, ctx->s_k));
        ++*launches;
        for (int w0 = 0; w0 < npass; w0 += ctx->long_capacity) {
            la.pass0 = w0;
            const int grid = std::min(ctx->long_capacity, npass - w0);
            void *args[] = { (void *)&la, (void *)&kp };
            TANW_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(32), args, 0, ctx->s_k));
            ++*launches;
        }
        return TANW_OK;
    };
    for (int b = 0; b < B - 1; ++b)
        if (int rc = fill_band(b, false)) return rc;
#ifdef TANW_TUNING
    static const bool skip_trace = getenv("TANW_DEBUG_SKIP_TRACE") != nullptr;   // tuning builds only
#else
    const bool skip_trace = false;
#endif
    for (int b = B - 1; b >= 0; --b) {
        if (int rc = fill_band(b, true)) return rc;
        if (skip_trace) continue;
        TANW_CUDA(ctx, launch_long_trace(la.ptr, (const PairDesc *)ctx->d_pairs.p + lp.p, la.cfull, la.r0, la.nb,
                                         b == B - 1, b == 0, state, (uint8_t *)ctx->d_ops.p,
                                         (int *)ctx->d_len.p + lp.p, ctx->s_k));
        ++*launches;
    }
    return TANW_OK;
}

// Pair descriptors, routes and work orders of one chunk (tanw_tables.cuh), on the compute stream.
int build_chunk_tables(tanw_ctx *ctx, int c)
{
    ChunkPlan &cp = ctx->chunk[c];
    if (cp.tables_built || cp.count <= 0) return TANW_OK;
    const TableArgs &ta = ctx->ta;
    TANW_CUDA(ctx, cudaMemsetAsync((int *)ctx->d_hist.p + (size_t)c * kHistStride, 0, sizeof(int) * kHistStride, ctx->s_k));
    const unsigned tiles = (unsigned)((cp.count + kTile - 1) / kTile);
    build_kernel<<<tiles, kTileThreads, 0, ctx->s_k>>>(ta, c, cp.page_shift);
    bins_kernel<<<1, 1024, 0, ctx->s_k>>>(ta, c, make_int4(cp.line16_class[0], cp.line16_class[1], cp.line16_class[2], cp.line16_class[3]),
                                          make_int4(cp.line_class[0], cp.line_class[1], cp.line_class[2], cp.line_class[3]));
    scatter_kernel<<<(unsigned)((cp.count + kTileThreads - 1) / kTileThreads), kTileThreads, 0, ctx->s_k>>>(ta, c, cp.page_shift);
    TANW_CUDA(ctx, cudaGetLastError());
    ctx->table_launches += 3;
    ctx->timing.table_launches = ctx->table_launches;
    cp.tables_built = true;
    return TANW_OK;
}

// 2-bit packing of the op strings of a chunk's pairs with the given routes (tanw_tables.cuh).
int pack_chunk(tanw_ctx *ctx, const ChunkPlan &cp, unsigned routes, cudaStream_t stream)
{
    if (cp.count <= 0) return TANW_OK;
    const int blocks = (int)std::min<int64_t>((cp.count + 7) / 8, (int64_t)ctx->sm_count * 8);
    pack_ops_kernel<<<blocks, 256, 0, stream>>>((const PairDesc *)ctx->d_pairs.p, (const unsigned char *)ctx->d_route.p,
                                                (const int *)ctx->d_len.p, (const uint8_t *)ctx->d_ops.p,
                                                (uint8_t *)ctx->d_pack.p, cp.first, cp.count, routes);
    TANW_CUDA(ctx, cudaGetLastError());
    return TANW_OK;
}

// ops_off is canonical when pair p's bytes start where pair p-1's capacity (n+m) ends.
bool layout_is_canonical(const int64_t *ops_off, const int32_t *n, const int32_t *m, int64_t P)
{
    if (P == 0) return true;
    int64_t diff = ops_off[0];
    for (int64_t p = 0; p + 1 < P; ++p) diff |= (ops_off[p + 1] - ops_off[p]) ^ ((int64_t)n[p] + m[p]);
    return diff == 0;
}

struct PrepareInput {
    const uint8_t *symbols;
    int64_t symbols_len;
    const int64_t *t_off, *o_off;
    const int32_t *n, *m;
    int64_t n_pairs;
    const tanw_scoring *sc;               // one system, or n_sc systems with per-pair indices
    int32_t n_sc;
    const int32_t *sidx;
    bool pipelined;                       // cut into chunks when that pays (tanw_align_batch)
};

// The survey of a handful of pairs, by the host from the arrays it was given: the per-pair rules of
// survey_kernel (tanw_tables.cuh) word for word.  A single page per call is the reference's own
// pattern (alignToOCR.py:273); waiting for the device's report and then for its list of
// chained-stripe pairs costs two round trips of ~30 us each, a sixth of such a call.
constexpr int64_t kHostSurveyPairs = 64;
// A small batch -- one page per call is the reference's own pattern -- is copied through page-locked
// staging of the context: a caller's ordinary (pageable) memory makes every cudaMemcpyAsync a
// synchronous, separately staged transfer, ~10 us apiece for a dozen copies of a few bytes.
constexpr int64_t kSmallIo = 128 << 10;
void host_survey(const TableArgs &a, const PrepareInput &in, Survey &sv)
{
    memset(&sv, 0, offsetof(Survey, long_list));
    ChunkSurvey &cs = sv.chunk[0];
    int64_t bad = -1;
    for (int64_t p = 0; p < in.n_pairs; ++p) {
        const long long np = in.n[p], mp = in.m[p], to = in.t_off[p], oo = in.o_off[p];
        if (np < 0 || mp < 0 || to < 0 || oo < 0 || to + np > a.symbols_len || oo + mp > a.symbols_len) {
            if (bad < 0) bad = p;
            continue;
        }
        cs.cap += np + mp;
        cs.cells += np * mp;
        cs.sym_end = std::max<long long>(cs.sym_end, std::max(to + np, oo + mp));
        sv.max_nm = std::max(sv.max_nm, (int)std::min<long long>(np + mp, 0x7fffffffll));
        const int cls = (np > 0 && mp > 0) ? line_c((int)mp) / 4 - 1 : 0;
        switch (route_of(a, np, mp)) {
        case kRouteLong:
            ++cs.n_long;
            if (sv.n_long < kMaxLongList) sv.long_list[sv.n_long] = (int)p;
            ++sv.n_long;
            break;
        case kRouteLine:
            ++cs.n_line;
            cs.max_line_slot = std::max(cs.max_line_slot, line_ptr_bytes((int)np, (int)mp));
            ++cs.line_class[cls];
            break;
        case kRouteLine16:
            ++cs.n_line16;
            cs.max_nm_line16 = std::max(cs.max_nm_line16, (int)(np + mp));
            cs.max_n_line16 = std::max(cs.max_n_line16, (int)np);
            cs.max_line16_slot = std::max(cs.max_line16_slot, line_ptr_bytes((int)np, (int)mp));
            ++cs.line16_class[cls];
            break;
        default:
            ++cs.n_page;
            cs.page_cells += np * mp;
            cs.max_slot = std::max(cs.max_slot, ptr_bytes((int)np, (int)mp));
            cs.max_page_cells = std::max(cs.max_page_cells, np * mp);
            cs.max_n_page = std::max(cs.max_n_page, (int)np);
            break;
        }
    }
    sv.bad = bad >= 0 ? 0xFFFFFFFFFFFFFFFFull - (unsigned long long)bad : 0;
}

int prepare_impl(tanw_ctx *ctx, const PrepareInput &in)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    HostTimer host_timer;
    ctx->prepared = false;
    ctx->ran = false;
    const int64_t P = in.n_pairs;
    if (P < 0 || in.symbols_len < 0) return fail(ctx, TANW_E_INVALID, "negative size");
    if (P > 0 && (!in.t_off || !in.n || !in.o_off || !in.m)) return fail(ctx, TANW_E_INVALID, "NULL pair table");
    if (in.symbols_len > 0 && !in.symbols) return fail(ctx, TANW_E_INVALID, "symbols is NULL");
    if (!in.sc) return fail(ctx, TANW_E_INVALID, "scoring is NULL");
    if (P > 0x7fffffff) return fail(ctx, TANW_E_INVALID, "more than 2^31-1 pairs in one batch");
    const bool multi = in.sidx != nullptr;
    if (multi && in.n_sc < 1) return fail(ctx, TANW_E_INVALID, "a multi batch needs at least one scoring system");
    const int sb = ctx->sym_bytes;
    const int max_k = sb == 1 ? 256 : kMaxWideSubstK;
    const tanw_scoring *sc = in.sc;
    if (multi) {
        if (sb != 1) return fail(ctx, TANW_E_INVALID, "per-pair scoring systems need 8-bit symbol codes");
        for (int32_t i = 0; i < in.n_sc; ++i)
            if (sc[i].subst) return fail(ctx, TANW_E_INVALID, "per-pair scoring systems must be equality scorers (no table)");
    } else if (sc->subst && (sc->subst_k < 1 || sc->subst_k > max_k)) {
        return fail(ctx, TANW_E_INVALID, "subst_k must be in 1..%d", max_k);
    }
    TANW_ENTER(ctx);

    // ---- uploads: pair arrays first (the survey needs only them; below), then the symbols in pieces ----
    const size_t Pz = (size_t)std::max<int64_t>(P, 1);
    const int64_t n_tiles = (P + kTile - 1) / kTile;
    if (ctx->d_n.reserve(sizeof(int) * Pz) != cudaSuccess || ctx->d_m.reserve(sizeof(int) * Pz) != cudaSuccess ||
        ctx->d_toff.reserve(sizeof(int64_t) * Pz) != cudaSuccess || ctx->d_ooff.reserve(sizeof(int64_t) * Pz) != cudaSuccess ||
        ctx->d_sym.reserve((size_t)in.symbols_len * (size_t)sb + 16) != cudaSuccess ||
        ctx->d_tilesums.reserve(sizeof(int64_t) * (size_t)std::max<int64_t>(n_tiles, 1)) != cudaSuccess ||
        ctx->d_survey.reserve(sizeof(Survey)) != cudaSuccess || ctx->d_misc.reserve(256) != cudaSuccess ||
        (multi && ctx->d_sidx.reserve(sizeof(int) * Pz) != cudaSuccess)) {
        cudaGetLastError();
        return fail(ctx, TANW_E_NOMEM, "device allocation failed (inputs, %lld symbol bytes)", (long long)in.symbols_len * sb);
    }
    // the previous batch of this context may still be running on s_k
    TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_idle, 0));
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_h2d0, ctx->s_in));
    int64_t h2d = 0;
    const int64_t sym_bytes_total = in.symbols_len * sb;
    const int64_t piece_bytes = std::max<int64_t>((sym_bytes_total + kPieces - 1) / kPieces / 256 * 256 + 256, 1 << 16);

    // ---- the survey: sizes, routes, validation, on the device; the host waits for 1 KB ----------
    int64_t limit = ctx->arena_limit;
    if (limit == 0) limit = ctx->total_mem / 10 * 4;     // no cudaMemGetInfo on the per-batch path
    TableArgs ta;
    memset(&ta, 0, sizeof ta);
    ta.t_off = (const long long *)ctx->d_toff.p;
    ta.o_off = (const long long *)ctx->d_ooff.p;
    ta.n = (const int *)ctx->d_n.p;
    ta.m = (const int *)ctx->d_m.p;
    ta.n_pairs = P;
    ta.symbols_len = in.symbols_len;
    // the survey reports on up to kMaxSlices slices of the batch; the host merges them into chunks
    const int64_t slice_pairs = std::max<int64_t>(((P + kMaxSlices - 1) / kMaxSlices + kTile - 1) / kTile * kTile, kTile);
    ta.chunk_pairs = slice_pairs;
    ta.long_cells = ctx->long_cells;
    ta.slot_limit = limit;
    ta.use_lines = (ctx->line_mode != 0 && !multi) ? 1 : 0;
    ta.line16_max_n = (ctx->line_mode == 1 && !multi && sb == 1) ? line16_limit(sc) : 0;
    ta.wide = sb == 2 ? 1 : 0;
    ta.tiny_batch = P <= 2 ? 1 : 0;
    ta.can_long = ctx->long_capacity > 0 ? 1 : 0;
    ta.survey = (Survey *)ctx->d_survey.p;
    ta.tile_sums = (long long *)ctx->d_tilesums.p;
    Survey &sv = *ctx->h_survey;
    const size_t survey_head = offsetof(Survey, long_list);
    const bool host_sv = P <= kHostSurveyPairs;
    if (host_sv) host_survey(ta, in, sv);
    // ... and when every one of them takes the chained-stripe path (a single page per call), the host
    // writes their descriptors too: no survey and no table kernels at all
    const bool host_tables = host_sv && P > 0 && sv.bad == 0 && sv.n_long == P && !ctx->packed_ops;
    // small inputs go through the context's page-locked staging (its previous use has left: the
    // stream has passed ev_h2d1 of the previous batch)
    const uint8_t *sym_src = in.symbols;
    const int32_t *n_src = in.n, *m_src = in.m;
    const int64_t *t_src = in.t_off, *o_src = in.o_off;
    if (host_sv && sym_bytes_total + 24 * P <= kSmallIo) {
        TANW_CUDA(ctx, cudaEventSynchronize(ctx->ev_h2d1));
        uint8_t *at = ctx->h_io;
        auto put = [&at](const void *src, size_t bytes) { uint8_t *p = at; if (bytes) memcpy(p, src, bytes); at += (bytes + 15) / 16 * 16; return p; };
        if (P > 0 && !host_tables) {
            t_src = (const int64_t *)put(in.t_off, sizeof(int64_t) * (size_t)P);
            o_src = (const int64_t *)put(in.o_off, sizeof(int64_t) * (size_t)P);
            n_src = (const int32_t *)put(in.n, sizeof(int32_t) * (size_t)P);
            m_src = (const int32_t *)put(in.m, sizeof(int32_t) * (size_t)P);
        }
        sym_src = put(in.symbols, (size_t)sym_bytes_total);
    }
    if (P > 0 && !host_tables) {                         // (an all-chained handful of pairs needs none of them on the device)
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_n.p, n_src, sizeof(int) * (size_t)P, cudaMemcpyHostToDevice, ctx->s_in));
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_m.p, m_src, sizeof(int) * (size_t)P, cudaMemcpyHostToDevice, ctx->s_in));
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_toff.p, t_src, sizeof(int64_t) * (size_t)P, cudaMemcpyHostToDevice, ctx->s_in));
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_ooff.p, o_src, sizeof(int64_t) * (size_t)P, cudaMemcpyHostToDevice, ctx->s_in));
        h2d += 24 * P;
        if (multi) {
            TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_sidx.p, in.sidx, sizeof(int) * (size_t)P, cudaMemcpyHostToDevice, ctx->s_in));
            h2d += 4 * P;
        }
    }
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_tab, ctx->s_in));
    TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_k, ctx->ev_tab, 0));
    if (!host_tables) {
        TANW_CUDA(ctx, cudaMemsetAsync(ctx->d_survey.p, 0, survey_head, ctx->s_k));
        if (n_tiles > 0) {                                // (also beside a host survey: the table kernels use its tile sums)
            survey_kernel<<<(unsigned)n_tiles, kTileThreads, 0, ctx->s_k>>>(ta);
            TANW_CUDA(ctx, cudaGetLastError());
        }
    }
    if (!host_sv) {
        TANW_CUDA(ctx, cudaMemcpyAsync(&sv, ctx->d_survey.p, survey_head, cudaMemcpyDeviceToHost, ctx->s_k));
        TANW_CUDA(ctx, cudaEventRecord(ctx->ev_survey, ctx->s_k));
    }
    // host work that does not need the survey runs while it is on its way: first of all the symbol
    // upload, in pieces with an event each, so that a chunk's kernels wait only for their own symbols
    for (int i = 0; i < kPieces; ++i) {
        const int64_t lo = std::min(sym_bytes_total, piece_bytes * i), hi = std::min(sym_bytes_total, piece_bytes * (i + 1));
        if (hi <= lo && i > 0) break;                   // a chunk waits for the piece its last symbol is in: never an empty one
        if (hi > lo)
            TANW_CUDA(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_sym.p + lo, sym_src + lo, (size_t)(hi - lo),
                                           cudaMemcpyHostToDevice, ctx->s_in));
        TANW_CUDA(ctx, cudaEventRecord(ctx->ev_piece[i], ctx->s_in));
    }
    h2d += sym_bytes_total;
    int64_t pmax = 1;
    int var = 2;
    for (int32_t i = 0; i < (multi ? in.n_sc : 1); ++i) {
        pmax = std::max(pmax, scoring_pmax(&sc[i]));
        var = std::min(var, variant_of(&sc[i]));
    }
    if (sb == 2) var = 0;
    if (!in.pipelined) {
        try {
            ctx->h_n.assign(in.n, in.n + P);
            ctx->h_m.assign(in.m, in.m + P);
        } catch (const std::bad_alloc &) {
            return fail(ctx, TANW_E_NOMEM, "out of host memory");
        }
    }
    if (!host_sv) TANW_CUDA(ctx, cudaEventSynchronize(ctx->ev_survey));
    else if (!in.pipelined && !host_tables) TANW_CUDA(ctx, cudaEventSynchronize(ctx->ev_tab));   // tanw.h: the pair arrays are consumed before prepare returns

    if (sv.bad != 0) {
        const int64_t p = (int64_t)(0xFFFFFFFFFFFFFFFFull - sv.bad);
        if (p >= 0 && p < P && (in.n[p] < 0 || in.m[p] < 0))
            return fail(ctx, TANW_E_INVALID, "pair %lld: negative length", (long long)p);
        return fail(ctx, TANW_E_INVALID, "pair %lld: offsets outside the symbol buffer", (long long)p);
    }
    if (!in_range(pmax, sv.max_nm)) return fail(ctx, TANW_E_RANGE, "%s", kRangeMessage);
    if (multi)
        for (int64_t p = 0; p < P; ++p)
            if (in.sidx[p] < 0 || in.sidx[p] >= in.n_sc)
                return fail(ctx, TANW_E_INVALID, "pair %lld: scoring index %d outside 0..%d", (long long)p, in.sidx[p], in.n_sc - 1);
    if (sv.n_long > kMaxLongList)
        return fail(ctx, TANW_E_INVALID, "%d pairs of this batch need the chained-stripe path; at most %d per batch",
                    sv.n_long, kMaxLongList);

    // ---- chained-stripe pairs: few, handled by the host one by one ---------------------------------
    ctx->longs.clear();
    int64_t max_long = 0, max_long_bnd = 0, max_ck = 0;
    if (sv.n_long > 0) {
        if (!host_sv) {
            TANW_CUDA(ctx, cudaMemcpyAsync(sv.long_list, (const uint8_t *)ctx->d_survey.p + survey_head,
                                           sizeof(int) * (size_t)sv.n_long, cudaMemcpyDeviceToHost, ctx->s_k));
            TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_k));
        }
        std::sort(sv.long_list, sv.long_list + sv.n_long);
        for (int i = 0; i < sv.n_long; ++i) {
            LongPair lp;
            lp.p = sv.long_list[i];
            lp.sidx = multi ? in.sidx[lp.p] : 0;
            const int64_t np = in.n[lp.p], mp = in.m[lp.p];
            lp.pd.t_off = in.t_off[lp.p]; lp.pd.o_off = in.o_off[lp.p]; lp.pd.ops_off = 0;
            lp.pd.n = (int)np; lp.pd.m = (int)mp;
            lp.cfull = long_stripe_c(ctx, (int)mp);
            const int64_t npass = (mp + 32 * lp.cfull - 1) / (32 * lp.cfull);
            lp.rows = long_band_rows(ctx, (int)np, (int)mp, lp.cfull, limit);
            if (lp.rows <= 0)
                return fail(ctx, TANW_E_NOMEM, "pair %lld: not even %d rows of traceback pointers (%lld bytes each) "
                            "fit the arena limit of %lld bytes", (long long)lp.p, kMinBandRows,
                            (long long)(npass * 32 * lp.cfull), (long long)limit);
            const int64_t bands = (np + lp.rows - 1) / lp.rows;
            max_long = std::max<int64_t>(max_long, ptr_bytes((int)std::min<int64_t>(lp.rows, np), (int)mp, lp.cfull));
            max_long_bnd = std::max(max_long_bnd, (npass + 1) * (std::min<int64_t>(lp.rows, np) + 4));
            max_ck = std::max(max_ck, (bands - 1) * 3 * mp);
            ctx->longs.push_back(lp);
        }
    }

    // ---- chunks: merge the survey's slices ----------------------------------------------------------
    const int n_slices = (int)std::max<int64_t>((P + slice_pairs - 1) / slice_pairs, 1);
    int64_t cells = 0, page_cells = 0, cap_total = 0, max_slot = 0, max_line_slot = 0, max_line16_slot = 0;
    int max_n = 0, n_page_total = 0, n_line_total = 0, max_nm16 = 0;
    for (int s = 0; s < n_slices; ++s) {
        const ChunkSurvey &cs = sv.chunk[s];
        cells += cs.cells; page_cells += cs.page_cells; cap_total += cs.cap;
        max_slot = std::max<int64_t>(max_slot, cs.max_slot);
        max_line_slot = std::max<int64_t>(max_line_slot, cs.max_line_slot);
        max_line16_slot = std::max<int64_t>(max_line16_slot, cs.max_line16_slot);
        max_n = std::max(max_n, cs.max_n_page);
        max_nm16 = std::max(max_nm16, cs.max_n_line16);
        n_page_total += cs.n_page; n_line_total += cs.n_line;
    }
    int S = 1;
    int per = n_slices;                                  // slices per chunk
    bool alternate = false;                              // page kernels of successive chunks on two streams
    bool alt_lines = false;                              // ... and the line kernels
    if (in.pipelined && n_slices > 1 && ctx->longs.empty() && !sc->subst) {
        const double copy_ms = (double)(sym_bytes_total + cap_total) / 45e6;
        const int64_t page_warps = (int64_t)ctx->sm_count * ctx->occ_plain * kWarpsPerBlock;
        if (n_page_total == 0 || page_cells * 20 < cells) {
            // Lines (with the odd page among them): chunks pay when the copies are long next to what
            // a chunk boundary costs (a few small launches), and a chunk should be whole rounds of
            // the line kernel's resident warps (eight pairs each).
            double want = std::sqrt(copy_ms / 0.012);           // (successive chunks' line kernels overlap: a boundary is cheap)
#ifdef TANW_TUNING
            if (const char *e = getenv("TANW_LINE_CHUNKS")) want = atof(e);                  // tuning builds only
#endif
            while (S * 2 <= std::min<int>(n_slices, kMaxChunks) && S * 2 <= want) S *= 2;
            if (S > 1) {
                const double per_round = (double)ctx->sm_count * ctx->occ_line16 * kL16Warps * 8 / (double)slice_pairs;
                const double rounds = std::max(1.0, std::floor((double)n_slices / S / per_round + 0.5));
                per = (int)std::max(1.0, std::floor(rounds * per_round));
                if ((n_slices + per - 1) / per > kMaxChunks) per = (n_slices + kMaxChunks - 1) / kMaxChunks;
                alt_lines = true;
            }
        } else if (n_line_total == 0 && n_page_total >= 3 * page_warps && copy_ms > 0.02 * (double)cells / 1.6e9) {
            // Pages: two chunks whose kernels run on two streams with pointer arenas of their own,
            // so that the second kernel's blocks move in as the first one's retire (no idle tail
            // between them) while the first chunk's results travel and the second's inputs arrive.
            S = 2;
            per = (n_slices + 1) / 2;
            alternate = true;
        }
        if (S == 1) per = n_slices;
    }
    S = (n_slices + per - 1) / per;
    ctx->n_chunks = S;
    int64_t ops_base = 0;
    int max_quads = 0, max_octets = 0, max_page = 0, n_line16_total = 0;
    for (int c = 0; c < S; ++c) {
        ChunkPlan &cp = ctx->chunk[c];
        cp = ChunkPlan();
        cp.first = (int64_t)c * per * slice_pairs;
        cp.count = std::min<int64_t>(P, cp.first + (int64_t)per * slice_pairs) - cp.first;
        cp.ops_base = ops_base;
        int64_t max_pc = 0, sym_end = 0;
        for (int s = c * per; s < std::min(n_slices, (c + 1) * per); ++s) {
            const ChunkSurvey &cs = sv.chunk[s];
            cp.cap += cs.cap; cp.cells += cs.cells;
            cp.n_page += cs.n_page; cp.n_line += cs.n_line; cp.n_line16 += cs.n_line16;
            for (int k = 0; k < 4; ++k) { cp.line_class[k] += cs.line_class[k]; cp.line16_class[k] += cs.line16_class[k]; }
            max_pc = std::max<int64_t>(max_pc, cs.max_page_cells);
            sym_end = std::max<int64_t>(sym_end, cs.sym_end);
        }
        for (int k = 0; k < 4; ++k) { cp.n_quads += (cp.line_class[k] + 3) / 4; cp.n_octets += (cp.line16_class[k] + 7) / 8; }
        while ((max_pc >> cp.page_shift) >= kPageKeys) ++cp.page_shift;
        cp.piece = (int)std::min<int64_t>(kPieces - 1, std::max<int64_t>(sym_end * sb - 1, 0) / piece_bytes);
        ops_base += cp.cap;
        max_quads = std::max(max_quads, cp.n_quads);
        max_octets = std::max(max_octets, cp.n_octets);
        n_line16_total += cp.n_line16;
        max_page = std::max(max_page, cp.n_page);
    }
    const int64_t chunk_pairs = (int64_t)per * slice_pairs;

    // ---- kernel parameters ------------------------------------------------------------------
    ctx->kp = make_kparams(sc);
    ctx->use_subst = !multi && sc->subst != nullptr;
    ctx->page_subst = ctx->use_subst ? 1 : 0;
    if (ctx->use_subst && sb == 1 && sc->subst_k <= kProfileMaxK) {
        bool small = true;                                  // the profile holds signed bytes
        for (int64_t i = 0; i < (int64_t)sc->subst_k * sc->subst_k && small; ++i) small = sc->subst[i] >= -127 && sc->subst[i] <= 127;
        if (small) ctx->page_subst = 2;
    }
    ctx->multi = multi;
    ctx->var = var;
    ctx->max_nm = sv.max_nm;
    ctx->batch_sym_bytes = sb;
    ctx->batch_packed = ctx->packed_ops;

    // ---- launch geometry and scratch ----------------------------------------------------------
    int occ = ctx->page_subst == 2 ? std::max(1, pairs_blocks_per_sm_profile(sc->subst_k))
            : ctx->use_subst ? ctx->occ_subst : ctx->occ_plain;
#ifdef TANW_TUNING
    if (const char *e = getenv("TANW_PAGE_BLOCKS")) occ = std::max(1, std::min(occ, atoi(e)));   // tuning builds only
#endif
    // Few pairs per resident warp: the makespan is the largest pair on one warp, and a warp that
    // shares its scheduler with three others runs at a quarter of the scheduler's rate.  Fewer CTAs
    // per SM make every warp faster while the batch still fills them (config 4, 4 096 pairs:
    // 4 / 3 / 2 CTAs per SM -> 6.03 / 5.82 / 5.73 ms; config 2, 10 000 pairs: 12.89 / 13.06 ms).
    const double pages_in_flight = alternate ? (double)n_page_total : (double)max_page;      // alternating chunks share the GPU
    while (occ > 2 && pages_in_flight / ((double)ctx->sm_count * occ * kWarpsPerBlock) < 2.5) --occ;
    int grid = ctx->sm_count * occ;
    const int64_t need_blocks = ((int64_t)max_page + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (need_blocks < grid) grid = (int)std::max<int64_t>(need_blocks, 1);
    const int64_t slot_bytes = (max_slot + 255) / 256 * 256;
    if (slot_bytes > 0) {
        int64_t max_blocks = limit / (slot_bytes * kWarpsPerBlock);
        if (max_blocks < 1)
            return fail(ctx, TANW_E_NOMEM,
                        "a pair needs %lld bytes of traceback pointers per warp; arena limit is %lld "
                        "and this device cannot run the chained-pass path", (long long)slot_bytes, (long long)limit);
        if (max_blocks < grid) grid = (int)max_blocks;
    }
    ctx->grid = grid;
    ctx->slot_bytes = slot_bytes;
    const int64_t slots = (int64_t)grid * kWarpsPerBlock;
    if (alternate && 2 * slots * slot_bytes > limit) alternate = false;       // one arena: the chunks' page kernels share a stream
    ctx->alternate = alternate;
    const int arenas = alternate ? 2 : 1;
    if (S == 1) alt_lines = false;
    int line_grid = ctx->sm_count * ctx->occ_line;
    if (((int64_t)max_quads + kWarpsPerBlock - 1) / kWarpsPerBlock < line_grid)
        line_grid = (int)std::max<int64_t>(((int64_t)max_quads + kWarpsPerBlock - 1) / kWarpsPerBlock, 1);
    const int64_t line_slot = (max_line_slot + 255) / 256 * 256;
    ctx->line_grid = line_grid;
    ctx->line_slot = line_slot;
    int64_t line_arena = max_quads ? (int64_t)line_grid * kWarpsPerBlock * 4 * line_slot : 0;
    int occ16 = ctx->occ_line16;
#ifdef TANW_TUNING
    if (const char *e = getenv("TANW_LINE16_BLOCKS")) occ16 = std::max(1, std::min(occ16, atoi(e)));   // tuning builds only
#endif
    int line16_grid = ctx->sm_count * occ16;
    if (((int64_t)max_octets + kL16Warps - 1) / kL16Warps < line16_grid)
        line16_grid = (int)std::max<int64_t>(((int64_t)max_octets + kL16Warps - 1) / kL16Warps, 1);
    const int64_t line16_slot = (max_line16_slot + 255) / 256 * 256;
    ctx->line16_grid = line16_grid;
    ctx->line16_slot = line16_slot;
    ctx->line16_max_n = n_line16_total > 0 ? std::max(max_nm16, 1) : 0;     // the tallest pair routed to the 16-bit kernel
    if (max_octets) line_arena = std::max(line_arena, (int64_t)line16_grid * kL16Warps * 8 * line16_slot);
    line_arena = (line_arena + 255) / 256 * 256;
    if (alt_lines && arenas * slots * slot_bytes + 2 * line_arena > limit) alt_lines = false;
    const int bnd_rows = max_n + 4;      // bnd[1..n] plus the prefetch overrun

    if (ctx->d_pairs.reserve(sizeof(PairDesc) * Pz) != cudaSuccess ||
        ctx->d_route.reserve(Pz) != cudaSuccess ||
        ctx->d_order.reserve(sizeof(int) * Pz) != cudaSuccess ||
        ctx->d_lsorted.reserve(sizeof(int) * Pz) != cudaSuccess ||
        ctx->d_hist.reserve(sizeof(int) * (size_t)kHistStride * (size_t)S) != cudaSuccess ||
        ctx->d_classes.reserve(sizeof(LineClasses) * 2 * kMaxChunks) != cudaSuccess ||
        ctx->d_counter.reserve(sizeof(unsigned) * 4 * kMaxChunks) != cudaSuccess ||
        ctx->d_arena.reserve((size_t)std::max<int64_t>(std::max(arenas * slots * slot_bytes + (alt_lines ? 2 : 1) * line_arena, max_long), 256)) != cudaSuccess ||
        ctx->d_bnd.reserve(sizeof(int2) * (size_t)std::max<int64_t>(arenas * slots * bnd_rows, 1)) != cudaSuccess ||
        reserve_zeroed(ctx, ctx->d_chain, sizeof(int4) * (size_t)max_long_bnd) != cudaSuccess ||
        ctx->d_ck.reserve(sizeof(int) * (size_t)(4 + max_ck)) != cudaSuccess ||
        ctx->d_ops.reserve((size_t)cap_total + 64) != cudaSuccess ||
        (ctx->packed_ops && ctx->d_pack.reserve((size_t)(cap_total / 4 + P + 64)) != cudaSuccess) ||
        ctx->d_len.reserve(sizeof(int) * Pz) != cudaSuccess ||
        ctx->d_scores.reserve(sizeof(int) * 3 * Pz) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, TANW_E_NOMEM, "device allocation failed (arena %lld bytes)", (long long)(slots * slot_bytes));
    }

    // ---- per-pair scoring systems / substitution table ------------------------------------------------
    if (multi) {
        if ((size_t)in.n_sc > ctx->h_kparams_cap) {
            if (ctx->h_kparams) cudaFreeHost(ctx->h_kparams);
            ctx->h_kparams = nullptr;
            ctx->h_kparams_cap = 0;
            if (cudaMallocHost((void **)&ctx->h_kparams, sizeof(KParams) * (size_t)in.n_sc) != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, TANW_E_NOMEM, "out of pinned host memory for the scoring systems");
            }
            ctx->h_kparams_cap = (size_t)in.n_sc;
        }
        for (int32_t i = 0; i < in.n_sc; ++i) ctx->h_kparams[i] = make_kparams(&sc[i]);
        if (ctx->d_kparams.reserve(sizeof(KParams) * (size_t)in.n_sc) != cudaSuccess)
            return fail(ctx, TANW_E_NOMEM, "device allocation failed (scoring systems)");
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_kparams.p, ctx->h_kparams, sizeof(KParams) * (size_t)in.n_sc,
                                       cudaMemcpyHostToDevice, ctx->s_in));
        h2d += (int64_t)sizeof(KParams) * in.n_sc;
    }
    if (ctx->use_subst) {
        int rc = upload_subst(ctx, sc, &h2d);
        if (rc) return rc;
    }
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_h2d1, ctx->s_in));
    if (ctx->use_subst && in.symbols_len > 0) {
        // symbol codes index the table: the largest one must be below K (the kernels clamp, so this
        // is an error report, not a safety check)
        int *d_max = (int *)ctx->d_misc.p + 1;
        TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_k, ctx->ev_h2d1, 0));
        TANW_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(int), ctx->s_k));
        const int blocks = (int)std::min<int64_t>((in.symbols_len + 1023) / 1024, ctx->sm_count * 8);
        if (sb == 1) max_symbol_kernel<uint8_t><<<blocks, 256, 0, ctx->s_k>>>((const uint8_t *)ctx->d_sym.p, in.symbols_len, d_max);
        else         max_symbol_kernel<uint16_t><<<blocks, 256, 0, ctx->s_k>>>((const uint16_t *)ctx->d_sym.p, in.symbols_len, d_max);
        TANW_CUDA(ctx, cudaGetLastError());
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->h_misc + 1, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_k));
        TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_k));
        if (ctx->h_misc[1] >= sc->subst_k)
            return fail(ctx, TANW_E_INVALID, "symbol code %d >= subst_k %d", ctx->h_misc[1], sc->subst_k);
    }

    // ---- the tables, chunk by chunk, on the device: now for the three-phase form; a pipelined batch
    // builds a chunk's tables right before its align kernels (run_impl), so that the first chunk's
    // alignment does not wait for the tables of the others --------------------------------------------
    ta.chunk_pairs = chunk_pairs;
    ta.pairs = (PairDesc *)ctx->d_pairs.p;
    ta.route = (unsigned char *)ctx->d_route.p;
    ta.order = (int *)ctx->d_order.p;
    ta.line_sorted = (int *)ctx->d_lsorted.p;
    ta.hist = (int *)ctx->d_hist.p;
    ta.classes = (LineClasses *)ctx->d_classes.p;
    ctx->ta = ta;
    ctx->table_launches = (n_tiles > 0 && !host_tables) ? 1 : 0;
    ctx->timing.table_launches = ctx->table_launches;
    for (int c = 0; c < S; ++c) ctx->chunk[c].tables_built = false;
    if (host_tables) {
        TANW_CUDA(ctx, cudaEventSynchronize(ctx->ev_small));           // the staging buffer's previous upload (if any) has left
        PairDesc *pd = reinterpret_cast<PairDesc *>(ctx->h_small);
        uint8_t *rt = ctx->h_small + (size_t)kHostSurveyPairs * sizeof(PairDesc);
        int64_t at = 0;
        for (int64_t p = 0; p < P; ++p) {
            pd[p].t_off = in.t_off[p]; pd[p].o_off = in.o_off[p]; pd[p].ops_off = at;
            pd[p].n = in.n[p]; pd[p].m = in.m[p];
            rt[p] = (uint8_t)kRouteLong;
            at += (int64_t)in.n[p] + in.m[p];
        }
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_pairs.p, pd, sizeof(PairDesc) * (size_t)P, cudaMemcpyHostToDevice, ctx->s_k));
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_route.p, rt, (size_t)P, cudaMemcpyHostToDevice, ctx->s_k));
        TANW_CUDA(ctx, cudaEventRecord(ctx->ev_small, ctx->s_k));
        for (int c = 0; c < S; ++c) ctx->chunk[c].tables_built = true;
    }
    if (!in.pipelined)
        for (int c = 0; c < S; ++c)
            if (int rc = build_chunk_tables(ctx, c)) return rc;

    BatchArgs &a = ctx->args;
    memset(&a, 0, sizeof a);
    a.sym = (const uint8_t *)ctx->d_sym.p;
    a.pairs = (const PairDesc *)ctx->d_pairs.p;
    a.ptr_arena = (uint8_t *)ctx->d_arena.p;
    a.slot_bytes = slot_bytes;
    a.bnd_arena = (int2 *)ctx->d_bnd.p;
    a.bnd_rows = bnd_rows;
    a.ops = (uint8_t *)ctx->d_ops.p;
    a.ops_len = (int *)ctx->d_len.p;
    a.scores = (int *)ctx->d_scores.p;
    a.kparams = (const KParams *)ctx->d_kparams.p;
    a.sidx = (const int *)ctx->d_sidx.p;
    a.check = (int *)ctx->d_misc.p;

    LineArgs &la = ctx->largs;
    memset(&la, 0, sizeof la);
    la.sym = a.sym;
    la.pairs = a.pairs;
    la.ptr_arena = a.ptr_arena + (size_t)(arenas * slots * slot_bytes);        // behind the page kernel's slots
    ctx->page_slots = slots;
    ctx->alt_lines = alt_lines;
    ctx->line_arena_bytes = (line_arena + 255) / 256 * 256;
    la.slot_bytes = line_slot;
    la.ops = a.ops;
    la.ops_len = a.ops_len;
    la.scores = a.scores;
    la.check = a.check;

    ctx->n_pairs = P;
    ctx->ops_total = cap_total;
    memset(&ctx->timing, 0, sizeof ctx->timing);
    ctx->timing.cells = cells;
    ctx->timing.ptr_bytes = cells;
    ctx->timing.h2d_bytes = h2d;
    ctx->timing.chunks = S;
    ctx->timing.table_launches = ctx->table_launches;
    ctx->timing.host_prepare_ms = host_timer.ms();
    ctx->prepared = true;
    return TANW_OK;
}

int run_impl(tanw_ctx *ctx, bool pipelined)
{
    HostTimer host_timer;
    TANW_ENTER(ctx);
    // everything uploaded -- or, pipelined, only what is not a symbol piece (scoring systems, table)
    if (!pipelined || ctx->multi || ctx->use_subst) TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_k, ctx->ev_h2d1, 0));
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_k0, ctx->s_k));
    int launches = 0;
    TANW_CUDA(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(unsigned) * 4 * kMaxChunks, ctx->s_k));
    TANW_CUDA(ctx, cudaMemsetAsync(ctx->d_misc.p, 0, sizeof(int), ctx->s_k));
    bool cp_forked[kMaxChunks] = {}, cp_lines[kMaxChunks] = {};
    // (Building the tables of all chunks of a line batch up front, before its first line kernel,
    // is slower -- config 3 end to end 1.29 -> 1.58 ms: a chunk's table kernels find free SM slots
    // beside the previous chunk's line kernel, so built lazily they cost nothing.)
    for (int c = 0; c < ctx->n_chunks; ++c) {
        if (int rc = build_chunk_tables(ctx, c)) return rc;
        const ChunkPlan &cp = ctx->chunk[c];
        const bool has_lines = cp.n_octets > 0 || cp.n_quads > 0;
        const bool forked = cp.n_page > 0 && (has_lines || ctx->alternate);
        const bool lines_away = ctx->alt_lines && has_lines;         // the chunk's line kernels leave s_k
        const int lpar = lines_away ? (c & 1) : 0;
        cudaStream_t ls = lines_away ? (lpar ? ctx->s_l1 : ctx->s_l0) : ctx->s_k;
        if (forked || lines_away) TANW_CUDA(ctx, cudaEventRecord(ctx->ev_fork[c], ctx->s_k));
        if (lines_away) TANW_CUDA(ctx, cudaStreamWaitEvent(ls, ctx->ev_fork[c], 0));
        if (pipelined && (has_lines || !forked)) TANW_CUDA(ctx, cudaStreamWaitEvent(ls, ctx->ev_piece[cp.piece], 0));
#ifdef TANW_TUNING
        TANW_CUDA(ctx, cudaEventRecord(ctx->tl_ls[c], ls));
#endif
        if (cp.n_octets > 0) {
            LineArgs la = ctx->largs;
            la.ptr_arena += (size_t)lpar * (size_t)ctx->line_arena_bytes;
            la.sorted = (const int *)ctx->d_lsorted.p + cp.first;
            la.classes = (const LineClasses *)ctx->d_classes.p + 2 * c;
            la.counter = (unsigned *)ctx->d_counter.p + 4 * c + 2;
            la.slot_bytes = ctx->line16_slot;
            la.n_quads = cp.n_octets;
            const int grid = (int)std::min<int64_t>(ctx->line16_grid, ((int64_t)cp.n_octets + kL16Warps - 1) / kL16Warps);
            TANW_CUDA(ctx, launch_lines16(la, ctx->kp, ctx->var, std::max(grid, 1), ls));
            ++launches;
        }
        if (cp.n_quads > 0) {
            LineArgs la = ctx->largs;
            la.ptr_arena += (size_t)lpar * (size_t)ctx->line_arena_bytes;
            la.sorted = (const int *)ctx->d_lsorted.p + cp.first;
            la.classes = (const LineClasses *)ctx->d_classes.p + 2 * c + 1;
            la.counter = (unsigned *)ctx->d_counter.p + 4 * c + 1;
            la.n_quads = cp.n_quads;
            const int grid = (int)std::min<int64_t>(ctx->line_grid, ((int64_t)cp.n_quads + kWarpsPerBlock - 1) / kWarpsPerBlock);
            TANW_CUDA(ctx, launch_lines(la, ctx->kp, ctx->var, ctx->use_subst, std::max(grid, 1), ls));
            ++launches;
        }
        if (ctx->batch_packed && has_lines) {
            if (int rc = pack_chunk(ctx, cp, (1u << kRouteLine) | (1u << kRouteLine16), ls)) return rc;
            ++launches;
        }
#ifdef TANW_TUNING
        TANW_CUDA(ctx, cudaEventRecord(ctx->tl_le[c], ls));
#endif
        if (lines_away) TANW_CUDA(ctx, cudaEventRecord(ctx->ev_lines[c], ls));
        cp_lines[c] = lines_away;
        if (cp.n_page > 0) {
            BatchArgs a = ctx->args;
            a.order = (const int *)ctx->d_order.p + cp.first;
            a.counter = (unsigned *)ctx->d_counter.p + 4 * c;
            a.n_pairs = cp.n_page;
            const int grid = (int)std::min<int64_t>(ctx->grid, ((int64_t)cp.n_page + kWarpsPerBlock - 1) / kWarpsPerBlock);
            const int parity = ctx->alternate ? (c & 1) : 0;
            cudaStream_t st = forked ? (parity ? ctx->s_k3 : ctx->s_k2) : ctx->s_k;
            a.ptr_arena += (size_t)parity * (size_t)(ctx->page_slots * ctx->slot_bytes);
            a.bnd_arena += (size_t)parity * (size_t)(ctx->page_slots * a.bnd_rows);
            if (forked) TANW_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_fork[c], 0));
            if (forked && pipelined) TANW_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_piece[cp.piece], 0));
            TANW_CUDA(ctx, launch_pairs(a, ctx->kp, ctx->var, ctx->page_subst, ctx->batch_sym_bytes, ctx->multi,
                                        std::max(grid, 1), st));
            ++launches;
            if (ctx->batch_packed) {
                if (int rc = pack_chunk(ctx, cp, 1u << kRoutePage, st)) return rc;
                ++launches;
            }
            if (forked) TANW_CUDA(ctx, cudaEventRecord(ctx->ev_pages[c], st));
        }
        cp_forked[c] = forked;
        if (c + 1 == ctx->n_chunks) {
            // join the page kernels before anything else touches the arena, and so that the
            // stream the caller sees (tanw_stream_handle) covers all of the batch's work
            for (int j = 0; j <= c; ++j)
                if (cp_forked[j]) TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_k, ctx->ev_pages[j], 0));
            for (int j = 0; j <= c; ++j)
                if (cp_lines[j]) TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_k, ctx->ev_lines[j], 0));
            for (const LongPair &lp : ctx->longs) {
                const KParams kp = ctx->multi ? ctx->h_kparams[lp.sidx] : ctx->kp;
                int rc = run_long_pair(ctx, lp, kp, ctx->var, &launches);
                if (rc) return rc;
            }
            if (ctx->batch_packed && !ctx->longs.empty()) {
                if (int rc = pack_chunk(ctx, cp, 1u << kRouteLong, ctx->s_k)) return rc;
                ++launches;
            }
        }
        TANW_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[c], ctx->s_k));
    }
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_k1, ctx->s_k));
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_idle, ctx->s_k));
    ctx->timing.kernel_launches = launches;
    ctx->timing.host_run_ms = host_timer.ms();
    ctx->ran = true;
    return TANW_OK;
}

int fetch_impl(tanw_ctx *ctx, uint8_t *ops, const int64_t *ops_off, int64_t ops_capacity,
               int32_t *ops_len, int32_t *scores, const int32_t *n, const int32_t *m)
{
    HostTimer host_timer;
    const int64_t P = ctx->n_pairs;
    if (P > 0 && (!ops_off || !ops_len)) return fail(ctx, TANW_E_INVALID, "NULL output table");
    if (ctx->ops_total > 0 && !ops) return fail(ctx, TANW_E_INVALID, "ops is NULL");
    const bool packed = ctx->batch_packed;
    if (packed && ops_capacity < ctx->ops_total / 4 + P + 1)
        return fail(ctx, TANW_E_INVALID, "packed op buffer too small: needs sum(n+m)/4 + pairs + 1 = %lld bytes",
                    (long long)(ctx->ops_total / 4 + P + 1));
    // the common case first: the caller uses the canonical layout (prefix sums of n+m)
    const bool canonical = packed || (layout_is_canonical(ops_off, n, m, P) && ctx->ops_total <= ops_capacity);
    if (!canonical) {
        for (int64_t p = 0; p < P; ++p) {
            const int64_t cap = (int64_t)n[p] + m[p];
            if (ops_off[p] < 0 || ops_off[p] + cap > ops_capacity)
                return fail(ctx, TANW_E_INVALID, "pair %lld: op buffer too small (needs n+m = %lld bytes at offset %lld)",
                            (long long)p, (long long)cap, (long long)ops_off[p]);
        }
    }
    TANW_ENTER(ctx);
    uint8_t *dst = ops;
    if (!canonical) {
        try {
            ctx->h_stage.resize((size_t)ctx->ops_total);
        } catch (const std::bad_alloc &) {
            return fail(ctx, TANW_E_NOMEM, "out of host memory for the op staging buffer");
        }
        dst = ctx->h_stage.data();
    }
    // small results come back through the context's page-locked staging (see kSmallIo)
    const int64_t ops_bytes = packed ? ctx->ops_total / 4 + P + 1 : ctx->ops_total;
    const int64_t ops_room = (ops_bytes + 15) / 16 * 16, len_room = (4 * P + 15) / 16 * 16;
    const bool small_out = canonical && P <= kHostSurveyPairs && ops_room + len_room + 12 * P + 16 <= kSmallIo;
    uint8_t *const user_ops = ops;
    int32_t *const user_len = ops_len, *const user_scores = scores;
    if (small_out) {
        uint8_t *base = ctx->h_io + kSmallIo;
        dst = base;
        ops_len = reinterpret_cast<int32_t *>(base + ops_room);
        if (scores) scores = reinterpret_cast<int32_t *>(base + ops_room + len_room);
    }
    int64_t d2h = 0;
    for (int c = 0; c < ctx->n_chunks; ++c) {
        const ChunkPlan &cp = ctx->chunk[c];
        TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_chunk[c], 0));
        if (cp.n_page > 0 && (cp.n_octets > 0 || cp.n_quads > 0 || ctx->alternate))
            TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_pages[c], 0));
        if (ctx->alt_lines && (cp.n_octets > 0 || cp.n_quads > 0))
            TANW_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_lines[c], 0));
        if (c == 0) TANW_CUDA(ctx, cudaEventRecord(ctx->ev_d2h0, ctx->s_out));
        if (packed && cp.count > 0) {
            const int64_t lo = (cp.ops_base >> 2) + cp.first, hi = ((cp.ops_base + cp.cap) >> 2) + cp.first + cp.count;
            TANW_CUDA(ctx, cudaMemcpyAsync(dst + lo, (const uint8_t *)ctx->d_pack.p + lo, (size_t)(hi - lo),
                                           cudaMemcpyDeviceToHost, ctx->s_out));
            d2h += hi - lo;
        } else if (cp.cap > 0) {
            TANW_CUDA(ctx, cudaMemcpyAsync(dst + cp.ops_base, (const uint8_t *)ctx->d_ops.p + cp.ops_base, (size_t)cp.cap,
                                           cudaMemcpyDeviceToHost, ctx->s_out));
            d2h += cp.cap;
        }
        if (cp.count > 0) {
            TANW_CUDA(ctx, cudaMemcpyAsync(ops_len + cp.first, (const int *)ctx->d_len.p + cp.first, sizeof(int) * (size_t)cp.count,
                                           cudaMemcpyDeviceToHost, ctx->s_out));
            d2h += (int64_t)sizeof(int) * cp.count;
            if (scores) {
                TANW_CUDA(ctx, cudaMemcpyAsync(scores + 3 * cp.first, (const int *)ctx->d_scores.p + 3 * cp.first,
                                               sizeof(int) * 3 * (size_t)cp.count, cudaMemcpyDeviceToHost, ctx->s_out));
                d2h += (int64_t)sizeof(int) * 3 * cp.count;
            }
        }
#ifdef TANW_TUNING
        TANW_CUDA(ctx, cudaEventRecord(ctx->tl_d[c], ctx->s_out));
#endif
    }
    TANW_CUDA(ctx, cudaMemcpyAsync(ctx->h_misc, ctx->d_misc.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_out));
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_d2h1, ctx->s_out));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
#ifdef TANW_TUNING
    if (getenv("TANW_TIMELINE")) {                                  // tuning builds only: ms since the first upload
        float t = 0.f;
        for (int i = 0; i < kPieces; ++i)
            if (cudaEventElapsedTime(&t, ctx->ev_h2d0, ctx->ev_piece[i]) == cudaSuccess) fprintf(stderr, "piece %d in %.3f\n", i, t);
        for (int c = 0; c < ctx->n_chunks; ++c) {
            float a = 0.f, b = 0.f, d = 0.f;
            cudaEventElapsedTime(&a, ctx->ev_h2d0, ctx->tl_ls[c]);
            cudaEventElapsedTime(&b, ctx->ev_h2d0, ctx->tl_le[c]);
            cudaEventElapsedTime(&d, ctx->ev_h2d0, ctx->tl_d[c]);
            fprintf(stderr, "chunk %d (piece %d, %lld pairs): lines %.3f .. %.3f, copied out %.3f\n", c, ctx->chunk[c].piece,
                    (long long)ctx->chunk[c].count, a, b, d);
        }
        if (cudaEventElapsedTime(&t, ctx->ev_h2d0, ctx->ev_d2h1) == cudaSuccess) fprintf(stderr, "done %.3f\n", t);
        cudaGetLastError();
    }
#endif
    if (ctx->h_misc[0] != 0)
        return fail(ctx, TANW_E_INTERNAL, "device assertion %d failed (TANW_CHECKED build)", ctx->h_misc[0]);
    if (small_out) {
        if (ops_bytes > 0) memcpy(user_ops, dst, (size_t)ops_bytes);
        if (P > 0) memcpy(user_len, ops_len, sizeof(int32_t) * (size_t)P);
        if (user_scores && P > 0) memcpy(user_scores, scores, sizeof(int32_t) * 3 * (size_t)P);
        ops_len = user_len;
    }
    if (!canonical) {
        int64_t at = 0;
        for (int64_t p = 0; p < P; ++p) {
            memcpy(ops + ops_off[p], ctx->h_stage.data() + at, (size_t)ops_len[p]);
            at += (int64_t)n[p] + m[p];
        }
    }
    ctx->timing.d2h_bytes = d2h;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_h2d0, ctx->ev_h2d1) == cudaSuccess) ctx->timing.h2d_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess) ctx->timing.kernel_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev_d2h0, ctx->ev_d2h1) == cudaSuccess) ctx->timing.d2h_ms = ms;
    cudaGetLastError();
    ctx->timing.host_fetch_ms = host_timer.ms();
    return TANW_OK;
}

int guarded_prepare(tanw_ctx *ctx, const PrepareInput &in)
{
    // No exception may cross the C ABI: host allocations can throw.
    try {
        return prepare_impl(ctx, in);
    } catch (const std::bad_alloc &) {
        return fail(ctx, TANW_E_NOMEM, "out of host memory while preparing the batch");
    } catch (...) {
        return fail(ctx, TANW_E_INVALID, "unexpected exception while preparing the batch");
    }
}

}  // namespace

extern "C" {

int tanw_version(void) { return 200; }

const char *tanw_last_error(const tanw_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_last_error.c_str();
}

int tanw_device_count(int *count)
{
    if (!count) return fail(nullptr, TANW_E_INVALID, "count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(nullptr, TANW_E_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = c;
    return TANW_OK;
}

int tanw_device_query(int device, tanw_device_info *out)
{
    if (!out) return fail(nullptr, TANW_E_INVALID, "out is NULL");
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess)
        return fail(nullptr, TANW_E_NODEVICE, "device %d: %s", device, cudaGetErrorString(e));
    memset(out, 0, sizeof *out);
    snprintf(out->name, sizeof out->name, "%.127s", prop.name);
    out->cc_major = prop.major;
    out->cc_minor = prop.minor;
    out->sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    out->clock_khz = khz;
    out->total_mem_bytes = (int64_t)prop.totalGlobalMem;
    size_t fr = 0, tot = 0;
    DeviceGuard guard(device);
    if (guard.err == cudaSuccess && cudaMemGetInfo(&fr, &tot) == cudaSuccess)
        out->free_mem_bytes = (int64_t)fr;
    cudaGetLastError();
    return TANW_OK;
}

int tanw_create(int device, tanw_ctx **out)
{
    if (!out) return fail(nullptr, TANW_E_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(nullptr, TANW_E_NODEVICE, "no CUDA device visible (libtanw has no CPU fallback)");
    if (device < 0 || device >= count)
        return fail(nullptr, TANW_E_NODEVICE, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    if (!device_is_blackwell(device, &prop))
        return fail(nullptr, TANW_E_NODEVICE,
                    "device %d (%s, sm_%d%d) is not an sm_100 part; libtanw is built for sm_100a only",
                    device, prop.name, prop.major, prop.minor);
    tanw_ctx *ctx = new (std::nothrow) tanw_ctx();
    if (!ctx) return fail(nullptr, TANW_E_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = (int64_t)prop.totalGlobalMem;
    memset(&ctx->timing, 0, sizeof ctx->timing);
    DeviceGuard guard(device);
    cudaError_t e = guard.err;
    cudaStream_t *streams[] = { &ctx->s_in, &ctx->s_k, &ctx->s_k2, &ctx->s_k3, &ctx->s_l0, &ctx->s_l1, &ctx->s_out };
    for (auto s : streams)
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
    cudaEvent_t *evs[] = { &ctx->ev_h2d0, &ctx->ev_h2d1, &ctx->ev_k0, &ctx->ev_k1, &ctx->ev_d2h0, &ctx->ev_d2h1 };
    for (auto ev : evs)
        if (e == cudaSuccess) e = cudaEventCreate(ev);
    cudaEvent_t *plain[] = { &ctx->ev_tab, &ctx->ev_survey, &ctx->ev_idle, &ctx->ev_small };
    for (auto ev : plain)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    for (int i = 0; i < kPieces; ++i)
#ifdef TANW_TUNING
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_piece[i]);
#else
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_piece[i], cudaEventDisableTiming);
#endif
    for (int i = 0; i < kMaxChunks; ++i) {
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_pages[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_lines[i], cudaEventDisableTiming);
#ifdef TANW_TUNING
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->tl_ls[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->tl_le[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->tl_d[i]);
#endif
    }
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_survey, sizeof(Survey));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_misc, 256);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_small, (size_t)kHostSurveyPairs * (sizeof(PairDesc) + 8));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_io, (size_t)(2 * kSmallIo));
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_idle, ctx->s_k);
    if (e != cudaSuccess) {
        int rc = fail(nullptr, TANW_E_CUDA, "context setup on device %d: %s", device, cudaGetErrorString(e));
        cudaGetLastError();
        tanw_destroy(ctx);
        return rc;
    }
    memset(ctx->h_misc, 0, 256);
    ctx->occ_plain = std::max(1, pairs_blocks_per_sm(false));
    ctx->occ_subst = std::max(1, pairs_blocks_per_sm(true));
    ctx->occ_line = std::max(1, lines_blocks_per_sm());
    ctx->occ_line16 = std::max(1, lines16_blocks_per_sm());
    {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        ctx->long_capacity = coop ? std::max(1, long_blocks_per_sm()) * ctx->sm_count : 0;
        cudaGetLastError();
    }
    *out = ctx;
    return TANW_OK;
}

int tanw_destroy(tanw_ctx *ctx)
{
    if (!ctx) return TANW_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t streams[] = { ctx->s_in, ctx->s_k, ctx->s_k2, ctx->s_k3, ctx->s_l0, ctx->s_l1, ctx->s_out };
    for (auto s : streams)
        if (s) cudaStreamSynchronize(s);
    DevBuf *bufs[] = { &ctx->d_sym, &ctx->d_n, &ctx->d_m, &ctx->d_toff, &ctx->d_ooff, &ctx->d_pairs, &ctx->d_route,
                       &ctx->d_order, &ctx->d_lsorted, &ctx->d_hist, &ctx->d_classes, &ctx->d_tilesums, &ctx->d_survey,
                       &ctx->d_counter, &ctx->d_arena, &ctx->d_bnd, &ctx->d_ops, &ctx->d_len, &ctx->d_scores,
                       &ctx->d_subst, &ctx->d_chain, &ctx->d_ck, &ctx->d_kparams, &ctx->d_sidx, &ctx->d_misc, &ctx->d_pack };
    for (auto b : bufs) b->release();
    cudaEvent_t evs[] = { ctx->ev_h2d0, ctx->ev_h2d1, ctx->ev_k0, ctx->ev_k1, ctx->ev_d2h0, ctx->ev_d2h1,
                          ctx->ev_tab, ctx->ev_survey, ctx->ev_idle, ctx->ev_small };
    for (auto ev : evs)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : ctx->ev_piece)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : ctx->ev_chunk)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : ctx->ev_fork)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : ctx->ev_pages)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : ctx->ev_lines)
        if (ev) cudaEventDestroy(ev);
#ifdef TANW_TUNING
    for (int i = 0; i < kMaxChunks; ++i) {
        if (ctx->tl_ls[i]) cudaEventDestroy(ctx->tl_ls[i]);
        if (ctx->tl_le[i]) cudaEventDestroy(ctx->tl_le[i]);
        if (ctx->tl_d[i]) cudaEventDestroy(ctx->tl_d[i]);
    }
#endif
    if (ctx->h_survey) cudaFreeHost(ctx->h_survey);
    if (ctx->h_misc) cudaFreeHost(ctx->h_misc);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->h_io) cudaFreeHost(ctx->h_io);
    if (ctx->h_subst) cudaFreeHost(ctx->h_subst);
    if (ctx->h_kparams) cudaFreeHost(ctx->h_kparams);
    for (auto s : streams)
        if (s) cudaStreamDestroy(s);
    cudaGetLastError();
    delete ctx;
    return TANW_OK;
}

int tanw_set_arena_limit(tanw_ctx *ctx, int64_t bytes)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (bytes < 0) return fail(ctx, TANW_E_INVALID, "arena limit must be >= 0");
    ctx->arena_limit = bytes;
    ctx->prepared = false;
    return TANW_OK;
}

int tanw_set_long_threshold(tanw_ctx *ctx, int64_t cells)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (cells < 1) return fail(ctx, TANW_E_INVALID, "long-pair threshold must be >= 1 cell");
    ctx->long_cells = cells;
    ctx->prepared = false;
    return TANW_OK;
}

int tanw_set_symbol_bytes(tanw_ctx *ctx, int bytes)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (bytes != 1 && bytes != 2) return fail(ctx, TANW_E_INVALID, "symbol width must be 1 or 2 bytes");
    ctx->sym_bytes = bytes;
    ctx->prepared = false;
    return TANW_OK;
}

int tanw_set_long_band_rows(tanw_ctx *ctx, int rows)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (rows < 0) return fail(ctx, TANW_E_INVALID, "band height must be >= 0 rows");
    ctx->long_band_rows = rows;
    ctx->prepared = false;
    return TANW_OK;
}

int tanw_set_packed_ops(tanw_ctx *ctx, int enabled)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    ctx->packed_ops = enabled != 0;
    ctx->prepared = false;
    return TANW_OK;
}

int tanw_set_line_kernel(tanw_ctx *ctx, int mode)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (mode < 0 || mode > 2) return fail(ctx, TANW_E_INVALID, "line kernel mode must be 0, 1 or 2");
    ctx->line_mode = mode;
    ctx->prepared = false;
    return TANW_OK;
}

int tanw_stream_handle(tanw_ctx *ctx, uint64_t *out)
{
    if (!ctx || !out) return fail(ctx, TANW_E_INVALID, "NULL argument");
    *out = (uint64_t)(uintptr_t)ctx->s_k;
    return TANW_OK;
}

int tanw_sync(tanw_ctx *ctx)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    TANW_ENTER(ctx);
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_k));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_k2));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_k3));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_l0));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_l1));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    return TANW_OK;
}

int tanw_batch_prepare(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                       const int64_t *t_off, const int32_t *n, const int64_t *o_off,
                       const int32_t *m, int64_t n_pairs, const tanw_scoring *sc)
{
    PrepareInput in = { symbols, symbols_len, t_off, o_off, n, m, n_pairs, sc, 1, nullptr, false };
    return guarded_prepare(ctx, in);
}

int tanw_batch_prepare_multi(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                             const int64_t *t_off, const int32_t *n, const int64_t *o_off,
                             const int32_t *m, int64_t n_pairs, const tanw_scoring *scorings,
                             int32_t n_scorings, const int32_t *scoring_idx)
{
    if (ctx && n_pairs > 0 && !scoring_idx) return fail(ctx, TANW_E_INVALID, "scoring_idx is NULL");
    static const int32_t none = 0;
    PrepareInput in = { symbols, symbols_len, t_off, o_off, n, m, n_pairs, scorings, n_scorings,
                        scoring_idx ? scoring_idx : &none, false };
    return guarded_prepare(ctx, in);
}

int tanw_batch_rescore(tanw_ctx *ctx, const tanw_scoring *sc)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (!ctx->prepared) return fail(ctx, TANW_E_STATE, "tanw_batch_rescore before tanw_batch_prepare");
    if (!sc) return fail(ctx, TANW_E_INVALID, "scoring is NULL");
    if (ctx->multi) return fail(ctx, TANW_E_INVALID, "rescore does not apply to a batch with per-pair scoring systems");
    if ((sc->subst != nullptr) != ctx->use_subst)
        return fail(ctx, TANW_E_INVALID, "rescore cannot switch between an equality scorer and a table");
    if (sc->subst && sc->subst_k != ctx->kp.subst_k)
        return fail(ctx, TANW_E_INVALID, "rescore needs a table of the same size (K = %d)", ctx->kp.subst_k);
    if (!in_range(scoring_pmax(sc), ctx->max_nm)) return fail(ctx, TANW_E_RANGE, "%s", kRangeMessage);
    if (ctx->line16_max_n > line16_limit(sc))
        return fail(ctx, TANW_E_STATE, "this scoring system cannot use the 16-bit line kernel the batch was routed to "
                    "(gap opens <= 0, match >= mismatch, small scores): prepare the batch again");
    TANW_ENTER(ctx);
    const KParams old = ctx->kp;
    ctx->kp = make_kparams(sc);
    ctx->kp.subst = old.subst;
    ctx->kp.subst_k = old.subst_k;
    ctx->var = ctx->batch_sym_bytes == 2 ? 0 : variant_of(sc);
    if (sc->subst) {
        int rc = upload_subst(ctx, sc, nullptr);
        if (rc) return rc;
        TANW_CUDA(ctx, cudaEventRecord(ctx->ev_h2d1, ctx->s_in));
    }
    ctx->ran = false;
    return TANW_OK;
}

int tanw_batch_run(tanw_ctx *ctx)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (!ctx->prepared) return fail(ctx, TANW_E_STATE, "tanw_batch_run before tanw_batch_prepare");
    return run_impl(ctx, false);
}

int tanw_batch_fetch(tanw_ctx *ctx, uint8_t *ops, const int64_t *ops_off, int64_t ops_capacity,
                     int32_t *ops_len, int32_t *scores)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (!ctx->ran) return fail(ctx, TANW_E_STATE, "tanw_batch_fetch before tanw_batch_run");
    if ((int64_t)ctx->h_n.size() != ctx->n_pairs)
        return fail(ctx, TANW_E_STATE, "tanw_batch_fetch needs a batch prepared by tanw_batch_prepare");
    return fetch_impl(ctx, ops, ops_off, ops_capacity, ops_len, scores, ctx->h_n.data(), ctx->h_m.data());
}

static int align_impl(tanw_ctx *ctx, PrepareInput in, uint8_t *ops, const int64_t *ops_off,
                      int64_t ops_capacity, int32_t *ops_len, int32_t *scores)
{
    in.pipelined = true;
    int rc = guarded_prepare(ctx, in);
    if (rc) return rc;
    ctx->h_n.clear();
    ctx->h_m.clear();
    rc = run_impl(ctx, true);
    if (rc) return rc;
    return fetch_impl(ctx, ops, ops_off, ops_capacity, ops_len, scores, in.n, in.m);
}

int tanw_align_batch(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                     const int64_t *t_off, const int32_t *n, const int64_t *o_off, const int32_t *m,
                     int64_t n_pairs, const tanw_scoring *scoring, uint8_t *ops, const int64_t *ops_off,
                     int64_t ops_capacity, int32_t *ops_len, int32_t *scores)
{
    PrepareInput in = { symbols, symbols_len, t_off, o_off, n, m, n_pairs, scoring, 1, nullptr, true };
    return align_impl(ctx, in, ops, ops_off, ops_capacity, ops_len, scores);
}

int tanw_align_batch_multi(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                           const int64_t *t_off, const int32_t *n, const int64_t *o_off, const int32_t *m,
                           int64_t n_pairs, const tanw_scoring *scorings, int32_t n_scorings,
                           const int32_t *scoring_idx, uint8_t *ops, const int64_t *ops_off,
                           int64_t ops_capacity, int32_t *ops_len, int32_t *scores)
{
    if (ctx && n_pairs > 0 && !scoring_idx) return fail(ctx, TANW_E_INVALID, "scoring_idx is NULL");
    static const int32_t none = 0;
    PrepareInput in = { symbols, symbols_len, t_off, o_off, n, m, n_pairs, scorings, n_scorings,
                        scoring_idx ? scoring_idx : &none, true };
    return align_impl(ctx, in, ops, ops_off, ops_capacity, ops_len, scores);
}

// One batch over several devices: shard d = pairs [bounds[d], bounds[d+1]) on ctxs[d], one host
// thread per shard.  Pairs are independent (alignToOCR.py:273 is one call per page), so there is
// no exchange step: every shard writes its op strings, lengths and scores straight into its slice
// of the caller's arrays.
int tanw_align_batch_sharded(tanw_ctx *const *ctxs, int32_t n_ctx, const int64_t *bounds,
                             const uint8_t *symbols, int64_t symbols_len,
                             const int64_t *t_off, const int32_t *n, const int64_t *o_off, const int32_t *m,
                             int64_t n_pairs, const tanw_scoring *scoring, uint8_t *ops, const int64_t *ops_off,
                             int64_t ops_capacity, int32_t *ops_len, int32_t *scores)
{
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return fail(nullptr, TANW_E_INVALID, "no context");
    tanw_ctx *first = ctxs[0];
    if (!bounds) return fail(first, TANW_E_INVALID, "bounds is NULL");
    if (n_pairs < 0) return fail(first, TANW_E_INVALID, "n_pairs must be >= 0");
    if (n_pairs > 0 && (!t_off || !n || !o_off || !m || !ops_off || !ops_len))
        return fail(first, TANW_E_INVALID, "NULL pair table");
    if (bounds[0] != 0 || bounds[n_ctx] != n_pairs) return fail(first, TANW_E_INVALID, "bounds must run from 0 to n_pairs");
    for (int d = 0; d < n_ctx; ++d) {
        if (!ctxs[d]) return fail(first, TANW_E_INVALID, "context %d is NULL", d);
        if (bounds[d + 1] < bounds[d]) return fail(first, TANW_E_INVALID, "bounds must not decrease");
        if (ctxs[d]->packed_ops) return fail(first, TANW_E_STATE, "packed op strings are per context: not with a sharded batch");
        if (ctxs[d]->sym_bytes != first->sym_bytes) return fail(first, TANW_E_STATE, "contexts differ in symbol width");
        for (int e = 0; e < d; ++e)
            if (ctxs[e] == ctxs[d]) return fail(first, TANW_E_INVALID, "context %d is listed twice", d);
    }
    std::vector<int> rcs((size_t)n_ctx, TANW_OK);
    auto shard = [&](int d) {
        const int64_t lo = bounds[d], hi = bounds[d + 1], P = hi - lo;
        if (P <= 0) return;
        try {
            // the smallest contiguous slice of the symbol buffer that holds the shard
            int64_t s_lo = INT64_MAX, s_hi = 0;
            for (int64_t p = lo; p < hi; ++p) {
                s_lo = std::min(s_lo, std::min(t_off[p], o_off[p]));
                s_hi = std::max(s_hi, std::max(t_off[p] + std::max(n[p], 0), o_off[p] + std::max(m[p], 0)));
            }
            if (s_lo < 0 || s_hi > symbols_len || s_lo > s_hi) { s_lo = 0; s_hi = symbols_len; }   // the shard's own check names the pair
            std::vector<int64_t> t((size_t)P), o((size_t)P), oo((size_t)P);
            const int64_t ops_lo = ops_off[lo];
            for (int64_t p = 0; p < P; ++p) {
                t[(size_t)p] = t_off[lo + p] - s_lo;
                o[(size_t)p] = o_off[lo + p] - s_lo;
                oo[(size_t)p] = ops_off[lo + p] - ops_lo;
            }
            const int64_t ops_hi = hi < n_pairs ? ops_off[hi] : ops_capacity;
            rcs[(size_t)d] = tanw_align_batch(ctxs[d], symbols + s_lo * ctxs[d]->sym_bytes, s_hi - s_lo, t.data(), n + lo,
                                              o.data(), m + lo, P, scoring, ops + ops_lo, oo.data(), ops_hi - ops_lo,
                                              ops_len + lo, scores ? scores + 3 * lo : nullptr);
        } catch (const std::bad_alloc &) {
            rcs[(size_t)d] = fail(ctxs[d], TANW_E_NOMEM, "out of host memory for the shard tables");
        }
    };
    std::vector<std::thread> threads;
    try {
        for (int d = 1; d < n_ctx; ++d) threads.emplace_back(shard, d);
    } catch (...) {
        for (auto &th : threads) th.join();
        return fail(first, TANW_E_INTERNAL, "could not start a host thread per device");
    }
    shard(0);
    for (auto &th : threads) th.join();
    for (int d = 0; d < n_ctx; ++d)
        if (rcs[(size_t)d] != TANW_OK) {
            if (d > 0) return fail(first, rcs[(size_t)d], "shard %d (pairs %lld..%lld): %s", d, (long long)bounds[d],
                                   (long long)bounds[d + 1], ctxs[d]->err.c_str());
            return rcs[0];
        }
    return TANW_OK;
}

int tanw_last_timing(tanw_ctx *ctx, tanw_timing *out)
{
    if (!ctx || !out) return fail(ctx, TANW_E_INVALID, "NULL argument");
    if (ctx->ran) {
        // valid once the stream has drained past the kernel (after fetch or tanw_sync)
        float ms = 0.f;
        if (cudaEventQuery(ctx->ev_k1) == cudaSuccess &&
            cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess)
            ctx->timing.kernel_ms = ms;
        if (cudaEventQuery(ctx->ev_h2d1) == cudaSuccess &&
            cudaEventElapsedTime(&ms, ctx->ev_h2d0, ctx->ev_h2d1) == cudaSuccess)
            ctx->timing.h2d_ms = ms;
        cudaGetLastError();
    }
    *out = ctx->timing;
    return TANW_OK;
}

int tanw_measure_int32_peak(tanw_ctx *ctx, int which, double *lane_ops_per_s)
{
    if (!ctx || !lane_ops_per_s) return fail(ctx, TANW_E_INVALID, "NULL argument");
    if (which < 0 || which > 3) return fail(ctx, TANW_E_INVALID, "which must be 0, 1, 2 or 3");
    TANW_ENTER(ctx);
    ctx->prepared = ctx->ran = false;                     // borrows the score / counter buffers of the batch
    if (ctx->d_counter.reserve(sizeof(unsigned) * 4 * kMaxChunks + 256) != cudaSuccess ||
        ctx->d_scores.reserve(4096 * sizeof(int)) != cudaSuccess)
        return fail(ctx, TANW_E_NOMEM, "device allocation failed");
    const int iters = 1 << 13, blocks = ctx->sm_count * 8, threads = 256;
    const int *src = (const int *)ctx->d_scores.p;       // any initialised words will do
    TANW_CUDA(ctx, cudaMemsetAsync(ctx->d_scores.p, 1, 4096 * sizeof(int), ctx->s_k));
    struct EventPair {                                    // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } ev;
    TANW_CUDA(ctx, cudaEventCreate(&ev.a));
    TANW_CUDA(ctx, cudaEventCreate(&ev.b));
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        TANW_CUDA(ctx, cudaEventRecord(e0, ctx->s_k));
        int *sink = (int *)ctx->d_counter.p + 4 * kMaxChunks;
        if (which == 0)      int32_peak_kernel<0><<<blocks, threads, 0, ctx->s_k>>>(iters, src, sink, 3, 5);
        else if (which == 1) int32_peak_kernel<1><<<blocks, threads, 0, ctx->s_k>>>(iters, src, sink, 3, 5);
        else if (which == 2) int32_peak_kernel<2><<<blocks, threads, 0, ctx->s_k>>>(iters, src, sink, 3, 5);
        else                 int32_peak_kernel<3><<<blocks, threads, 0, ctx->s_k>>>(iters, src, sink, 3, 5);
        TANW_CUDA(ctx, cudaGetLastError());
        TANW_CUDA(ctx, cudaEventRecord(e1, ctx->s_k));
        TANW_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        TANW_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        // instructions per chain-iteration: IADD3 x1 (two adds merged), VIMNMX x2, VIADDMNMX x2, VIADD + LOP3
        const double instr = (double)blocks * threads * (double)iters * 16.0 * (which == 0 ? 1.0 : 2.0);
        if (rep > 0) best = std::max(best, instr / (ms * 1e-3));
    }
    *lane_ops_per_s = best;
    return TANW_OK;
}

}  // extern "C"
