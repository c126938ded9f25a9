// tanw.cu -- host side of libtanw.so: the C ABI declared in include/tanw.h.
//
// Replaces the per-page call textSeqCompare.perform_alignment (textSeqCompare.py:13-177,
// call site alignToOCR.py:273-274) by a batched device implementation.  No CPU fallback:
// every entry either runs on an sm_100 device or returns an error.
#include "tanw.h"
#include "tanw_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

using namespace tanw;

namespace {

thread_local std::string g_last_error;   // for failures that have no context yet

// Page-locked host vectors: the pair / order / quad tables are rebuilt for every batch and
// uploaded with cudaMemcpyAsync, which is only asynchronous (and only reaches full PCIe speed)
// from pinned memory.  Capacity is kept across batches, so the allocation cost is paid once.
template <class T>
struct PinnedAlloc {
    typedef T value_type;
    PinnedAlloc() = default;
    template <class U> PinnedAlloc(const PinnedAlloc<U> &) {}
    T *allocate(size_t n)
    {
        void *p = nullptr;
        if (cudaMallocHost(&p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); throw std::bad_alloc(); }
        return static_cast<T *>(p);
    }
    void deallocate(T *p, size_t) { cudaFreeHost(p); }
    template <class U> bool operator==(const PinnedAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const PinnedAlloc<U> &) const { return false; }
};
template <class T> using pinned_vector = std::vector<T, PinnedAlloc<T>>;

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
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct tanw_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_h2d0 = nullptr, ev_h2d1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr,
                ev_d2h0 = nullptr, ev_d2h1 = nullptr;
    std::string err;
    int64_t arena_limit = 0;
    int64_t max_nm = 0;                   // largest n+m of the prepared batch (range check on rescore)
    int64_t total_mem = 0;

    DevBuf d_sym, d_pairs, d_order, d_counter, d_arena, d_bnd, d_ops, d_len, d_scores, d_subst, d_prog, d_quads;
    std::vector<int> h_line;              // pairs routed to the four-per-warp line kernel
    std::vector<int> h_line_key, h_line_count, h_line_sorted, h_line_skey;   // counting sort scratch (kept across batches)
    pinned_vector<int4> h_quads;
    LineArgs largs;
    int line_grid = 0, occ_line = 0;
    bool use_lines = true;
    std::vector<int> h_long;              // pairs routed to the chained-pass (whole-GPU) path
    int long_capacity = 0;                // resident warps for a cooperative launch
    int long_epoch = 0;                   // stamps the chain records of a launch
    DevBuf d_chain;
    DevBuf d_ck;                          // row-band checkpoints: 4 ints of traceback state, then 3*m ints per band edge
    std::vector<int2> h_long_geo;         // per long pair: (stripe strip width, rows per band)
    int long_band_rows = 0;               // 0 = one band unless the pointer block exceeds the arena limit
    int sym_bytes = 1;                    // 1: uint8 symbol codes; 2: uint16 (page kernel only)
    int batch_sym_bytes = 1;              // width the prepared batch was uploaded with
    int64_t long_cells = int64_t(1) << 26;   // pairs with n*m >= this use the chained-pass path
    pinned_vector<PairDesc> h_pairs;
    pinned_vector<int> h_order, h_order_sorted;
    std::vector<int64_t> h_ops_off;       // canonical device layout: prefix sums of n+m
    std::vector<uint8_t> h_stage;         // used when the caller's op layout is not canonical

    // state of the prepared batch
    bool prepared = false, ran = false;
    int64_t n_pairs = 0, ops_total = 0;
    KParams kp;
    bool use_subst = false;
    bool opens_nonpositive = false;       // gap_open_x <= 0 and gap_open_y <= 0
    BatchArgs args;
    int grid = 0;
    int occ_plain = 0, occ_subst = 0;
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

bool device_is_blackwell(int device, cudaDeviceProp *prop_out)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return false;
    if (prop_out) *prop_out = prop;
    return prop.major == 10;
}

// int32 fixed point: every finite intermediate must stay far away from kNeg = -2^30.
// |value| <= (n+m+2) * max|param| ; carried << 6 and offset by up to ex*n once more: 2^22 * 2 * 64 = 2^29 < 2^30.
bool scoring_in_range(const tanw_scoring *s, int64_t max_n_plus_m)
{
    int64_t pmax = 1;
    auto upd = [&](int64_t v) { pmax = std::max<int64_t>(pmax, v < 0 ? -v : v); };
    upd(s->match); upd(s->mismatch); upd(s->boundary_gap);
    upd((int64_t)s->gap_open_x + s->gap_extend_x); upd(s->gap_extend_x);
    upd((int64_t)s->gap_open_y + s->gap_extend_y); upd(s->gap_extend_y);
    if (s->subst)
        for (int64_t i = 0; i < (int64_t)s->subst_k * s->subst_k; ++i) upd(s->subst[i]);
    return (max_n_plus_m + 2) * pmax < (int64_t(1) << 22);
}

void fill_kparams(KParams &kp, const tanw_scoring *sc)
{
    const int *keep_tab = kp.subst;
    const int keep_k = kp.subst_k;
    memset(&kp, 0, sizeof kp);
    kp.maT = (sc->match << kShift) | kTagM;
    kp.miT = (sc->mismatch << kShift) | kTagM;
    kp.ox = (sc->gap_open_x + sc->gap_extend_x) * (1 << kShift);
    kp.ex = sc->gap_extend_x * (1 << kShift);
    kp.oy = (sc->gap_open_y + sc->gap_extend_y) * (1 << kShift);
    kp.ey = sc->gap_extend_y * (1 << kShift);
    kp.bg = sc->boundary_gap * (1 << kShift);
    kp.subst = keep_tab;
    kp.subst_k = keep_k;
}

int upload_subst(tanw_ctx *ctx, const tanw_scoring *sc, int64_t *h2d)
{
    const size_t kk = (size_t)sc->subst_k * (size_t)sc->subst_k;
    std::vector<int> tab(kk);
    for (size_t i = 0; i < kk; ++i) tab[i] = (sc->subst[i] * (1 << kShift)) | kTagM;
    if (ctx->d_subst.reserve(sizeof(int) * kk) != cudaSuccess)
        return fail(ctx, TANW_E_NOMEM, "device allocation failed (substitution table)");
    TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_subst.p, tab.data(), sizeof(int) * kk, cudaMemcpyHostToDevice, ctx->stream));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // tab is a local
    ctx->kp.subst = (const int *)ctx->d_subst.p;
    ctx->kp.subst_k = sc->subst_k;
    if (h2d) *h2d += (int64_t)(sizeof(int) * kk);
    return TANW_OK;
}

// Which instantiation serves the prepared scoring system: -1 substitution table, 0 general,
// 1 gap opens <= 0, 2 gap opens <= 0 and gap_extend_y == 0.
int kernel_variant(const tanw_ctx *ctx)
{
    if (ctx->use_subst) return -1;
    if (!ctx->opens_nonpositive) return 0;
    return ctx->kp.ey == 0 ? 2 : 1;
}

// The chain records carry an epoch stamp; fresh memory must not contain a live one.
cudaError_t reserve_zeroed(tanw_ctx *ctx, DevBuf &buf, size_t bytes)
{
    if (bytes <= buf.cap) return cudaSuccess;
    cudaError_t e = buf.reserve(bytes);
    if (e == cudaSuccess && buf.cap) {
        e = cudaMemsetAsync(buf.p, 0, buf.cap, ctx->stream);
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
    if (const char *e = getenv("TANW_LONG_C")) {          // tuning experiments only
        const int c = atoi(e);
        if (c >= 4 && c <= kMaxC && c % 4 == 0 && (m + 32 * c - 1) / (32 * c) <= ctx->long_capacity) return c;
    }
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
constexpr int kLineKeys = 4 * (kLineMaxN + 1);      // quad sort keys: 4 strip-width classes x (n + 1)
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
int run_long_pair(tanw_ctx *ctx, int p, int2 geo, int *launches)
{
    const PairDesc &pd = ctx->h_pairs[(size_t)p];
    const int R = geo.y;
    const int B = (pd.n + R - 1) / R;
    int *state = (int *)ctx->d_ck.p;
    int *ck = state + 4;
    LongArgs la;
    la.T = (const uint8_t *)ctx->d_sym.p + pd.t_off;
    la.O = (const uint8_t *)ctx->d_sym.p + pd.o_off;
    la.n = pd.n;
    la.m = pd.m;
    la.ptr = (uint8_t *)ctx->d_arena.p;
    la.chain = (int4 *)ctx->d_chain.p;
    la.chain_stride = (long long)std::min(R, pd.n) + 4;
    la.cfull = geo.x;
    la.scores = (int *)ctx->d_scores.p + 3 * (size_t)p;
    const int npass = (pd.m + 32 * la.cfull - 1) / (32 * la.cfull);
    const int var = kernel_variant(ctx);
    const void *fn = var < 0 ? (const void *)align_long_kernel<true, 0>
                   : var == 2 ? (const void *)align_long_kernel<false, 2>
                   : var == 1 ? (const void *)align_long_kernel<false, 1>
                              : (const void *)align_long_kernel<false, 0>;
    auto fill_band = [&](int b, bool store) -> int {
        la.r0 = b * R;
        la.nb = std::min(R, pd.n - la.r0);
        la.ck_in = b > 0 ? ck + (size_t)(b - 1) * 3 * (size_t)pd.m : nullptr;
        la.ck_out = (!store && b < B - 1) ? ck + (size_t)b * 3 * (size_t)pd.m : nullptr;
        la.store = store ? 1 : 0;
        la.epoch = ++ctx->long_epoch;
        if (la.epoch == 0) la.epoch = ++ctx->long_epoch;
        long_col0_kernel<<<(la.nb + 255) / 256, 256, 0, ctx->stream>>>(
            la.chain + (size_t)npass * (size_t)la.chain_stride, la.nb, la.r0, ctx->kp.bg, la.epoch);
        TANW_CUDA(ctx, cudaGetLastError());
        ++*launches;
        for (int w0 = 0; w0 < npass; w0 += ctx->long_capacity) {
            la.pass0 = w0;
            const int grid = std::min(ctx->long_capacity, npass - w0);
            void *args[] = { (void *)&la, (void *)&ctx->kp };
            TANW_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(32), args, 0, ctx->stream));
            ++*launches;
        }
        return TANW_OK;
    };
    for (int b = 0; b < B - 1; ++b)
        if (int rc = fill_band(b, false)) return rc;
    static const bool skip_trace = getenv("TANW_DEBUG_SKIP_TRACE") != nullptr;   // timing experiments only
    for (int b = B - 1; b >= 0; --b) {
        if (int rc = fill_band(b, true)) return rc;
        if (skip_trace) continue;
        trace_long_kernel<<<1, 32, 0, ctx->stream>>>(la.ptr, pd.n, pd.m, la.cfull, la.r0, la.nb, b == B - 1, b == 0,
                                                     state, (uint8_t *)ctx->d_ops.p + pd.ops_off,
                                                     (int *)ctx->d_len.p + p);
        TANW_CUDA(ctx, cudaGetLastError());
        ++*launches;
    }
    return TANW_OK;
}

}  // namespace

extern "C" {

int tanw_version(void) { return 100; }

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
    snprintf(out->name, sizeof out->name, "%s", prop.name);
    out->cc_major = prop.major;
    out->cc_minor = prop.minor;
    out->sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    out->clock_khz = khz;
    out->total_mem_bytes = (int64_t)prop.totalGlobalMem;
    size_t fr = 0, tot = 0;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cudaSetDevice(device) == cudaSuccess && cudaMemGetInfo(&fr, &tot) == cudaSuccess)
        out->free_mem_bytes = (int64_t)fr;
    cudaSetDevice(cur);
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
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaEvent_t *evs[] = { &ctx->ev_h2d0, &ctx->ev_h2d1, &ctx->ev_k0, &ctx->ev_k1, &ctx->ev_d2h0, &ctx->ev_d2h1 };
    for (auto ev : evs)
        if (e == cudaSuccess) e = cudaEventCreate(ev);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_plain, align_pairs_kernel<false, 2>,
                                                          kWarpsPerBlock * 32, 0);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_subst, align_pairs_kernel<true, 0>,
                                                          kWarpsPerBlock * 32, 0);
    if (e != cudaSuccess) {
        int rc = fail(nullptr, TANW_E_CUDA, "context setup on device %d: %s", device, cudaGetErrorString(e));
        tanw_destroy(ctx);
        return rc;
    }
    if (ctx->occ_plain < 1) ctx->occ_plain = 1;
    if (ctx->occ_subst < 1) ctx->occ_subst = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_line, align_lines_kernel<true, 0>, kWarpsPerBlock * 32, 0);
    if (ctx->occ_line < 1) ctx->occ_line = 1;
    {
        int occ_long = 0, coop = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_long, align_long_kernel<true, 0>, 32, 0);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        ctx->long_capacity = coop ? std::max(1, occ_long) * ctx->sm_count : 0;
        cudaGetLastError();
    }
    *out = ctx;
    return TANW_OK;
}

int tanw_destroy(tanw_ctx *ctx)
{
    if (!ctx) return TANW_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = { &ctx->d_sym, &ctx->d_pairs, &ctx->d_order, &ctx->d_counter, &ctx->d_arena,
                       &ctx->d_bnd, &ctx->d_ops, &ctx->d_len, &ctx->d_scores, &ctx->d_subst, &ctx->d_prog, &ctx->d_quads, &ctx->d_chain, &ctx->d_ck };
    for (auto b : bufs) b->release();
    cudaEvent_t evs[] = { ctx->ev_h2d0, ctx->ev_h2d1, ctx->ev_k0, ctx->ev_k1, ctx->ev_d2h0, ctx->ev_d2h1 };
    for (auto ev : evs)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TANW_OK;
}

int tanw_set_arena_limit(tanw_ctx *ctx, int64_t bytes)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (bytes < 0) return fail(ctx, TANW_E_INVALID, "arena limit must be >= 0");
    ctx->arena_limit = bytes;
    return TANW_OK;
}

int tanw_set_long_threshold(tanw_ctx *ctx, int64_t cells)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (cells < 1) return fail(ctx, TANW_E_INVALID, "long-pair threshold must be >= 1 cell");
    ctx->long_cells = cells;
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

int tanw_set_line_kernel(tanw_ctx *ctx, int enabled)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    ctx->use_lines = enabled != 0;
    return TANW_OK;
}

int tanw_stream_handle(tanw_ctx *ctx, uint64_t *out)
{
    if (!ctx || !out) return fail(ctx, TANW_E_INVALID, "NULL argument");
    *out = (uint64_t)(uintptr_t)ctx->stream;
    return TANW_OK;
}

int tanw_sync(tanw_ctx *ctx)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    TANW_CUDA(ctx, cudaSetDevice(ctx->device));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TANW_OK;
}

static int prepare_impl(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                        const int64_t *t_off, const int32_t *n, const int64_t *o_off,
                        const int32_t *m, int64_t n_pairs, const tanw_scoring *sc);

// No exception may cross the C ABI: host allocations (std::vector, pinned tables) can throw.
int tanw_batch_prepare(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                       const int64_t *t_off, const int32_t *n, const int64_t *o_off,
                       const int32_t *m, int64_t n_pairs, const tanw_scoring *sc)
{
    try {
        return prepare_impl(ctx, symbols, symbols_len, t_off, n, o_off, m, n_pairs, sc);
    } catch (const std::bad_alloc &) {
        return fail(ctx, TANW_E_NOMEM, "out of host memory while building the batch tables");
    } catch (...) {
        return fail(ctx, TANW_E_INVALID, "unexpected exception in tanw_batch_prepare");
    }
}

static int prepare_impl(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                        const int64_t *t_off, const int32_t *n, const int64_t *o_off,
                        const int32_t *m, int64_t n_pairs, const tanw_scoring *sc)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    HostTimer host_timer;
    static const bool trace_phases = getenv("TANW_DEBUG_PREP") != nullptr;      // host-side tuning only
    auto mark = [&](const char *what) {
        if (trace_phases) fprintf(stderr, "[tanw prepare] %-28s %.3f ms\n", what, host_timer.ms());
    };
    ctx->prepared = false;
    ctx->ran = false;
    if (n_pairs < 0 || symbols_len < 0) return fail(ctx, TANW_E_INVALID, "negative size");
    if (n_pairs > 0 && (!t_off || !n || !o_off || !m)) return fail(ctx, TANW_E_INVALID, "NULL pair table");
    if (symbols_len > 0 && !symbols) return fail(ctx, TANW_E_INVALID, "symbols is NULL");
    if (!sc) return fail(ctx, TANW_E_INVALID, "scoring is NULL");
    if (n_pairs > 0x7fffffff) return fail(ctx, TANW_E_INVALID, "more than 2^31-1 pairs in one batch");
    const int sb = ctx->sym_bytes;
    const int max_k = sb == 1 ? 256 : kMaxWideSubstK;
    if (sc->subst && (sc->subst_k < 1 || sc->subst_k > max_k))
        return fail(ctx, TANW_E_INVALID, "subst_k must be in 1..%d", max_k);

    // ---- start the symbol upload first: it overlaps the host-side table building below (the
    // copy is asynchronous when the caller's buffer is pinned) -------------------------------
    TANW_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->d_sym.reserve((size_t)symbols_len * (size_t)sb + 16) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, TANW_E_NOMEM, "device allocation failed (symbols, %lld bytes)", (long long)symbols_len);
    }
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_h2d0, ctx->stream));
    if (symbols_len > 0)
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_sym.p, symbols, (size_t)symbols_len * (size_t)sb,
                                       cudaMemcpyHostToDevice, ctx->stream));

    mark("symbol upload issued");
    // ---- pair table, canonical op layout, size statistics --------------------------------
    ctx->h_pairs.resize((size_t)n_pairs);
    ctx->h_ops_off.resize((size_t)n_pairs);
    ctx->h_long.clear();
    ctx->h_long_geo.clear();
    ctx->h_line.clear();
    ctx->h_line_key.clear();
    ctx->h_line_count.assign((size_t)kLineKeys + 1, 0);
    int64_t limit = ctx->arena_limit;
    if (limit == 0) limit = ctx->total_mem / 10 * 4;     // no cudaMemGetInfo on the per-batch path
    int64_t max_line_slot = 0, max_ck = 0;
    int64_t ops_total = 0, cells = 0, ptr_total = 0, max_nm = 0, max_slot = 0, max_long = 0, max_long_bnd = 0;
    int max_n = 0, max_long_pass = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        const int64_t np = n[p], mp = m[p];
        if (np < 0 || mp < 0) return fail(ctx, TANW_E_INVALID, "pair %lld: negative length", (long long)p);
        if (t_off[p] < 0 || o_off[p] < 0 || t_off[p] + np > symbols_len || o_off[p] + mp > symbols_len)
            return fail(ctx, TANW_E_INVALID, "pair %lld: offsets outside the symbol buffer", (long long)p);
        PairDesc &pd = ctx->h_pairs[(size_t)p];
        pd.t_off = t_off[p]; pd.o_off = o_off[p]; pd.n = (int)np; pd.m = (int)mp;
        pd.ops_off = ops_total;
        ctx->h_ops_off[(size_t)p] = ops_total;
        ops_total += np + mp;
        cells += np * mp;
        ptr_total += np * mp;
        max_nm = std::max(max_nm, np + mp);
        // a batch of one or two pages (the drop-in single call) would occupy one or two warps:
        // spread each page over its stripes instead (latency 1.4 ms -> ~0.5 ms per page)
        const bool tiny_batch = n_pairs <= 2 && mp > kLineMaxM && np * mp >= (int64_t(1) << 16);
        // a page whose pointer block does not fit one warp's share of the arena goes to the
        // chained-pass path too, which can cut it into row bands
        const bool oversize = np > 0 && mp > 0 && (ptr_bytes((int)np, (int)mp) + 255) / 256 * 256 * kWarpsPerBlock > limit &&
                              !(sb == 1 && ctx->use_lines && mp <= kLineMaxM && np <= kLineMaxN);
        if (sb == 2) {
            // 16-bit symbol codes: the page kernel only (one warp per pair, any size the arena holds)
            if (oversize)
                return fail(ctx, TANW_E_NOMEM, "pair %lld: %lld bytes of traceback pointers per warp exceed the arena "
                            "limit (pairs with 16-bit symbols have no striped path)", (long long)p,
                            (long long)ptr_bytes((int)np, (int)mp));
            max_slot = std::max<int64_t>(max_slot, ptr_bytes((int)np, (int)mp));
            max_n = std::max(max_n, (int)np);
        } else if ((np * mp >= ctx->long_cells || tiny_batch || oversize) && ctx->long_capacity > 0) {
            // whole-manuscript pair: one warp per column stripe, all stripes resident at once
            const int cf = long_stripe_c(ctx, (int)mp);
            const int64_t npass = (mp + 32 * cf - 1) / (32 * cf);
            const int rows = long_band_rows(ctx, (int)np, (int)mp, cf, limit);
            if (rows <= 0)
                return fail(ctx, TANW_E_NOMEM, "pair %lld: not even %d rows of traceback pointers (%lld bytes each) "
                            "fit the arena limit of %lld bytes", (long long)p, kMinBandRows,
                            (long long)(npass * 32 * cf), (long long)limit);
            ctx->h_long.push_back((int)p);
            ctx->h_long_geo.push_back(make_int2(cf, rows));
            const int64_t bands = (np + rows - 1) / rows;
            max_long = std::max<int64_t>(max_long, ptr_bytes((int)std::min<int64_t>(rows, np), (int)mp, cf));
            max_long_bnd = std::max(max_long_bnd, (npass + 1) * (std::min<int64_t>(rows, np) + 4));
            max_long_pass = std::max<int>(max_long_pass, (int)npass);
            max_ck = std::max(max_ck, (bands - 1) * 3 * mp);
        } else if (ctx->use_lines && mp <= kLineMaxM && np <= kLineMaxN) {
            // short pair: 8 lanes per pair, four pairs per warp.  Sort key for the quads, counted
            // here while the pair is at hand: descending (strip-width class, n); cell-less pairs last
            ctx->h_line.push_back((int)p);
            max_line_slot = std::max<int64_t>(max_line_slot, line_ptr_bytes((int)np, (int)mp));
            const bool act = np > 0 && mp > 0;
            const int key = kLineKeys - 1 - ((act ? line_c((int)mp) / 4 - 1 : 0) * (kLineMaxN + 1) + (act ? (int)np : 0));
            ctx->h_line_key.push_back(key);
            ++ctx->h_line_count[(size_t)key + 1];
        } else {
            max_slot = std::max<int64_t>(max_slot, ptr_bytes((int)np, (int)mp));
            max_n = std::max(max_n, (int)np);
        }
    }
    if (!scoring_in_range(sc, max_nm))
        return fail(ctx, TANW_E_RANGE,
                    "scores do not fit the int32 fixed-point representation: (n+m+2)*max|param| must be < 2^22");
    if (sc->subst) {
        int maxsym = 0;
        if (sb == 1)
            for (int64_t i = 0; i < symbols_len; ++i) maxsym = std::max<int>(maxsym, symbols[i]);
        else
            for (int64_t i = 0; i < symbols_len; ++i)
                maxsym = std::max<int>(maxsym, reinterpret_cast<const uint16_t *>(symbols)[i]);
        if (symbols_len > 0 && maxsym >= sc->subst_k)
            return fail(ctx, TANW_E_INVALID, "symbol code %d >= subst_k %d", maxsym, sc->subst_k);
    }

    mark("pair table + range check");
    // ---- work order: largest pairs first (greedy longest-processing-time) ------------------
    ctx->h_order.clear();
    ctx->h_order.reserve((size_t)n_pairs);
    {
        size_t li = 0, si = 0;
        for (int64_t p = 0; p < n_pairs; ++p) {
            if (li < ctx->h_long.size() && ctx->h_long[li] == (int)p) { ++li; continue; }
            if (si < ctx->h_line.size() && ctx->h_line[si] == (int)p) { ++si; continue; }
            ctx->h_order.push_back((int)p);
        }
    }
    // ---- quads for the line kernel: equal strip width, similar height --------------------------
    ctx->h_quads.clear();
    if (!ctx->h_line.empty()) {
        // counting sort on (strip-width class, n) descending; keys and counts come from the loop above
        const size_t nl = ctx->h_line.size();
        std::vector<int> &count = ctx->h_line_count, &sorted = ctx->h_line_sorted, &skey = ctx->h_line_skey;
        const std::vector<int> &key = ctx->h_line_key;
        for (int b = 1; b <= kLineKeys; ++b) count[(size_t)b] += count[(size_t)b - 1];
        sorted.resize(nl);
        skey.resize(nl);
        for (size_t i = 0; i < nl; ++i) {
            const int at = count[(size_t)key[i]]++;
            sorted[(size_t)at] = ctx->h_line[i];
            skey[(size_t)at] = key[i];
        }
        ctx->h_quads.reserve(nl / 4 + 8);
        size_t i = 0;
        while (i < nl) {
            const int cls = (kLineKeys - 1 - skey[i]) / (kLineMaxN + 1);
            int q[4] = { -1, -1, -1, -1 };
            int c = 0;
            while (c < 4 && i < nl && (kLineKeys - 1 - skey[i]) / (kLineMaxN + 1) == cls) q[c++] = sorted[i++];
            ctx->h_quads.push_back(make_int4(q[0], q[1], q[2], q[3]));
        }
    }
    mark("line quads");
    const int64_t n_quads = (int64_t)ctx->h_quads.size();
    const int64_t n_batch = (int64_t)ctx->h_order.size();
    if (n_batch > 1) {
        // Largest pairs first (greedy longest-processing-time) only needs an approximate order:
        // one stable counting sort on n*m quantised to 16 bits, O(pairs), instead of a
        // comparison sort (which cost 8 ms of host time per 125k line pairs).
        const pinned_vector<PairDesc> &hp = ctx->h_pairs;
        int64_t max_cells = 1;
        for (int p : ctx->h_order) max_cells = std::max(max_cells, (int64_t)hp[(size_t)p].n * hp[(size_t)p].m);
        int shift = 0;
        while ((max_cells >> shift) >= 65536) ++shift;
        std::vector<int> count(65537, 0);
        for (int p : ctx->h_order) {
            const int64_t c = (int64_t)hp[(size_t)p].n * hp[(size_t)p].m;
            ++count[(size_t)(65535 - (c >> shift)) + 1];
        }
        for (size_t b = 1; b <= 65536; ++b) count[b] += count[b - 1];
        pinned_vector<int> &sorted = ctx->h_order_sorted;
        sorted.resize((size_t)n_batch);
        for (int p : ctx->h_order) {
            const int64_t c = (int64_t)hp[(size_t)p].n * hp[(size_t)p].m;
            sorted[(size_t)count[(size_t)(65535 - (c >> shift))]++] = p;
        }
        ctx->h_order.swap(sorted);
    }

    mark("page order");
    // ---- kernel parameters ------------------------------------------------------------------
    fill_kparams(ctx->kp, sc);
    ctx->use_subst = sc->subst != nullptr;
    ctx->opens_nonpositive = sc->gap_open_x <= 0 && sc->gap_open_y <= 0;
    ctx->max_nm = max_nm;

    TANW_CUDA(ctx, cudaSetDevice(ctx->device));

    // ---- launch geometry and scratch ----------------------------------------------------------
    const int occ = ctx->use_subst ? ctx->occ_subst : ctx->occ_plain;
    int grid = ctx->sm_count * occ;
    const int64_t need_blocks = (n_batch + kWarpsPerBlock - 1) / kWarpsPerBlock;
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
    const int64_t slots = (int64_t)grid * kWarpsPerBlock;
    int line_grid = ctx->sm_count * ctx->occ_line;
    if ((n_quads + kWarpsPerBlock - 1) / kWarpsPerBlock < line_grid)
        line_grid = (int)std::max<int64_t>((n_quads + kWarpsPerBlock - 1) / kWarpsPerBlock, 1);
    const int64_t line_slot = (max_line_slot + 255) / 256 * 256;
    ctx->line_grid = line_grid;
    const int64_t line_arena = n_quads ? (int64_t)line_grid * kWarpsPerBlock * 4 * line_slot : 0;
    const int bnd_rows = max_n + 4;      // bnd[1..n] plus the prefetch overrun

    // ---- device buffers and uploads -------------------------------------------------------
    if (ctx->d_pairs.reserve(sizeof(PairDesc) * (size_t)std::max<int64_t>(n_pairs, 1)) != cudaSuccess ||
        ctx->d_order.reserve(sizeof(int) * (size_t)std::max<int64_t>(n_pairs, 1)) != cudaSuccess ||
        ctx->d_counter.reserve(256) != cudaSuccess ||
        ctx->d_arena.reserve((size_t)std::max<int64_t>(std::max(std::max(slots * slot_bytes, max_long), line_arena), 256)) != cudaSuccess ||
        ctx->d_quads.reserve(sizeof(int4) * (size_t)std::max<int64_t>(n_quads, 1)) != cudaSuccess ||
        ctx->d_bnd.reserve(sizeof(int2) * (size_t)std::max<int64_t>(slots * bnd_rows, 1)) != cudaSuccess ||
        reserve_zeroed(ctx, ctx->d_chain, sizeof(int4) * (size_t)max_long_bnd) != cudaSuccess ||
        ctx->d_ck.reserve(sizeof(int) * (size_t)(4 + max_ck)) != cudaSuccess ||
        ctx->d_ops.reserve((size_t)ops_total + 64) != cudaSuccess ||
        ctx->d_len.reserve(sizeof(int) * (size_t)std::max<int64_t>(n_pairs, 1)) != cudaSuccess ||
        ctx->d_scores.reserve(sizeof(int) * 3 * (size_t)std::max<int64_t>(n_pairs, 1)) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, TANW_E_NOMEM, "device allocation failed (arena %lld bytes)", (long long)(slots * slot_bytes));
    }
    mark("device buffers");
    int64_t h2d = symbols_len * sb;
    ctx->batch_sym_bytes = sb;
    if (n_pairs > 0) {
        TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_pairs.p, ctx->h_pairs.data(), sizeof(PairDesc) * (size_t)n_pairs,
                                       cudaMemcpyHostToDevice, ctx->stream));
        if (n_batch > 0)
            TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.data(), sizeof(int) * (size_t)n_batch,
                                           cudaMemcpyHostToDevice, ctx->stream));
        if (n_quads > 0)
            TANW_CUDA(ctx, cudaMemcpyAsync(ctx->d_quads.p, ctx->h_quads.data(), sizeof(int4) * (size_t)n_quads,
                                           cudaMemcpyHostToDevice, ctx->stream));
        h2d += (int64_t)sizeof(PairDesc) * n_pairs + (int64_t)sizeof(int) * n_batch + (int64_t)sizeof(int4) * n_quads;
    }
    if (ctx->use_subst) {
        int rc = upload_subst(ctx, sc, &h2d);
        if (rc) return rc;
    }
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_h2d1, ctx->stream));
    mark("table uploads issued");

    BatchArgs &a = ctx->args;
    a.sym = (const uint8_t *)ctx->d_sym.p;
    a.pairs = (const PairDesc *)ctx->d_pairs.p;
    a.order = (const int *)ctx->d_order.p;
    a.counter = (unsigned *)ctx->d_counter.p;
    a.n_pairs = (int)n_batch;
    a.ptr_arena = (uint8_t *)ctx->d_arena.p;
    a.slot_bytes = slot_bytes;
    a.bnd_arena = (int2 *)ctx->d_bnd.p;
    a.bnd_rows = bnd_rows;
    a.ops = (uint8_t *)ctx->d_ops.p;
    a.ops_len = (int *)ctx->d_len.p;
    a.scores = (int *)ctx->d_scores.p;

    LineArgs &la = ctx->largs;
    la.sym = a.sym;
    la.pairs = a.pairs;
    la.quads = (const int4 *)ctx->d_quads.p;
    la.counter = (unsigned *)ctx->d_counter.p + 1;
    la.n_quads = (int)n_quads;
    la.ptr_arena = a.ptr_arena;
    la.slot_bytes = line_slot;
    la.ops = a.ops;
    la.ops_len = a.ops_len;
    la.scores = a.scores;

    ctx->n_pairs = n_pairs;
    ctx->ops_total = ops_total;
    memset(&ctx->timing, 0, sizeof ctx->timing);
    ctx->timing.cells = cells;
    ctx->timing.ptr_bytes = ptr_total;
    ctx->timing.h2d_bytes = h2d;
    ctx->timing.host_prepare_ms = host_timer.ms();
    ctx->prepared = true;
    return TANW_OK;
}

int tanw_batch_rescore(tanw_ctx *ctx, const tanw_scoring *sc)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (!ctx->prepared) return fail(ctx, TANW_E_STATE, "tanw_batch_rescore before tanw_batch_prepare");
    if (!sc) return fail(ctx, TANW_E_INVALID, "scoring is NULL");
    if ((sc->subst != nullptr) != ctx->use_subst)
        return fail(ctx, TANW_E_INVALID, "rescore cannot switch between an equality scorer and a table");
    if (sc->subst && sc->subst_k != ctx->kp.subst_k)
        return fail(ctx, TANW_E_INVALID, "rescore needs a table of the same size (K = %d)", ctx->kp.subst_k);
    if (!scoring_in_range(sc, ctx->max_nm))
        return fail(ctx, TANW_E_RANGE,
                    "scores do not fit the int32 fixed-point representation: (n+m+2)*max|param| must be < 2^22");
    TANW_CUDA(ctx, cudaSetDevice(ctx->device));
    fill_kparams(ctx->kp, sc);
    ctx->opens_nonpositive = sc->gap_open_x <= 0 && sc->gap_open_y <= 0;
    if (sc->subst) {
        try {
            int rc = upload_subst(ctx, sc, nullptr);
            if (rc) return rc;
        } catch (const std::bad_alloc &) {
            return fail(ctx, TANW_E_NOMEM, "out of host memory for the substitution table");
        }
    }
    ctx->ran = false;
    return TANW_OK;
}

int tanw_batch_run(tanw_ctx *ctx)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (!ctx->prepared) return fail(ctx, TANW_E_STATE, "tanw_batch_run before tanw_batch_prepare");
    HostTimer host_timer;
    TANW_CUDA(ctx, cudaSetDevice(ctx->device));
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_k0, ctx->stream));
    int launches = 0;
    if (ctx->args.n_pairs > 0 || ctx->largs.n_quads > 0)
        TANW_CUDA(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, 2 * sizeof(unsigned), ctx->stream));
    if (ctx->largs.n_quads > 0) {
        const int threads = kWarpsPerBlock * 32;
        switch (kernel_variant(ctx)) {
        case -1: align_lines_kernel<true, 0><<<ctx->line_grid, threads, 0, ctx->stream>>>(ctx->largs, ctx->kp); break;
        case 2:  align_lines_kernel<false, 2><<<ctx->line_grid, threads, 0, ctx->stream>>>(ctx->largs, ctx->kp); break;
        case 1:  align_lines_kernel<false, 1><<<ctx->line_grid, threads, 0, ctx->stream>>>(ctx->largs, ctx->kp); break;
        default: align_lines_kernel<false, 0><<<ctx->line_grid, threads, 0, ctx->stream>>>(ctx->largs, ctx->kp); break;
        }
        TANW_CUDA(ctx, cudaGetLastError());
        ++launches;
    }
    if (ctx->args.n_pairs > 0) {
        // tabulated scorer; general equality scorer; gap opens <= 0; gap opens <= 0 and
        // gap_extend_y == 0 (the reference's default_sys) -- see Strip in tanw_kernels.cuh
        const int threads = kWarpsPerBlock * 32;
        if (ctx->batch_sym_bytes == 2) {
            switch (kernel_variant(ctx)) {
            case -1: align_pairs_kernel<true, 0, uint16_t><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
            case 2:  align_pairs_kernel<false, 2, uint16_t><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
            case 1:  align_pairs_kernel<false, 1, uint16_t><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
            default: align_pairs_kernel<false, 0, uint16_t><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
            }
        } else
        switch (kernel_variant(ctx)) {
        case -1: align_pairs_kernel<true, 0><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
        case 2:  align_pairs_kernel<false, 2><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
        case 1:  align_pairs_kernel<false, 1><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
        default: align_pairs_kernel<false, 0><<<ctx->grid, threads, 0, ctx->stream>>>(ctx->args, ctx->kp); break;
        }
        TANW_CUDA(ctx, cudaGetLastError());
        ++launches;
    }
    for (size_t i = 0; i < ctx->h_long.size(); ++i) {
        int rc = run_long_pair(ctx, ctx->h_long[i], ctx->h_long_geo[i], &launches);
        if (rc) return rc;
    }
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_k1, ctx->stream));
    ctx->timing.kernel_launches = launches;
    ctx->timing.host_run_ms = host_timer.ms();
    ctx->ran = true;
    return TANW_OK;
}

int tanw_batch_fetch(tanw_ctx *ctx, uint8_t *ops, const int64_t *ops_off, int64_t ops_capacity,
                     int32_t *ops_len, int32_t *scores)
{
    if (!ctx) return fail(nullptr, TANW_E_INVALID, "ctx is NULL");
    if (!ctx->ran) return fail(ctx, TANW_E_STATE, "tanw_batch_fetch before tanw_batch_run");
    HostTimer host_timer;
    const int64_t P = ctx->n_pairs;
    if (P > 0 && (!ops_off || !ops_len)) return fail(ctx, TANW_E_INVALID, "NULL output table");
    if (ctx->ops_total > 0 && !ops) return fail(ctx, TANW_E_INVALID, "ops is NULL");
    // the common case first: the caller uses the canonical layout (prefix sums of n+m)
    bool canonical = P == 0 || (memcmp(ops_off, ctx->h_ops_off.data(), sizeof(int64_t) * (size_t)P) == 0 &&
                                ctx->ops_total <= ops_capacity);
    if (!canonical) {
        for (int64_t p = 0; p < P; ++p) {
            const int64_t cap = (int64_t)ctx->h_pairs[(size_t)p].n + ctx->h_pairs[(size_t)p].m;
            if (ops_off[p] < 0 || ops_off[p] + cap > ops_capacity)
                return fail(ctx, TANW_E_INVALID, "pair %lld: op buffer too small (needs n+m = %lld bytes at offset %lld)",
                            (long long)p, (long long)cap, (long long)ops_off[p]);
        }
    }
    TANW_CUDA(ctx, cudaSetDevice(ctx->device));
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_d2h0, ctx->stream));
    int64_t d2h = 0;
    uint8_t *dst = ops;
    if (!canonical) {
        try {
            ctx->h_stage.resize((size_t)ctx->ops_total);
        } catch (const std::bad_alloc &) {
            return fail(ctx, TANW_E_NOMEM, "out of host memory for the op staging buffer");
        }
        dst = ctx->h_stage.data();
    }
    if (ctx->ops_total > 0) {
        TANW_CUDA(ctx, cudaMemcpyAsync(dst, ctx->d_ops.p, (size_t)ctx->ops_total, cudaMemcpyDeviceToHost, ctx->stream));
        d2h += ctx->ops_total;
    }
    if (P > 0) {
        TANW_CUDA(ctx, cudaMemcpyAsync(ops_len, ctx->d_len.p, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
        d2h += (int64_t)sizeof(int) * P;
        if (scores) {
            TANW_CUDA(ctx, cudaMemcpyAsync(scores, ctx->d_scores.p, sizeof(int) * 3 * (size_t)P,
                                           cudaMemcpyDeviceToHost, ctx->stream));
            d2h += (int64_t)sizeof(int) * 3 * P;
        }
    }
    TANW_CUDA(ctx, cudaEventRecord(ctx->ev_d2h1, ctx->stream));
    TANW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!canonical)
        for (int64_t p = 0; p < P; ++p)
            memcpy(ops + ops_off[p], ctx->h_stage.data() + ctx->h_ops_off[(size_t)p], (size_t)ops_len[p]);
    ctx->timing.d2h_bytes = d2h;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_h2d0, ctx->ev_h2d1) == cudaSuccess) ctx->timing.h2d_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess) ctx->timing.kernel_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev_d2h0, ctx->ev_d2h1) == cudaSuccess) ctx->timing.d2h_ms = ms;
    cudaGetLastError();
    ctx->timing.host_fetch_ms = host_timer.ms();
    return TANW_OK;
}

int tanw_align_batch(tanw_ctx *ctx, const uint8_t *symbols, int64_t symbols_len,
                     const int64_t *t_off, const int32_t *n, const int64_t *o_off, const int32_t *m,
                     int64_t n_pairs, const tanw_scoring *scoring, uint8_t *ops, const int64_t *ops_off,
                     int64_t ops_capacity, int32_t *ops_len, int32_t *scores)
{
    int rc = tanw_batch_prepare(ctx, symbols, symbols_len, t_off, n, o_off, m, n_pairs, scoring);
    if (rc) return rc;
    rc = tanw_batch_run(ctx);
    if (rc) return rc;
    return tanw_batch_fetch(ctx, ops, ops_off, ops_capacity, ops_len, scores);
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
    TANW_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->prepared = ctx->ran = false;                     // borrows the score / counter buffers of the batch
    if (ctx->d_counter.reserve(256) != cudaSuccess || ctx->d_scores.reserve(4096 * sizeof(int)) != cudaSuccess)
        return fail(ctx, TANW_E_NOMEM, "device allocation failed");
    const int iters = 1 << 13, blocks = ctx->sm_count * 8, threads = 256;
    const int *src = (const int *)ctx->d_scores.p;       // any initialised words will do
    TANW_CUDA(ctx, cudaMemsetAsync(ctx->d_scores.p, 1, 4096 * sizeof(int), ctx->stream));
    struct EventPair {                                    // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } ev;
    TANW_CUDA(ctx, cudaEventCreate(&ev.a));
    TANW_CUDA(ctx, cudaEventCreate(&ev.b));
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        TANW_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        int *sink = (int *)ctx->d_counter.p + 8;
        if (which == 0)      int32_peak_kernel<0><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink, 3, 5);
        else if (which == 1) int32_peak_kernel<1><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink, 3, 5);
        else if (which == 2) int32_peak_kernel<2><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink, 3, 5);
        else                 int32_peak_kernel<3><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink, 3, 5);
        TANW_CUDA(ctx, cudaGetLastError());
        TANW_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
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
