/*
 * tanw.h -- C ABI of libtanw.so: batched affine-gap Needleman-Wunsch (three-matrix Gotoh
 * form with the reference's exact tie-breaks) on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of DDMAL/text_alignment:
 *     textSeqCompare.perform_alignment(transcript, ocr, scoring_system, verbose)
 *     /root/reference/textSeqCompare.py:13-177, called from alignToOCR.py:273-274.
 * The Python module text_alignment_b200/textSeqCompare.py keeps that signature and calls the
 * functions below through ctypes.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, a TANW_E_* code otherwise; the message is
 *     available from tanw_last_error(ctx) (or tanw_last_error(NULL) for create failures);
 *   - the caller owns every host buffer and keeps it alive for the duration of the call;
 *     the library owns device memory, streams and staging inside the context.  One exception
 *     in the three-phase form: tanw_batch_prepare returns while `symbols` is still being copied
 *     (that is what lets a double-buffering caller overlap it with another context's kernel), so
 *     `symbols` must stay valid and unmodified until tanw_batch_fetch or tanw_sync has returned;
 *     the pair arrays and the scoring system(s) are consumed before prepare returns;
 *   - a context is single-caller; distinct contexts (one per GPU) may be driven from
 *     distinct host threads concurrently (that is the multi-GPU model);
 *   - every entry runs on the context's device and restores the caller's current CUDA device
 *     before it returns;
 *   - there is no CPU fallback: without an sm_100 device tanw_create fails;
 *   - scores are int32 fixed point on the device: a batch is refused with TANW_E_RANGE unless
 *     (max(n+m) + 2) * max|parameter| < 2^22, where the parameters are match, mismatch,
 *     gap_open+gap_extend, gap_extend (both axes), boundary_gap and every table entry -- e.g.
 *     default_sys (max 10) allows n+m up to 419 428, a callable returning +-100 up to 41 941.
 *
 * Data model (replaces the Python lists of textSeqCompare.py:13, :21-22)
 *   symbols : uint8 codes (uint16 after tanw_set_symbol_bytes(ctx, 2)), all sequences of the
 *             batch concatenated; equal codes <=> elements that compare equal in Python (the
 *             shim interns them per pair);
 *   pair p  : transcript = symbols[t_off[p] .. t_off[p]+n[p]), OCR = symbols[o_off[p] .. +m[p]);
 *   ops     : per pair, the alignment columns left to right, one byte per column:
 *             0 = (T[x], O[y])  diagonal          (textSeqCompare.py:115-125)
 *             1 = (T[x], '_')   gap in the OCR     (textSeqCompare.py:128-135, :160-164)
 *             2 = ('_', O[y])   gap in transcript  (textSeqCompare.py:138-145, :154-158)
 *             written at ops[ops_off[p] .. ops_off[p]+ops_len[p]); capacity n[p]+m[p];
 *   scores  : (M, X, Y)[n][m] per pair as int32 (the reference keeps them in mat / x_mat /
 *             y_mat, textSeqCompare.py:45-47, and returns none); TANW_NEG_INF stands for the
 *             reference's -1e100 sentinel (:55, :60).
 */
#ifndef TANW_H
#define TANW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TANW_OK            0
#define TANW_E_INVALID     1   /* bad argument (null pointer, negative size, ...)            */
#define TANW_E_RANGE       2   /* scores would not fit the int32 fixed-point representation   */
#define TANW_E_CUDA        3   /* a CUDA runtime call failed                                  */
#define TANW_E_NODEVICE    4   /* no sm_100 device / device index out of range                */
#define TANW_E_NOMEM       5   /* host or device allocation failed                            */
#define TANW_E_STATE       6   /* call order violated (run before prepare, ...)               */
#define TANW_E_INTERNAL    7   /* a device assertion failed (TANW_CHECKED builds only)        */

#define TANW_NEG_INF       (-1073741824)   /* int32 stand-in for -1e100 in score outputs */

typedef struct tanw_ctx tanw_ctx;

/* Scoring system after the parsing of textSeqCompare.py:24-42.
 * subst == NULL : score(a,b) = (a == b) ? match : mismatch      (:31-32, :36-37)
 * subst != NULL : score(a,b) = subst[a*subst_k + b], the tabulated user callable (:27-29);
 *                 every symbol code in the batch must be < subst_k (<= 256; <= 2048 with 16-bit codes).
 * boundary_gap  : the module-level constant gap_extend that initialises row 0 / column 0
 *                 (:9, :54-59) -- NOT the call's gap parameters. */
typedef struct tanw_scoring {
    int32_t match;
    int32_t mismatch;
    int32_t gap_open_x;
    int32_t gap_open_y;
    int32_t gap_extend_x;
    int32_t gap_extend_y;
    int32_t boundary_gap;
    int32_t subst_k;
    const int32_t *subst;
} tanw_scoring;

typedef struct tanw_device_info {
    char     name[128];
    int32_t  cc_major, cc_minor;
    int32_t  sm_count;
    int32_t  clock_khz;
    int64_t  total_mem_bytes;
    int64_t  free_mem_bytes;
} tanw_device_info;

/* CUDA-event timings of the most recent batch on this context, milliseconds. */
typedef struct tanw_timing {
    float   h2d_ms;        /* prepare: host -> device copies of symbols + pair table           */
    float   kernel_ms;     /* run: fill + traceback kernels                                    */
    float   d2h_ms;        /* fetch: device -> host copies of ops, lengths, scores             */
    int32_t kernel_launches;  /* kernels launched by run                                       */
    int64_t cells;         /* sum of n*m over the batch                                        */
    int64_t ptr_bytes;     /* traceback-pointer bytes the fill kernel writes (algorithmic)     */
    int64_t h2d_bytes, d2h_bytes;
    /* host wall-clock spent inside the three phases (includes the waits on the device) */
    float   host_prepare_ms, host_run_ms, host_fetch_ms;
    int32_t chunks;           /* chunks the batch was pipelined in (1 in the three-phase form)  */
    int32_t table_launches;   /* small kernels launched by prepare to build the batch tables    */
} tanw_timing;

/* ---- library / device queries ------------------------------------------------------------ */
int  tanw_version(void);                                  /* 100*major + minor                 */
int  tanw_device_count(int *count);
int  tanw_device_query(int device, tanw_device_info *out);
const char *tanw_last_error(const tanw_ctx *ctx);         /* ctx may be NULL                    */

/* ---- context -------------------------------------------------------------------------------- */
int  tanw_create(int device, tanw_ctx **out);
int  tanw_destroy(tanw_ctx *ctx);
/* Upper bound for the traceback-pointer arena in bytes (0 = default: 40% of device memory). */
int  tanw_set_arena_limit(tanw_ctx *ctx, int64_t bytes);
/* Pairs with n*m >= cells take the chained-pass path: one warp per 512-column stripe, all
 * stripes resident at once (cooperative launch), so that a single whole-manuscript pair
 * (BASELINE config 5, 100k x 80k) uses the whole GPU.  Default 2^26.  Results are identical
 * on both paths; the threshold only moves work between them. */
int  tanw_set_long_threshold(tanw_ctx *ctx, int64_t cells);
/* A chained-pass pair whose traceback pointers (1 byte per cell) exceed the arena limit is cut
 * into bands of rows: one forward fill that keeps only the per-column state at every band edge,
 * then, bottom band first, a second fill of each band that stores its pointers and the
 * traceback through it (about twice the fill work, memory bounded by one band).  Pages too large
 * for one warp's share of the arena take the same route.  rows > 0 forces bands of that height
 * (tests; tuning), 0 = only when needed.  Results are identical either way. */
int  tanw_set_long_band_rows(tanw_ctx *ctx, int rows);
/* Width of a symbol code in bytes: 1 (default) or 2.  With 2, `symbols` in the batch calls points
 * to uint16 codes (pass the array's address), symbols_len / t_off / o_off still count symbols,
 * and subst_k may be up to 2048.  For pairs with more than 256 distinct elements (the reference
 * accepts any hashable, textSeqCompare.py:13-22).  Such batches run on the page kernel (one warp
 * per pair, general recurrences); a pair whose pointers exceed one warp's share of the arena takes
 * the chained-stripe path and its row bands like any other. */
int  tanw_set_symbol_bytes(tanw_ctx *ctx, int bytes);
/* Pairs with m <= 128 and n <= 4096 are aligned by the line kernels (8 lanes per pair; BASELINE
 * config 3): eight per warp, two per 32-bit register, when their scores fit 16 bits (gap opens
 * <= 0, equality scorer with match >= mismatch, (2n + 132) * max|param| <= 8000), else four per
 * warp in int32.  mode 0 sends them through the page kernel instead, 2 through the int32 line
 * kernel only (same results on every route; used by the tests to compare them).  Default: 1. */
int  tanw_set_line_kernel(tanw_ctx *ctx, int mode);

/* Packed op strings: an op is 0, 1 or 2, so with enabled != 0 the batch calls deliver FOUR OPS PER
 * BYTE -- op q of pair p in bits 2*(q & 3) of byte (ops_off_canonical[p] >> 2) + p + (q >> 2) of
 * `ops`, where ops_off_canonical is the prefix sum of n+m (every pair starts on a fresh byte; no
 * other layout is offered, `ops_off` is ignored) and `ops_capacity` must be at least
 * sum(n+m)/4 + n_pairs + 1.  A quarter of the bytes cross PCIe on the way back, which is what
 * bounds short-line batches on a box with eight GPUs.  Same alignments, another wire format. */
int  tanw_set_packed_ops(tanw_ctx *ctx, int enabled);

/* ---- one-call batch alignment: the entry the reference's call site maps to ------------------
 * Replaces N calls of textSeqCompare.perform_alignment (textSeqCompare.py:13) -- copies the
 * inputs to the device, runs fill + traceback, copies ops / lengths / scores back.
 * scores may be NULL.  ops_off[p] must leave n[p]+m[p] bytes for pair p. */
int  tanw_align_batch(tanw_ctx *ctx,
                      const uint8_t *symbols, int64_t symbols_len,
                      const int64_t *t_off, const int32_t *n,
                      const int64_t *o_off, const int32_t *m,
                      int64_t n_pairs, const tanw_scoring *scoring,
                      uint8_t *ops, const int64_t *ops_off, int64_t ops_capacity,
                      int32_t *ops_len, int32_t *scores);

/* The same with a scoring system per pair: pair p is aligned under scorings[scoring_idx[p]].  This
 * is the reference's parameter sweep (evaluate_text_alignment.py:134-194: 729 integer scoring
 * vectors x 3 pages = 2187 independent alignments) as ONE batch: the pages' symbols are uploaded
 * once and several pairs may name the same t_off / o_off.  Equality scorers only (subst == NULL),
 * 8-bit symbol codes. */
int  tanw_align_batch_multi(tanw_ctx *ctx,
                            const uint8_t *symbols, int64_t symbols_len,
                            const int64_t *t_off, const int32_t *n,
                            const int64_t *o_off, const int32_t *m,
                            int64_t n_pairs, const tanw_scoring *scorings, int32_t n_scorings,
                            const int32_t *scoring_idx,
                            uint8_t *ops, const int64_t *ops_off, int64_t ops_capacity,
                            int32_t *ops_len, int32_t *scores);

/* One batch over several devices (pages are independent, alignToOCR.py:273 is one
 * call per page -- the batch is partitioned, there is no exchange step).  Shard d = the pairs
 * bounds[d] .. bounds[d+1]-1 (bounds[0] = 0, bounds[n_ctx] = n_pairs, non-decreasing; an empty
 * shard is allowed) is aligned on ctxs[d] -- contexts of different devices, or several contexts of
 * one device -- by one host thread per shard, and writes its op strings, lengths and scores
 * straight into its slice of the caller's arrays: the gather is the layout itself.  All other
 * arguments as tanw_align_batch for the WHOLE batch.  Returns the first shard's error, with its
 * text in ctxs[0]'s tanw_last_error.  Not with tanw_set_packed_ops contexts.  The caller must
 * not use any of the contexts from another thread during the call. */
int  tanw_align_batch_sharded(tanw_ctx *const *ctxs, int32_t n_ctx, const int64_t *bounds,
                              const uint8_t *symbols, int64_t symbols_len,
                              const int64_t *t_off, const int32_t *n,
                              const int64_t *o_off, const int32_t *m,
                              int64_t n_pairs, const tanw_scoring *scoring,
                              uint8_t *ops, const int64_t *ops_off, int64_t ops_capacity,
                              int32_t *ops_len, int32_t *scores);

/* ---- the same in three phases (bench.py times `run` alone with inputs resident in HBM) ------ */
int  tanw_batch_prepare(tanw_ctx *ctx,
                        const uint8_t *symbols, int64_t symbols_len,
                        const int64_t *t_off, const int32_t *n,
                        const int64_t *o_off, const int32_t *m,
                        int64_t n_pairs, const tanw_scoring *scoring);
int  tanw_batch_prepare_multi(tanw_ctx *ctx,
                              const uint8_t *symbols, int64_t symbols_len,
                              const int64_t *t_off, const int32_t *n,
                              const int64_t *o_off, const int32_t *m,
                              int64_t n_pairs, const tanw_scoring *scorings, int32_t n_scorings,
                              const int32_t *scoring_idx);
int  tanw_batch_run(tanw_ctx *ctx);                       /* asynchronous on the ctx stream     */
/* Replace the scoring system of the prepared batch; the sequences stay resident in HBM: prepare
 * once, then { rescore, run, fetch } per system (one launch each; tanw_align_batch_multi runs a
 * whole sweep as one launch).  An equality scorer may be replaced by another equality scorer, a
 * table by a table of the same K. */
int  tanw_batch_rescore(tanw_ctx *ctx, const tanw_scoring *scoring);
int  tanw_batch_fetch(tanw_ctx *ctx,
                      uint8_t *ops, const int64_t *ops_off, int64_t ops_capacity,
                      int32_t *ops_len, int32_t *scores);
int  tanw_sync(tanw_ctx *ctx);

int  tanw_last_timing(tanw_ctx *ctx, tanw_timing *out);

/* Raw CUDA stream of the context (cudaStream_t as an integer) so that a caller holding
 * device memory of its own (e.g. torch) can order work against it. */
int  tanw_stream_handle(tanw_ctx *ctx, uint64_t *out);

/* ---- measurement helper: dependency-free int32 issue rate of this device -----------------------
 * Measures the roofline denominator SURVEY.md 8(d) asks the builder to measure: warp-lane
 * INSTRUCTIONS per second issued from every SM, 16 independent chains per thread.
 * `which`: 0 = IADD3 (add.s32), 1 = VIMNMX (max.s32/min.s32), 2 = VIADDMNMX (fused add+max),
 *          3 = VIADD + LOP3 alternating: both integer pipes busy (VIADD runs on either), i.e.
 *              the issue ceiling of a kernel that mixes the pipes.
 * Uses the context's scratch buffers: a prepared batch must be prepared again afterwards. */
int  tanw_measure_int32_peak(tanw_ctx *ctx, int which, double *lane_ops_per_s);

/* ---- the steps either side of the alignment, on packed arrays (host code, no context) -------------
 * OCRopus .llocs records in, syllable boxes / Rodan JSON out, without building a
 * Python object per character.  All three return TANW_OK or a TANW_E_* code with the message in
 * tanw_consumer_last_error() (thread-local). */
const char *tanw_consumer_last_error(void);

/* latinSyllabification.syllabify_text (latinSyllabification.py:22-109, :170-174) on bytes: the
 * syllables of `text` (words = what split(' ') gives) as [first, one-past-last) character ranges,
 * 2 ints per syllable.  ASCII letters, digits and spaces only: anything else returns TANW_E_STATE
 * and the caller syllabifies in Python.  A word without a vowel -- on which the reference never
 * terminates -- is TANW_E_INVALID.  *n_out = number of syllables (also when `capacity` was too small). */
int  tanw_syllabify_text(const char *text, int64_t text_len, int32_t *bounds, int64_t capacity, int64_t *n_out);

/* One text line of `ocropus-rpred --llocs` output (alignToOCR.py:153-182): UTF-8 records
 * "<character> TAB <x of its RIGHT edge inside the strip> NEWLINE".  A character's box runs from
 * the previous record's x to its own (np.round(x + x_min), half to even) over the strip's full
 * height; records whose character is '~' or empty advance x but are dropped (:171-173).  Writes the
 * code point and (ulx, uly, lrx, lry) of every kept character; *n_out = their number (also when
 * `capacity` was too small: TANW_E_NOMEM, call again). */
int  tanw_parse_llocs(const char *text, int64_t text_len, int32_t x_min, int32_t y_min, int32_t y_max,
                      uint32_t *chars, int32_t *boxes, int64_t capacity, int64_t *n_out);

/* alignToOCR.py:285-324 for many pages at once.  Per page: its op string (as tanw_align_batch
 * returns it), its syllables as [first, one-past-last) transcript indices (syl_bounds, 2 ints per
 * syllable, ascending; page pg owns syllables syl_off[pg] .. syl_off[pg+1]) and the boxes of its
 * OCR characters (4 ints per character; page pg owns characters box_off[pg] .. box_off[pg+1], which
 * must be as many as the op string consumes -- the reference's assertion at :291).  For every
 * syllable: out_has = 0 if no OCR character is aligned to it (:313), else its box in out_boxes
 * (union of the aligned characters' boxes, lowest text line only, :318-324).  Equivalent to the
 * reference's regular-expression search when no transcript character is a regex metacharacter
 * or '_' (the Python wrapper checks and otherwise keeps the regex). */
int  tanw_syllable_boxes(int64_t n_pages, const uint8_t *ops, const int64_t *ops_off, const int32_t *ops_len,
                         const int32_t *syl_bounds, const int64_t *syl_off,
                         const int32_t *boxes, const int64_t *box_off,
                         int32_t *out_boxes, uint8_t *out_has);

/* alignToOCR.to_JSON_dict (:333-351) serialised as json.dumps writes it: syllable s is the UTF-8
 * bytes syl_utf8[syl_text_off[s] .. syl_text_off[s+1]); syllables without a box are left out;
 * `median_line_spacing` is the number already formatted by the caller.  *out_len = bytes needed. */
int  tanw_boxes_to_json(const char *syl_utf8, const int64_t *syl_text_off, int64_t n_syl,
                        const int32_t *syl_boxes, const uint8_t *has_box, const char *median_line_spacing,
                        char *out, int64_t capacity, int64_t *out_len);

#ifdef __cplusplus
}
#endif
#endif /* TANW_H */
