// tanw_consumer.cu -- the steps either side of the alignment, on packed arrays (host code).
//
// The reference turns OCRopus' .llocs records into CharBox objects (alignToOCR.py:153-182),
// inserts gap boxes along the aligned OCR string (:285-292), finds every syllable of the transcript
// in the aligned transcript with a regular expression and unions the boxes aligned to it
// (:297-324), and writes the result as JSON (:333-351).  With the alignment at a microsecond per
// page that per-object Python work (about 12 ms per page) is all that is left of a page's cost.
// Here the same results come from one pass over the op string:
//
//   a syllable's letters s0 .. sk are consecutive transcript characters, so the regular expression
//   s0 _* s1 _* .. sk (searched from the end of the previous syllable) matches exactly the columns
//   from the one holding s0 to the one holding sk -- provided no transcript character is a regular
//   expression metacharacter or the gap symbol, which the caller checks (else it keeps the regex
//   path).  The OCR characters in those columns are a contiguous index range [y_lo, y_hi), found
//   by counting ops; the box is the union over that range, after the "lowest text line wins" rule.
//
// Plain C ABI like the rest of the library (include/tanw.h); no device work, no context.
#include "tanw.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_consumer_error;

int cfail(int code, const char *fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_consumer_error = buf;
    return code;
}

// One UTF-8 code point at p (p < end); returns its length, 0 if the bytes are not valid UTF-8.
int utf8_decode(const unsigned char *p, const unsigned char *end, uint32_t *cp)
{
    const unsigned c = p[0];
    if (c < 0x80) { *cp = c; return 1; }
    int len = (c >= 0xF0 && c < 0xF8) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC0) ? 2 : 0;
    if (c >= 0xF8 || len == 0 || p + len > end) return 0;
    uint32_t v = c & (0xFF >> (len + 1));
    for (int i = 1; i < len; ++i) {
        if ((p[i] & 0xC0) != 0x80) return 0;
        v = (v << 6) | (p[i] & 0x3F);
    }
    *cp = v;
    return len;
}

struct Out {
    char *buf;
    int64_t cap, len;
    void put(const char *s, size_t k)
    {
        if (len + (int64_t)k <= cap) memcpy(buf + len, s, k);
        len += (int64_t)k;
    }
    void put(const char *s) { put(s, strlen(s)); }
    void num(long long v)
    {
        char t[32];
        put(t, (size_t)snprintf(t, sizeof t, "%lld", v));
    }
};

// A string as json.dumps writes it with ensure_ascii=True.
bool json_string(Out &o, const unsigned char *s, const unsigned char *end)
{
    o.put("\"");
    while (s < end) {
        uint32_t cp;
        const int k = utf8_decode(s, end, &cp);
        if (k == 0) return false;
        s += k;
        char t[16];
        switch (cp) {
        case '"':  o.put("\\\""); break;
        case '\\': o.put("\\\\"); break;
        case '\n': o.put("\\n"); break;
        case '\r': o.put("\\r"); break;
        case '\t': o.put("\\t"); break;
        case '\b': o.put("\\b"); break;
        case '\f': o.put("\\f"); break;
        default:
            if (cp < 0x20 || (cp >= 0x80 && cp < 0x10000)) {
                o.put(t, (size_t)snprintf(t, sizeof t, "\\u%04x", cp));
            } else if (cp >= 0x10000) {
                const uint32_t v = cp - 0x10000;
                o.put(t, (size_t)snprintf(t, sizeof t, "\\u%04x", 0xD800 + (v >> 10)));
                o.put(t, (size_t)snprintf(t, sizeof t, "\\u%04x", 0xDC00 + (v & 0x3FF)));
            } else {
                t[0] = (char)cp;
                o.put(t, 1);
            }
        }
    }
    o.put("\"");
    return true;
}

}  // namespace

// ---- Latin syllabification (latinSyllabification.py:22-109, :170-174) on bytes ------------------------
// The reference cuts a word into units -- consonant clusters, then diphthongs, each class in its
// listed order and each occurrence taken left to right as str.split does; what is left falls apart
// into single letters -- marks vowels and diphthongs as seeds, and then, until only seeds remain,
// glues every non-seed unit directly before a seed onto it and after that every non-seed unit
// directly after a seed.  Units are contiguous pieces of the word, so a unit is (start, length, flag).
struct Unit { int start, len; bool done, seed; };   // done: taken out by a pattern (no further cuts)

const char *const kClusters[] = { "qu", "ch", "ph", "fl", "fr", "st", "br", "cr", "cl", "pr", "tr", "ct", "th",
                                  "ae", "au", "ei", "oe", "ui", "ya", "ex", "ix" };
constexpr int kConsonantGroups = 13, kPatterns = 21;
constexpr int kMaxWord = 512;

bool is_vowel(char c) { return c == 'a' || c == 'e' || c == 'i' || c == 'o' || c == 'u' || c == 'y'; }

// Syllables of word[0..len) as lengths; returns their number, -1 if the word has no seed (the
// reference never terminates on such a word), -2 if it is too long for the fixed buffers.
int syllabify_word_bytes(const char *w, int len, int *out_len)
{
    if (len == 0) return 0;
    if (len > kMaxWord) return -2;
    // latinSyllabification.py:30-35
    if (len == 6 && memcmp(w, "euouae", 6) == 0) { out_len[0] = 1; out_len[1] = 1; out_len[2] = 1; out_len[3] = 1; out_len[4] = 2; return 5; }
    if (len == 5 && memcmp(w, "cuius", 5) == 0) { out_len[0] = 2; out_len[1] = 3; return 2; }
    if (len == 4 && memcmp(w, "eius", 4) == 0) { out_len[0] = 1; out_len[1] = 3; return 2; }
    static thread_local Unit a[kMaxWord + 2], b[kMaxWord + 2];
    Unit *cur = a, *nxt = b;
    int n = 1;
    cur[0] = { 0, len, false, false };
    // Units are pieces of the word, so a cluster that occurs nowhere in the word cuts nothing: one
    // look at the word's letter pairs says which of the 21 passes can do anything (usually 0-2).
    struct PairTable {
        int8_t idx[128][128];
        PairTable() { memset(idx, -1, sizeof idx); for (int p = 0; p < kPatterns; ++p) idx[(int)kClusters[p][0]][(int)kClusters[p][1]] = (int8_t)p; }
    };
    static const PairTable table;
    unsigned present = 0;
    for (int q = 0; q + 1 < len; ++q) {
        const unsigned char c0 = (unsigned char)w[q], c1 = (unsigned char)w[q + 1];
        if (c0 < 128 && c1 < 128 && table.idx[c0][c1] >= 0) present |= 1u << table.idx[c0][c1];
    }
    for (int p = 0; p < kPatterns; ++p) {
        if (!(present >> p & 1u)) continue;
        const char c0 = kClusters[p][0], c1 = kClusters[p][1];
        int k = 0;
        for (int u = 0; u < n; ++u) {
            const Unit part = cur[u];
            if (part.done) { nxt[k++] = part; continue; }
            int pos = part.start;
            const int end = part.start + part.len;
            for (int q = pos; q + 1 < end;) {
                if (w[q] == c0 && w[q + 1] == c1) {
                    if (q > pos) nxt[k++] = { pos, q - pos, false, false };
                    nxt[k++] = { q, 2, true, p >= kConsonantGroups };
                    q += 2;
                    pos = q;
                } else {
                    ++q;
                }
            }
            if (end > pos) nxt[k++] = { pos, end - pos, false, false };
        }
        Unit *t = cur; cur = nxt; nxt = t;
        n = k;
    }
    // single letters; vowels and diphthongs are seeds (:66-68)
    int k = 0;
    bool any_seed = false;
    for (int u = 0; u < n; ++u) {
        if (cur[u].done) {
            nxt[k] = cur[u];
            nxt[k].done = false;
            any_seed = any_seed || nxt[k].seed;
            ++k;
        } else {
            for (int q = 0; q < cur[u].len; ++q) {
                const bool v = is_vowel(w[cur[u].start + q]);
                nxt[k++] = { cur[u].start + q, 1, false, v };
                any_seed = any_seed || v;
            }
        }
    }
    { Unit *t = cur; cur = nxt; nxt = t; }
    n = k;
    if (!any_seed) return -1;
    // glue until every unit is a seed (:71-105)
    for (;;) {
        bool all = true;
        for (int u = 0; u < n; ++u) all = all && cur[u].seed;
        if (all) break;
        for (int pass = 0; pass < 2; ++pass) {                  // consonant + seed, then seed + consonant
            k = 0;
            for (int i = 0; i < n;) {
                if (i + 1 < n) {
                    const bool as = cur[i].seed, bs = cur[i + 1].seed;
                    if (pass == 0 ? (bs && !as) : (as && !bs)) {
                        nxt[k++] = { cur[i].start, cur[i].len + cur[i + 1].len, false, true };
                        i += 2;
                        continue;
                    }
                }
                nxt[k++] = cur[i++];
            }
            Unit *t = cur; cur = nxt; nxt = t;
            n = k;
        }
    }
    for (int u = 0; u < n; ++u) out_len[u] = cur[u].len;
    return n;
}

extern "C" {

const char *tanw_consumer_last_error(void) { return g_consumer_error.c_str(); }

int tanw_syllabify_text(const char *text, int64_t text_len, int32_t *bounds, int64_t capacity, int64_t *n_out)
{
    if (!n_out || (text_len > 0 && !text) || text_len < 0 || capacity < 0 || (capacity > 0 && !bounds))
        return cfail(TANW_E_INVALID, "tanw_syllabify_text: bad argument");
    int64_t count = 0;
    static thread_local int lens[kMaxWord + 2];
    for (int64_t i = 0; i <= text_len;) {
        int64_t j = i;
        while (j < text_len && text[j] != ' ') {
            const unsigned char c = (unsigned char)text[j];
            const bool alnum = (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
            if (!alnum) return cfail(TANW_E_STATE, "character 0x%02x at %lld: the byte path takes ASCII letters, digits and spaces", c, (long long)j);
            ++j;
        }
        const int k = syllabify_word_bytes(text + i, (int)std::min<int64_t>(j - i, kMaxWord + 1), lens);
        if (k == -1) {
            std::string word(text + i, (size_t)(j - i));
            return cfail(TANW_E_INVALID, "cannot syllabify '%s': no vowel (the reference loops forever here)", word.c_str());
        }
        if (k == -2) return cfail(TANW_E_STATE, "word of %lld characters at %lld", (long long)(j - i), (long long)i);
        int64_t at = i;
        for (int u = 0; u < k; ++u) {
            if (count < capacity) { bounds[2 * count] = (int32_t)at; bounds[2 * count + 1] = (int32_t)(at + lens[u]); }
            at += lens[u];
            ++count;
        }
        i = j + 1;
    }
    *n_out = count;
    if (count > capacity) return cfail(TANW_E_NOMEM, "%lld syllables, capacity %lld", (long long)count, (long long)capacity);
    return TANW_OK;
}

int tanw_parse_llocs(const char *text, int64_t text_len, int32_t x_min, int32_t y_min, int32_t y_max,
                     uint32_t *chars, int32_t *boxes, int64_t capacity, int64_t *n_out)
{
    if (!n_out || (text_len > 0 && !text) || text_len < 0 || capacity < 0 || (capacity > 0 && (!chars || !boxes)))
        return cfail(TANW_E_INVALID, "tanw_parse_llocs: bad argument");
    const unsigned char *p = (const unsigned char *)text, *end = p + text_len;
    int64_t count = 0, rec = 0;
    long long prev_x = x_min;
    while (p < end) {
        // one record: <character> TAB <x of its right edge> [TAB ...] NEWLINE   (alignToOCR.py:157-169)
        const unsigned char *eol = p;
        while (eol < end && *eol != '\n' && *eol != '\r') ++eol;
        const unsigned char *tab = p;
        while (tab < eol && *tab != '\t') ++tab;
        if (tab == eol)
            return cfail(TANW_E_INVALID, "llocs record %lld has no position field", (long long)rec);
        const unsigned char *num = tab + 1, *num_end = num;
        while (num_end < eol && *num_end != '\t') ++num_end;
        char tmp[64];
        const size_t nl = (size_t)(num_end - num);
        if (nl == 0 || nl >= sizeof tmp)
            return cfail(TANW_E_INVALID, "llocs record %lld: position field is not a number", (long long)rec);
        memcpy(tmp, num, nl);
        tmp[nl] = 0;
        char *stop = nullptr;
        const double x = strtod(tmp, &stop);
        while (stop && (*stop == ' ')) ++stop;
        if (!stop || *stop != 0 || stop == tmp)
            return cfail(TANW_E_INVALID, "llocs record %lld: position field is not a number", (long long)rec);
        const long long cur_x = (long long)std::nearbyint(x + (double)x_min);      // np.round: half to even (:166)
        // the character field without '~' (clean_special_chars, :61-72); '~' alone or nothing: set aside (:171-173)
        uint32_t cp = 0;
        int kept = 0;
        bool only_tilde_or_empty = true;
        for (const unsigned char *q = p; q < tab;) {
            uint32_t c;
            const int k = utf8_decode(q, tab, &c);
            if (k == 0) return cfail(TANW_E_INVALID, "llocs record %lld: invalid UTF-8", (long long)rec);
            q += k;
            if (c != '~') { cp = c; ++kept; }
        }
        only_tilde_or_empty = (tab == p) || (tab - p == 1 && *p == '~');
        if (!only_tilde_or_empty) {
            if (kept != 1)
                return cfail(TANW_E_INVALID, "llocs record %lld: the character field holds %d characters "
                             "(the array path needs exactly one)", (long long)rec, kept);
            if (count < capacity) {
                chars[count] = cp;
                boxes[4 * count + 0] = (int32_t)prev_x;
                boxes[4 * count + 1] = y_min;
                boxes[4 * count + 2] = (int32_t)cur_x;
                boxes[4 * count + 3] = y_max;
            }
            ++count;
        }
        prev_x = cur_x;
        ++rec;
        p = eol;
        if (p < end && *p == '\r') ++p;
        if (p < end && *p == '\n') ++p;
    }
    *n_out = count;
    if (count > capacity) return cfail(TANW_E_NOMEM, "llocs: %lld characters, capacity %lld", (long long)count, (long long)capacity);
    return TANW_OK;
}

int tanw_syllable_boxes(int64_t n_pages, const uint8_t *ops, const int64_t *ops_off, const int32_t *ops_len,
                        const int32_t *syl_bounds, const int64_t *syl_off,
                        const int32_t *boxes, const int64_t *box_off,
                        int32_t *out_boxes, uint8_t *out_has)
{
    if (n_pages < 0) return cfail(TANW_E_INVALID, "tanw_syllable_boxes: negative page count");
    if (n_pages > 0 && (!ops_off || !ops_len || !syl_off || !box_off))
        return cfail(TANW_E_INVALID, "tanw_syllable_boxes: NULL table");
    for (int64_t pg = 0; pg < n_pages; ++pg) {
        const uint8_t *op = ops + ops_off[pg];
        const int64_t L = ops_len[pg];
        const int64_t s0 = syl_off[pg], s1 = syl_off[pg + 1];
        const int32_t *bx = boxes + 4 * box_off[pg];
        const int64_t m = box_off[pg + 1] - box_off[pg];
        int64_t s = s0;
        int64_t x = 0, y = 0, ylo = 0;
        for (int64_t s2 = s0; s2 < s1; ++s2) {
            out_has[s2] = 0;
            if (syl_bounds[2 * s2] >= syl_bounds[2 * s2 + 1] || (s2 > s0 && syl_bounds[2 * s2] < syl_bounds[2 * s2 - 1]))
                return cfail(TANW_E_INVALID, "page %lld: syllable %lld is empty or out of order", (long long)pg, (long long)(s2 - s0));
        }
        for (int64_t c = 0; c < L; ++c) {
            const int o = op[c];
            if (o > 2) return cfail(TANW_E_INVALID, "page %lld: op %d at column %lld", (long long)pg, o, (long long)c);
            if (o != 2) {                                       // a transcript character sits in this column
                if (s < s1 && x == syl_bounds[2 * s]) ylo = y;
                if (s < s1 && x == syl_bounds[2 * s + 1] - 1) {
                    const int64_t yhi = y + (o == 0 ? 1 : 0);
                    if (yhi > m)
                        return cfail(TANW_E_INVALID, "page %lld: all_chars not same length as alignment", (long long)pg);
                    if (yhi > ylo) {
                        // several text lines: keep the lowest (largest uly), alignToOCR.py:318-320
                        int32_t low = bx[4 * ylo + 1];
                        for (int64_t k = ylo + 1; k < yhi; ++k) low = bx[4 * k + 1] > low ? bx[4 * k + 1] : low;
                        bool first = true;
                        int32_t ulx = 0, lrx = 0, lry = 0;
                        for (int64_t k = ylo; k < yhi; ++k) {
                            if (bx[4 * k + 1] != low) continue;
                            if (first || bx[4 * k + 0] < ulx) ulx = bx[4 * k + 0];
                            if (first || bx[4 * k + 2] > lrx) lrx = bx[4 * k + 2];
                            if (first || bx[4 * k + 3] > lry) lry = bx[4 * k + 3];
                            first = false;
                        }
                        out_boxes[4 * s + 0] = ulx; out_boxes[4 * s + 1] = low;
                        out_boxes[4 * s + 2] = lrx; out_boxes[4 * s + 3] = lry;
                        out_has[s] = 1;
                    }
                    ++s;
                }
                ++x;
            }
            if (o != 1) ++y;
        }
        if (y != m)
            return cfail(TANW_E_INVALID, "page %lld: all_chars not same length as alignment: %lld vs %lld OCR characters",
                         (long long)pg, (long long)m, (long long)y);
        if (s != s1)
            return cfail(TANW_E_INVALID, "page %lld: %lld syllables lie beyond the aligned transcript", (long long)pg, (long long)(s1 - s));
    }
    return TANW_OK;
}

int tanw_boxes_to_json(const char *syl_utf8, const int64_t *syl_text_off, int64_t n_syl,
                       const int32_t *syl_boxes, const uint8_t *has_box, const char *median_line_spacing,
                       char *out, int64_t capacity, int64_t *out_len)
{
    if (!out_len || n_syl < 0 || capacity < 0 || !median_line_spacing || (n_syl > 0 && (!syl_utf8 || !syl_text_off || !syl_boxes || !has_box)))
        return cfail(TANW_E_INVALID, "tanw_boxes_to_json: bad argument");
    Out o = { out, out ? capacity : 0, 0 };
    o.put("{\"median_line_spacing\": ");
    o.put(median_line_spacing);
    o.put(", \"syl_boxes\": [");
    bool first = true;
    for (int64_t s = 0; s < n_syl; ++s) {
        if (!has_box[s]) continue;                              // aligned to nothing in the OCR (:313)
        if (!first) o.put(", ");
        first = false;
        o.put("{\"syl\": ");
        if (!json_string(o, (const unsigned char *)syl_utf8 + syl_text_off[s], (const unsigned char *)syl_utf8 + syl_text_off[s + 1]))
            return cfail(TANW_E_INVALID, "syllable %lld is not valid UTF-8", (long long)s);
        o.put(", \"ul\": [");
        o.num(syl_boxes[4 * s + 0]); o.put(", "); o.num(syl_boxes[4 * s + 1]);
        o.put("], \"lr\": [");
        o.num(syl_boxes[4 * s + 2]); o.put(", "); o.num(syl_boxes[4 * s + 3]);
        o.put("]}");
    }
    o.put("]}");
    *out_len = o.len;
    if (o.len > capacity || !out) return cfail(TANW_E_NOMEM, "JSON needs %lld bytes, capacity %lld", (long long)o.len, (long long)capacity);
    return TANW_OK;
}

}  // extern "C"
