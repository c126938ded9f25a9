/* examples/abi_example.c -- using libtanw.so from plain C (no Python, no torch).
 *
 *   gcc -Iinclude examples/abi_example.c -o abi_example -Ltext_alignment_b200 -ltanw \
 *       -Wl,-rpath,$PWD/text_alignment_b200 && ./abi_example
 *
 * Aligns two (transcript, OCR) pairs with the reference's default scoring system
 * [8, -4, -7, -7, -3, 0] and boundary constant gap_extend = -1 (textSeqCompare.py:9-10) and
 * prints the gap-padded sequences exactly as textSeqCompare.perform_alignment would return them. */
#include <stdio.h>
#include <string.h>
#include "tanw.h"

static void show(const char *T, const char *O, const uint8_t *ops, int len)
{
    char tra[256], ocr[256];
    int x = 0, y = 0;
    for (int k = 0; k < len; ++k) {
        tra[k] = ops[k] == 2 ? '_' : T[x++];
        ocr[k] = ops[k] == 1 ? '_' : O[y++];
    }
    tra[len] = ocr[len] = 0;
    printf("%s\n%s\n", tra, ocr);
}

int main(void)
{
    const char *T[2] = { "dominus", "alleluia" }, *O[2] = { "dns", "a l l e l u y a" };
    uint8_t sym[64], ops[64];
    int64_t t_off[2], o_off[2], ops_off[2];
    int32_t n[2], m[2], ops_len[2], scores[6];
    int64_t pos = 0, opos = 0;
    for (int p = 0; p < 2; ++p) {
        n[p] = (int32_t)strlen(T[p]); m[p] = (int32_t)strlen(O[p]);
        t_off[p] = pos; memcpy(sym + pos, T[p], (size_t)n[p]); pos += n[p];
        o_off[p] = pos; memcpy(sym + pos, O[p], (size_t)m[p]); pos += m[p];
        ops_off[p] = opos; opos += n[p] + m[p];
    }
    tanw_ctx *ctx = NULL;
    if (tanw_create(0, &ctx)) { fprintf(stderr, "tanw_create: %s\n", tanw_last_error(NULL)); return 1; }
    tanw_scoring sc = { 8, -4, -7, -7, -3, 0, -1, 0, NULL };
    if (tanw_align_batch(ctx, sym, pos, t_off, n, o_off, m, 2, &sc, ops, ops_off, opos, ops_len, scores)) {
        fprintf(stderr, "tanw_align_batch: %s\n", tanw_last_error(ctx));
        return 1;
    }
    for (int p = 0; p < 2; ++p) {
        show(T[p], O[p], ops + ops_off[p], ops_len[p]);
        printf("(M, X, Y)[n][m] = (%d, %d, %d)\n", scores[3 * p], scores[3 * p + 1], scores[3 * p + 2]);
    }
    tanw_destroy(ctx);
    return 0;
}
