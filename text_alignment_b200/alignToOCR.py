# -*- coding: utf-8 -*-
"""The consumer of the alignment: OCR character boxes + transcript -> syllable boxes.

Mirrors the part of the reference's ``alignToOCR.process`` that surrounds the hot path
(/root/reference/alignToOCR.py:247-324): abbreviation expansion on the OCR boxes, the call of
``textSeqCompare.perform_alignment`` (:273-276), gap insertion and its length invariant
(:285-292), and the per-syllable regex + box union (:297-324).  ``CharBox`` (:35-58) and
``to_JSON_dict`` (:333-351) keep their reference shape.  Image preprocessing, OCRopus and the
rotation back into raw-image coordinates (:216-245, :327-328) are Gamera / subprocess work and
stay out of scope: this module starts from the list of OCR character boxes.

Host-side glue, not a kernel; the alignment it
calls is the CUDA path.  ``boxes_for_pages`` aligns many pages in one launch.
"""
import re

import numpy as np

from . import latinSyllabification as latsyl
from . import textSeqCompare as tsc


class CharBox(object):
    """A character (or syllable) with its bounding box; same attributes as the reference's
    record (alignToOCR.py:35-58): ``ul``/``lr`` corner tuples (None marks a gap, and then the
    scalar fields are left unset exactly as in the reference), ``ulx, uly, lrx, lry, width,
    height``."""
    __slots__ = ['char', 'ul', 'lr', 'ulx', 'lrx', 'uly', 'lry', 'width', 'height']

    def __init__(self, char, ul=None, lr=None):
        self.char = char
        has_box = ul is not None and lr is not None
        self.ul = tuple(ul) if has_box else None
        self.lr = tuple(lr) if has_box else None
        if has_box:
            (self.ulx, self.uly), (self.lrx, self.lry) = (ul[0], ul[1]), (lr[0], lr[1])
            self.width, self.height = lr[0] - ul[0], lr[1] - ul[1]

    def __repr__(self):
        where = '{}, {}'.format(self.ul, self.lr) if (self.ul and self.lr) else 'empty'
        return '{}: {}'.format(self.char, where)

    def __eq__(self, other):
        return (isinstance(other, CharBox) and self.char == other.char and
                self.ul == other.ul and self.lr == other.lr)

    __hash__ = None


def expand_abbreviations(all_chars, abbreviations=None):
    """alignToOCR.py:251-264.  For every abbreviation (dict order), while it occurs in the OCR
    string, replace its boxes: segment i of the expansion inherits the box of the i-th
    abbreviation character, one CharBox per expanded letter."""
    abbreviations = latsyl.abbreviations if abbreviations is None else abbreviations
    all_chars = list(all_chars)
    for abb, segments in abbreviations.items():
        while True:
            ocr_str = ''.join(str(x.char) for x in all_chars)
            idx = ocr_str.find(abb)
            if idx == -1:
                break
            ins = []
            for i, segment in enumerate(segments):
                src = all_chars[i + idx]
                ins += [CharBox(x, src.ul, src.lr) for x in segment]
            all_chars = all_chars[:idx] + ins + all_chars[idx + len(abb):]
    return all_chars


def insert_gaps(all_chars, ocr_align):
    """alignToOCR.py:285-292: a box-less CharBox('_') wherever the aligned OCR string has a
    gap, so that the box list is index-aligned with ``tra_align``.  One O(L) pass instead of
    the reference's repeated list.insert; the result (and the assertion) are the same."""
    n_gaps = sum(1 for c in ocr_align if c == '_')
    assert len(all_chars) + n_gaps == len(ocr_align), 'all_chars not same length as alignment: ' \
        '{} vs {}'.format(len(all_chars) + n_gaps, len(ocr_align))
    it = iter(all_chars)
    return [CharBox('_') if c == '_' else next(it) for c in ocr_align]


def _syllable_pattern(syl):
    """A syllable's letters with any number of gap symbols between them, exactly the regular
    expression the reference builds (unescaped; alignToOCR.py:301-304)."""
    if len(syl) == 1:
        return syl
    return syl[0] + syl[1:-1].replace('', '_*') + syl[-1]


def syllable_boxes(transcript, tra_align, aligned_chars):
    """alignToOCR.py:277, :297-324: for every syllable of the transcript find, from a moving
    offset, the stretch of ``tra_align`` that spells it with optional gaps in between, and
    union the boxes of the OCR characters aligned to that stretch.  A syllable aligned to no
    OCR character yields no box (:313); a syllable spanning two text lines keeps only the
    boxes of the lower line (:318-320)."""
    out = []
    cursor = 0
    for syl in latsyl.syllabify_text(transcript):
        if not syl:
            continue
        hit = re.search(_syllable_pattern(syl), tra_align[cursor:])
        lo, hi = cursor + hit.start(), cursor + hit.end()
        cursor = hi
        boxed = [c for c in aligned_chars[lo:hi] if c.lr is not None]
        if not boxed:
            continue
        tops = set(c.uly for c in boxed)
        if len(tops) > 1:
            boxed = [c for c in boxed if c.uly == max(tops)]
        out.append(CharBox(syl,
                           (min(c.ulx for c in boxed), min(c.uly for c in boxed)),
                           (max(c.lrx for c in boxed), max(c.lry for c in boxed))))
    return out


def boxes_for_page(transcript, all_chars, seq_align_params=None, device=0):
    """One page: what ``process`` does between OCR and un-rotation (alignToOCR.py:247-324).
    Returns (syl_boxes, all_chars_after_expansion, tra_align, ocr_align)."""
    return boxes_for_pages([(transcript, all_chars)], seq_align_params, devices=[device])[0]


def boxes_for_pages(pages, seq_align_params=None, devices=None):
    """Many pages, one alignment launch per device."""
    expanded = [expand_abbreviations(chars) for _, chars in pages]
    pairs = [(list(transcript), list(''.join(x.char for x in chars)))          # :267, :273
             for (transcript, _), chars in zip(pages, expanded)]
    aligned = tsc.perform_alignment_batch(pairs, scoring_system=seq_align_params, devices=devices)
    out = []
    for (transcript, _), chars, (tra, ocr) in zip(pages, expanded, aligned):
        tra_align = ''.join(tra)                                               # :275-276
        ocr_align = ''.join(ocr)
        aligned_chars = insert_gaps(chars, ocr_align)
        out.append((syllable_boxes(transcript, tra_align, aligned_chars), chars, tra_align, ocr_align))
    return out


def to_JSON_dict(syl_boxes, lines_peak_locs):
    """alignToOCR.py:333-351."""
    med_line_spacing = np.quantile(np.diff(lines_peak_locs), 0.75)
    data = {'median_line_spacing': med_line_spacing, 'syl_boxes': []}
    for s in syl_boxes:
        data['syl_boxes'].append({'syl': s.char,
                                  'ul': [int(s.ul[0]), int(s.ul[1])],
                                  'lr': [int(s.lr[0]), int(s.lr[1])]})
    return data


# ---- wire formats either side of the path --------------------------------

def clean_special_chars(inp):
    """alignToOCR.py:61-72: OCRopus' '~' never reaches the aligner."""
    return inp.replace('~', '')


def parse_llocs(lines, x_min, y_min, y_max):
    """One text line of OCRopus ``--llocs`` output -> CharBoxes (alignToOCR.py:153-182).

    Each TSV record is (character, x position of its RIGHT edge inside the strip); a
    character's box therefore runs from the previous record's x to its own x, over the
    strip's full height.  Records whose character is '~' or empty advance the x position but
    are set aside (returned second)."""
    chars, other = [], []
    prev_x = x_min
    for rec in lines:
        fields = rec.rstrip('\n').split('\t')
        cur_x = int(np.round(float(fields[1]) + x_min))
        box = CharBox(fields[0] if fields[0] in ('~', '') else clean_special_chars(fields[0]),
                      (prev_x, y_min), (cur_x, y_max))
        (other if fields[0] in ('~', '') else chars).append(box)
        prev_x = cur_x
    return chars, other


def read_llocs_files(paths, strips):
    """All text lines of a page: ``strips[i]`` gives (offset_x, offset_y, height) of line i."""
    import io
    all_chars = []
    for path, (off_x, off_y, height) in zip(paths, strips):
        with io.open(path, encoding='utf-8') as f:
            chars, _ = parse_llocs(list(f), off_x, off_y, off_y + height)
        all_chars += chars
    return all_chars


# ---- the same consumer on packed arrays ---------------------------
#
# The object path above builds ~3 000 CharBox objects and runs ~400 regular-expression searches per
# page (about 12 ms per page in CPython) behind an aligner that needs a microsecond per page.  The
# array path keeps a page as (transcript str, OCR str, boxes int32[k, 4]) and does the per-column
# work natively (csrc/tanw_consumer.cu): op strings straight from the device batch, syllable
# spans as transcript index ranges, boxes by segmented min/max.  Results are identical to the
# object path (tests/test_consumer.py checks both against the reference's own process()).

import functools
import json

from . import _native


@functools.lru_cache(maxsize=1 << 16)
def _word_syllables(word):
    """syllabify_word, memoised: chant texts repeat a small vocabulary."""
    return tuple(latsyl.syllabify_word(word))


def regex_free(transcript):
    """True when the reference's per-syllable regular expression (alignToOCR.py:297-309) reduces
    to "the columns from the syllable's first letter to its last": every character is a letter,
    a digit or a space -- no regex metacharacter and no gap symbol."""
    body = transcript.replace(' ', '')
    return body == '' or body.isalnum()


def syllable_spans(transcript, strings=True):
    """-> (syllables, bounds int32[S, 2]): syllable s is transcript[bounds[s, 0]:bounds[s, 1]].
    Words are what ``split(' ')`` gives (latinSyllabification.py:171); a word's syllables are
    consecutive pieces of it (None is returned if that ever fails to hold, and the caller keeps
    the regular-expression path)."""
    bounds = _native.syllable_bounds(transcript)
    if bounds is not None:                                   # the native syllabifier (ASCII text)
        return (_native.text_slices(transcript, bounds) if strings else None), bounds
    words = transcript.split(' ')
    per_word = [_word_syllables(w) for w in words]
    syls = [s for ws in per_word for s in ws]
    S = len(syls)
    if S == 0:
        return syls, np.zeros((0, 2), np.int32)
    counts = np.fromiter(map(len, per_word), dtype=np.int64, count=len(words))
    lens = np.fromiter(map(len, syls), dtype=np.int64, count=S)
    word_lens = np.fromiter(map(len, words), dtype=np.int64, count=len(words))
    if lens.min() < 1 or not np.array_equal(np.add.reduceat(np.append(lens, 0), np.minimum(np.cumsum(counts) - counts, S))
                                            * (counts > 0), word_lens * (counts > 0)) or np.any((counts == 0) & (word_lens > 0)):
        return None
    word_start = np.cumsum(word_lens + 1) - (word_lens + 1)     # after the words and single spaces before it
    excl = np.cumsum(lens) - lens                               # letters of all syllables before this one
    first = np.minimum(np.cumsum(counts) - counts, S - 1)       # a word's first syllable
    starts = np.repeat(word_start - excl[first], counts) + excl
    return syls, np.stack([starts, starts + lens], axis=1).astype(np.int32)


def expand_abbreviations_arrays(ocr, boxes, abbreviations=None):
    """alignToOCR.py:251-264 on (OCR string, boxes int32[k, 4]): same replacements in the same
    order as ``expand_abbreviations``; every expanded letter inherits the box of the abbreviation
    character its segment stands for.

    Which box a character uses is carried in a second STRING (character i = chr(box index)), so
    that the reference's list surgery (two slices and a concatenation of the whole page per
    occurrence) becomes string slicing, a memcpy.  After a replacement at ``idx`` the next leftmost
    occurrence cannot lie wholly left of ``idx - len(abb) + 1`` (it would have been found first),
    so the search resumes there instead of at 0 -- the same occurrences in the same order."""
    abbreviations = latsyl.abbreviations if abbreviations is None else abbreviations
    if len(ocr) >= 0xD000:                      # box indices must stay below the surrogate range
        return _expand_abbreviations_arrays_lists(ocr, boxes, abbreviations)
    src = None
    for abb, segments in abbreviations.items():
        if not abb:
            continue
        start = 0
        ins = ''.join(segments)
        while True:
            idx = ocr.find(abb, start)
            if idx == -1:
                break
            if src is None:
                src = _identity_string(len(ocr))
            ins_src = ''.join([src[idx + i] * len(seg) for i, seg in enumerate(segments)])
            ocr = ocr[:idx] + ins + ocr[idx + len(abb):]
            src = src[:idx] + ins_src + src[idx + len(abb):]
            start = max(0, idx - len(abb) + 1)
    if src is None:
        return ocr, boxes
    index = np.frombuffer(src.encode('utf-32-le'), dtype=np.uint32).astype(np.intp)
    return ocr, np.ascontiguousarray(np.asarray(boxes, dtype=np.int32).reshape(-1, 4)[index])


_IDENTITY = ['']


def _identity_string(k):
    """chr(0) chr(1) ... chr(k-1): built once, sliced afterwards."""
    if len(_IDENTITY[0]) < k:
        _IDENTITY[0] = ''.join(map(chr, range(max(k, 4096))))
    return _IDENTITY[0][:k]


def _expand_abbreviations_arrays_lists(ocr, boxes, abbreviations):
    """The same with a list of box indices (pages too long for the string form)."""
    src = None                                  # index of the box each current character uses
    for abb, segments in abbreviations.items():
        while True:
            idx = ocr.find(abb)
            if idx == -1:
                break
            if src is None:
                src = list(range(len(ocr)))
            ins = ''.join(segments)
            ins_src = [src[idx + i] for i, seg in enumerate(segments) for _ in seg]
            ocr = ocr[:idx] + ins + ocr[idx + len(abb):]
            src = src[:idx] + ins_src + src[idx + len(abb):]
    if src is None:
        return ocr, boxes
    return ocr, np.ascontiguousarray(np.asarray(boxes, dtype=np.int32).reshape(-1, 4)[src])


def page_from_llocs(llocs, strips):
    """All text lines of a page from raw .llocs bytes: ``llocs[i]`` is the content of line i's
    file, ``strips[i]`` its (offset_x, offset_y, height) (alignToOCR.py:153-182).
    -> (OCR string, boxes int32[k, 4])."""
    cps, bxs = [], []
    for text, (off_x, off_y, height) in zip(llocs, strips):
        c, b = _native.parse_llocs(text, off_x, off_y, off_y + height)
        cps.append(c)
        bxs.append(b)
    cp = np.concatenate(cps) if cps else np.zeros(0, np.uint32)
    boxes = np.concatenate(bxs) if bxs else np.zeros((0, 4), np.int32)
    return cp.astype('<u4').tobytes().decode('utf-32-le', 'surrogatepass'), boxes


def boxes_for_pages_arrays(pages, seq_align_params=None, devices=None):
    """Many pages on arrays: ``pages`` = [(transcript str, OCR str, boxes int32[k, 4])] (the OCR
    characters before abbreviation expansion).  One alignment launch per device, one native pass
    over the op strings.  -> per page (syllables, syl_boxes int32[s, 4]): the syllables that got a
    box, in order, exactly the ``CharBox`` list ``boxes_for_pages`` returns.

    Pages whose transcript holds a regular-expression metacharacter or '_' take the object path
    (the reference's regex then means something else than "first letter to last letter")."""
    results = [None] * len(pages)
    fast = []
    for k, (transcript, ocr, boxes) in enumerate(pages):
        spans = syllable_spans(transcript, strings=False) if regex_free(transcript) else None
        if spans is None:
            chars = [CharBox(c, (int(b[0]), int(b[1])), (int(b[2]), int(b[3]))) for c, b in zip(ocr, np.asarray(boxes).reshape(-1, 4))]
            syl_boxes = boxes_for_page(transcript, chars, seq_align_params, device=(devices or [0])[0])[0]
            results[k] = ([b.char for b in syl_boxes],
                          np.array([[b.ulx, b.uly, b.lrx, b.lry] for b in syl_boxes], dtype=np.int32).reshape(-1, 4))
        else:
            ocr2, boxes2 = expand_abbreviations_arrays(ocr, boxes)
            fast.append((k, transcript, ocr2, np.asarray(boxes2, dtype=np.int32).reshape(-1, 4), spans))
    if fast:
        ops, ops_off, ops_len = tsc.align_strings([(t, o) for _, t, o, _, _ in fast], seq_align_params, devices)
        syl_off = np.zeros(len(fast) + 1, dtype=np.int64)
        box_off = np.zeros(len(fast) + 1, dtype=np.int64)
        np.cumsum([f[4][1].shape[0] for f in fast], out=syl_off[1:])
        np.cumsum([f[3].shape[0] for f in fast], out=box_off[1:])
        out, has = _native.syllable_boxes(ops, ops_off, ops_len, np.concatenate([f[4][1] for f in fast]), syl_off,
                                          np.concatenate([f[3] for f in fast]), box_off)
        for j, (k, transcript, _, _, (syls, bounds)) in enumerate(fast):
            lo, hi = int(syl_off[j]), int(syl_off[j + 1])
            keep = has[lo:hi]
            if syls is None:                      # the native syllabifier's ranges: only the kept syllables become strings
                kept = _native.text_slices(transcript, bounds, keep=keep)
            else:
                kept = [s for s, h in zip(syls, keep.tolist()) if h]
            results[k] = (kept, out[lo:hi][keep])
    return results


def to_JSON_bytes(syllables, syl_boxes, lines_peak_locs):
    """alignToOCR.to_JSON_dict (:333-351) + json.dumps, written natively from the arrays of
    ``boxes_for_pages_arrays``."""
    med_line_spacing = float(np.quantile(np.diff(lines_peak_locs), 0.75))
    if not np.isfinite(med_line_spacing):
        return json.dumps({'median_line_spacing': med_line_spacing,
                           'syl_boxes': [{'syl': s, 'ul': [int(b[0]), int(b[1])], 'lr': [int(b[2]), int(b[3])]}
                                         for s, b in zip(syllables, syl_boxes)]}).encode()
    return _native.boxes_to_json(syllables, syl_boxes, np.ones(len(syllables), np.uint8), med_line_spacing)
