"""The consumer of the alignment: syllabifier,
abbreviation expansion, gap insertion, syllable -> box mapping.

CPU tests pin the host-side port against golden output of the reference's own
alignToOCR.process (driven with mocked Gamera / OCR, tests/golden/make_golden.py) with the
CPU oracle standing in for the aligner, and against the live reference when it is present.
The GPU test runs the real pipeline: CUDA alignment + this glue == the reference's boxes."""
import random

import pytest

from conftest import load_golden
from text_alignment_b200 import alignToOCR as atocr
from text_alignment_b200 import latinSyllabification as latsyl
from text_alignment_b200 import synth


@pytest.fixture(scope='module')
def golden():
    return load_golden('consumer.json')


def _oracle_batch(pairs, scoring_system=None, devices=None):
    from oracle import nw_oracle
    return [nw_oracle.perform_alignment(a, b, scoring_system) for a, b in pairs]


def _boxes(page):
    return [atocr.CharBox(c, tuple(ul), tuple(lr)) for c, ul, lr in page['boxes']]


def _as_tuples(syl_boxes):
    return [[b.char, [int(v) for v in b.ul], [int(v) for v in b.lr]] for b in syl_boxes]


def test_syllabifier_golden(golden):
    for word, want in golden['syllables']:
        if want is None:                       # the reference never terminates on this word
            with pytest.raises(ValueError):
                latsyl.syllabify_word(word)
        else:
            assert latsyl.syllabify_word(word) == want, word


def test_syllabifier_reference_demo():
    # latinSyllabification.py:215-219, output recorded in SURVEY.md Appendix B
    text = 'quaecumque ejus michi antiphonum assistens alleluya dixit extra exhibeamus'
    assert ' '.join(latsyl.syllabify_text(text)) == \
        'quae cum que e jus mi chi an ti pho num as si stens al le lu ya dix it ex tra ex hi be a mus'
    assert latsyl.syllabify_word('euouae') == ['e', 'u', 'o', 'u', 'ae']
    assert latsyl.syllabify_word('cuius') == ['cu', 'ius'] and latsyl.syllabify_word('eius') == ['e', 'ius']


def test_abbreviation_expansion_golden(golden):
    for page in golden['pages']:
        chars = atocr.expand_abbreviations(_boxes(page))
        assert ''.join(c.char for c in chars) == page['expanded_ocr']


def test_abbreviation_expansion_on_arrays_equals_the_object_path():
    """expand_abbreviations_arrays (string surgery, search resumed near the last replacement) against
    the object path that follows alignToOCR.py:251-264 literally, and against its own list form:
    texts dense in abbreviations, adjacent and nested candidates, user-supplied tables whose
    expansions create new occurrences."""
    import numpy as np
    rng = random.Random(5)
    tables = [None,
              {'ab': ['b', 'a'], 'bb': ['ab']},                    # an expansion re-creates a key to its left
              {'aa': ['a'], 'ba': ['a', 'b']},
              {'x': ['yy'], 'yyy': ['x', 'z', 'y']}]
    alphabets = ['dnsūealā^ēō l', 'ab', 'ab', 'xyz']
    for table, alpha in zip(tables, alphabets):
        for _ in range(150):
            text = ''.join(rng.choice(alpha) for _ in range(rng.randint(0, 60)))
            boxes = np.array([[i, 2 * i, i + 5, 2 * i + 7] for i in range(len(text))], dtype=np.int32).reshape(-1, 4)
            kw = {} if table is None else {'abbreviations': table}
            got_s, got_b = atocr.expand_abbreviations_arrays(text, boxes, **kw)
            lst_s, lst_b = atocr._expand_abbreviations_arrays_lists(text, boxes, table or latsyl.abbreviations)
            chars = atocr.expand_abbreviations([atocr.CharBox(c, (int(b[0]), int(b[1])), (int(b[2]), int(b[3])))
                                                for c, b in zip(text, boxes)], **kw)
            assert got_s == lst_s == ''.join(c.char for c in chars)
            want = np.array([[c.ulx, c.uly, c.lrx, c.lry] for c in chars], dtype=np.int32).reshape(-1, 4)
            assert np.array_equal(np.asarray(got_b).reshape(-1, 4), want)
            assert np.array_equal(np.asarray(lst_b).reshape(-1, 4), want)


def test_text_slices_helper_and_python_fallback(monkeypatch):
    """_native.text_slices: the CPython helper (csrc/tanw_pylist.c) and the pure-Python form give
    the same lists, with and without a keep mask; a range outside the text is an error."""
    import numpy as np
    from text_alignment_b200 import _native
    text = 'kyrieeleisonā ē'
    bounds = np.array([[0, 2], [2, 5], [5, 5], [5, 12], [12, 13], [14, 15]], dtype=np.int32)
    keep = np.array([1, 0, 1, 1, 0, 1], dtype=np.uint8)
    want_all = [text[a:b] for a, b in bounds.tolist()]
    want_kept = [w for w, h in zip(want_all, keep.tolist()) if h]
    assert _native.text_slices(text, bounds) == want_all
    assert _native.text_slices(text, bounds, keep=keep) == want_kept
    assert _native.text_slices(text, bounds, keep=keep.astype(bool)) == want_kept
    assert _native.text_slices('', np.zeros((0, 2), np.int32)) == []
    if _native.pylist() is not None:
        with pytest.raises(ValueError):
            _native.text_slices('abc', np.array([[1, 9]], dtype=np.int32))
    monkeypatch.setattr(_native, 'pylist', lambda: None)
    assert _native.text_slices(text, bounds) == want_all
    assert _native.text_slices(text, bounds, keep=keep) == want_kept


def test_abbreviation_boxes_inherit_from_source_character():
    chars = [atocr.CharBox(c, (10 * k, 0), (10 * k + 9, 5)) for k, c in enumerate('a dns b')]
    out = atocr.expand_abbreviations(chars)
    assert ''.join(c.char for c in out) == 'a dominus b'
    # 'd' -> 'do', 'n' -> 'mi', 's' -> 'nus' (latinSyllabification.py:10, alignToOCR.py:261-263)
    assert [c.ul[0] for c in out[2:9]] == [20, 20, 30, 30, 40, 40, 40]


def test_insert_gaps_and_invariant():
    chars = [atocr.CharBox(c, (k, 0), (k + 1, 1)) for k, c in enumerate('abc')]
    out = atocr.insert_gaps(chars, 'a__b_c')
    assert [c.char for c in out] == list('a__b_c') and out[1].ul is None and out[3].ul == (1, 0)
    with pytest.raises(AssertionError):
        atocr.insert_gaps(chars, 'a_b')


def test_pages_golden_with_oracle_aligner(golden, monkeypatch):
    monkeypatch.setattr(atocr.tsc, 'perform_alignment_batch', _oracle_batch)
    for page in golden['pages']:
        syl_boxes, chars, tra, ocr = atocr.boxes_for_page(page['transcript'], _boxes(page), page['params'])
        assert _as_tuples(syl_boxes) == page['syl_boxes'], page['seed']
        assert len(tra) == len(ocr)


def test_json_shape():
    b = [atocr.CharBox('glo', (100, 200), (154, 260))]
    d = atocr.to_JSON_dict(b, [100, 240, 380, 520])
    assert d['syl_boxes'] == [{'syl': 'glo', 'ul': [100, 200], 'lr': [154, 260]}]
    assert d['median_line_spacing'] == 140


def test_live_reference_process_100_pages(monkeypatch):
    """SURVEY 8(d) parity gate: identical syllable boxes through the reference's own
    alignToOCR.process on >= 100 pages (small pages so that the pure-Python reference
    aligner finishes in about a minute)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference checkout not present')
    import make_golden
    A = ref_loader.load_aligntoocr()
    monkeypatch.setattr(atocr.tsc, 'perform_alignment_batch', _oracle_batch)
    monkeypatch.setattr(atocr.tsc, 'align_strings', _oracle_align_strings)
    rng = random.Random(12)
    for k in range(100):
        n = rng.randint(40, 110)
        t, boxes = synth.make_page(30000 + k, n, int(n * rng.uniform(1.0, 1.6)), 1, 9, abbreviations=k % 3 == 0)
        params = None if k % 4 else [rng.choice([5, 8, 11]), rng.choice([-4, -7, -10]), rng.choice([-2, -5, -7]),
                                     rng.choice([-2, -5, -7]), rng.choice([0, -3, -5]), rng.choice([0, -3, -5])]
        ref_boxes, _, _, ref_chars = make_golden.run_reference_process(A, t, boxes, params)
        got, chars, _, _ = atocr.boxes_for_page(t, [atocr.CharBox(c, ul, lr) for c, ul, lr in boxes], params)
        assert _as_tuples(got) == _as_tuples(ref_boxes), (k, t)
        assert [c.char for c in chars] == [c.char for c in ref_chars]
        arr = atocr.boxes_for_pages_arrays([(t,) + _page_arrays(boxes)], params)[0]       # the array path too
        assert _arrays_as_tuples(arr) == _as_tuples(ref_boxes), (k, t)


@pytest.mark.gpu
def test_pages_golden_on_gpu(golden):
    """The real pipeline: CUDA alignment behind the drop-in module + host glue reproduces the
    syllable boxes of the reference's alignToOCR.process, page by page and as one batch."""
    pages = [(p['transcript'], _boxes(p)) for p in golden['pages'] if p['params'] is None]
    want = [p['syl_boxes'] for p in golden['pages'] if p['params'] is None]
    out = atocr.boxes_for_pages(pages)
    assert [_as_tuples(o[0]) for o in out] == want
    for page in golden['pages']:
        syl_boxes, _, _, _ = atocr.boxes_for_page(page['transcript'], _boxes(page), page['params'])
        assert _as_tuples(syl_boxes) == page['syl_boxes']


# ---- wire format: OCRopus .llocs -> CharBox (alignToOCR.py:153-182) -------------------------------

LLOCS = [u'g\t12.4\n', u'l\t20.0\n', u'~\t22.5\n', u'o\t31.6\n', u'\t33.0\n', u'r\t40.49\n', u'ū\t55.5\n', u'a\t70.51\n']


def test_parse_llocs_right_edges_become_boxes():
    chars, other = atocr.parse_llocs(LLOCS, 100, 50, 110)
    assert ''.join(c.char for c in chars) == u'glorūa'
    assert [(c.ul, c.lr) for c in chars][:3] == [((100, 50), (112, 110)), ((112, 50), (120, 110)), ((122, 50), (132, 110))]
    assert [c.char for c in other] == ['~', ''] and other[0].ul == (120, 50)
    assert chars[-1].lr == (171, 110)          # np.round(70.51 + 100)


def test_parse_llocs_matches_reference_ocr_reader(tmp_path, monkeypatch):
    """Drive the reference's perform_ocr_with_ocropus with the subprocess patched out and
    .llocs files prepared on disk; our reader must produce the same boxes."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference checkout not present')
    A = ref_loader.load_aligntoocr()

    class Strip(object):
        def __init__(self, ox, oy, h):
            self.offset_x, self.offset_y, self.height = ox, oy, h

        def save_image(self, path):
            pass
    strips = [Strip(100, 50, 60), Strip(90, 190, 58)]
    monkeypatch.chdir(tmp_path)
    (tmp_path / 'wk').mkdir()
    import io
    for i, recs in enumerate([LLOCS, [u'd\t9.5\n', u'n\t19.5\n', u's\t30.2\n', u'~\t31.0\n']]):
        with io.open(str(tmp_path / 'wk' / '_{}.llocs'.format(i)), 'w', encoding='utf-8') as f:
            f.writelines(recs)
    monkeypatch.setattr(A.subprocess, 'check_call', lambda *a, **k: 0)
    ref = A.perform_ocr_with_ocropus(strips, 'nomodel', 'wk')
    got = atocr.read_llocs_files([str(tmp_path / 'wk' / '_{}.llocs'.format(i)) for i in range(2)],
                                 [(s.offset_x, s.offset_y, s.height) for s in strips])
    assert [(c.char, tuple(c.ul), tuple(c.lr)) for c in ref] == [(c.char, c.ul, c.lr) for c in got]


# ---- the same consumer on packed arrays (native: csrc/tanw_consumer.cu) ---------------------------

def _page_arrays(boxes):
    import numpy as np
    ocr = ''.join(c for c, _, _ in boxes)
    arr = np.array([[ul[0], ul[1], lr[0], lr[1]] for _, ul, lr in boxes], dtype=np.int32).reshape(-1, 4)
    return ocr, arr


def _oracle_align_strings(pairs, scoring_system=None, devices=None):
    """Stand-in for textSeqCompare.align_strings: op strings from the CPU oracle."""
    import numpy as np
    from oracle import nw_oracle
    ops, off, lens = [], [0], []
    for t, o in pairs:
        tra, ocr = nw_oracle.perform_alignment(list(t), list(o), scoring_system)
        # '_' in an aligned sequence is a gap unless the other one has a gap there too (never both)
        it_t, it_o = iter(t), iter(o)
        row = []
        x = y = 0
        for a, b in zip(tra, ocr):
            if b == '_' and not (y < len(o) and o[y] == '_' and a == '_'):
                row.append(1); x += 1
            elif a == '_' and not (x < len(t) and t[x] == '_'):
                row.append(2); y += 1
            else:
                row.append(0); x += 1; y += 1
        del it_t, it_o
        ops += row + [9] * (len(t) + len(o) - len(row))
        lens.append(len(row))
        off.append(off[-1] + len(t) + len(o))
    return np.array(ops, dtype=np.uint8), np.array(off[:-1], dtype=np.int64), np.array(lens, dtype=np.int32)


def _arrays_as_tuples(res):
    syls, boxes = res
    return [[s, [int(b[0]), int(b[1])], [int(b[2]), int(b[3])]] for s, b in zip(syls, boxes.tolist())]


def test_syllable_spans_are_the_syllables_in_place():
    import numpy as np
    for text in ['gloria in excelsis deo', 'a', '', '  double  spaces here ', 'quaecumque ejus michi euouae cuius']:
        syls, bounds = atocr.syllable_spans(text)
        assert syls == latsyl.syllabify_text(text)
        assert [text[a:b] for a, b in bounds.tolist()] == syls
        assert bounds.dtype == np.int32
    assert atocr.regex_free('gloria in excelsis') and atocr.regex_free('') and atocr.regex_free('dūs 12')
    assert not atocr.regex_free('glo.ria') and not atocr.regex_free('a_b') and not atocr.regex_free('quid?')


def test_array_consumer_golden_with_oracle_aligner(golden, monkeypatch):
    monkeypatch.setattr(atocr.tsc, 'align_strings', _oracle_align_strings)
    monkeypatch.setattr(atocr.tsc, 'perform_alignment_batch', _oracle_batch)
    default = [p for p in golden['pages'] if p['params'] is None]
    out = atocr.boxes_for_pages_arrays([(p['transcript'],) + _page_arrays(p['boxes']) for p in default])
    assert [_arrays_as_tuples(o) for o in out] == [p['syl_boxes'] for p in default]
    for page in golden['pages']:
        got = atocr.boxes_for_pages_arrays([(page['transcript'],) + _page_arrays(page['boxes'])], page['params'])[0]
        assert _arrays_as_tuples(got) == page['syl_boxes'], page['seed']


def test_array_consumer_equals_object_path_on_odd_pages(monkeypatch):
    """Multi-line syllables (lowest line wins), syllables aligned to nothing, abbreviations, and a
    transcript with a regex metacharacter (which must take the object path)."""
    import numpy as np
    monkeypatch.setattr(atocr.tsc, 'align_strings', _oracle_align_strings)
    monkeypatch.setattr(atocr.tsc, 'perform_alignment_batch', _oracle_batch)
    rng = random.Random(5)
    pages = []
    for k in range(40):
        n = rng.randint(30, 90)
        t, boxes = synth.make_page(41000 + k, n, int(n * rng.uniform(0.5, 1.7)), 1, 7, abbreviations=k % 2 == 0)
        if k % 5 == 0:
            t = t.replace(' ', '. ', 1)                   # a metacharacter: regex path
        boxes = [(c, (ul[0], ul[1] + (7 if rng.random() < 0.1 else 0)), lr) for c, ul, lr in boxes]   # ragged lines
        pages.append((t, boxes))
    want = [atocr.boxes_for_page(t, [atocr.CharBox(c, ul, lr) for c, ul, lr in boxes])[0] for t, boxes in pages]
    got = atocr.boxes_for_pages_arrays([(t,) + _page_arrays(boxes) for t, boxes in pages])
    for w, g in zip(want, got):
        assert _as_tuples(w) == _arrays_as_tuples(g)
        assert g[1].dtype == np.int32


def test_native_llocs_reader_equals_python_reader():
    import numpy as np
    from text_alignment_b200 import _native
    text = u''.join(LLOCS).encode('utf-8')
    cps, boxes = _native.parse_llocs(text, 100, 50, 110)
    chars, _ = atocr.parse_llocs(LLOCS, 100, 50, 110)
    assert ''.join(map(chr, cps.tolist())) == ''.join(c.char for c in chars)
    assert boxes.tolist() == [[c.ulx, c.uly, c.lrx, c.lry] for c in chars]
    # half-to-even rounding as np.round, CRLF line ends, a last record without newline
    cps, boxes = _native.parse_llocs(b'a\t0.5\r\nb\t1.5\r\nc\t2.5', 0, 0, 9)
    assert boxes[:, 2].tolist() == [int(np.round(v)) for v in (0.5, 1.5, 2.5)] == [0, 2, 2]
    ocr, arr = atocr.page_from_llocs([text, b'd\t9.5\nn\t19.5\ns\t30.2\n~\t31.0\n'], [(100, 50, 60), (90, 190, 58)])
    assert ocr == u'glorūadns' and arr.shape == (9, 4) and arr[6].tolist() == [90, 190, 100, 248]
    for bad in (b'a\n', b'ab\t3\n', b'a\tx\n', b'\xff\t3\n'):
        with pytest.raises(ValueError):
            _native.parse_llocs(bad, 0, 0, 1)


def test_native_json_equals_json_dumps():
    import json
    import numpy as np
    syls = ['glo', u'dūs', 'a"b\\', u'\U0001d11e', 'x\ty']
    boxes = np.array([[100, 200, 154, 260], [-3, 0, 7, 9], [1, 2, 3, 4], [5, 6, 7, 8], [9, 9, 9, 9]], dtype=np.int32)
    peaks = [100, 240, 380, 521]
    want = json.dumps(atocr.to_JSON_dict([atocr.CharBox(s, (b[0], b[1]), (b[2], b[3])) for s, b in zip(syls, boxes.tolist())], peaks))
    assert atocr.to_JSON_bytes(syls, boxes, peaks) == want.encode()
    assert json.loads(atocr.to_JSON_bytes([], np.zeros((0, 4), np.int32), peaks)) == {'median_line_spacing': 140.5, 'syl_boxes': []}


def test_native_syllable_boxes_refuses_inconsistent_pages():
    import numpy as np
    from text_alignment_b200 import _native
    ops = np.array([0, 0, 1, 2], dtype=np.uint8)
    bounds = np.array([[0, 2], [2, 3]], dtype=np.int32)
    boxes = np.array([[0, 0, 5, 5], [5, 0, 9, 5], [9, 0, 12, 5]], dtype=np.int32)
    out, has = _native.syllable_boxes(ops, [0], [4], bounds, [0, 2], boxes, [0, 3])
    assert out.tolist() == [[0, 0, 9, 5], [0, 0, 0, 0]] and has.tolist() == [True, False]
    with pytest.raises(AssertionError):                      # one OCR box too many (alignToOCR.py:291)
        _native.syllable_boxes(ops, [0], [4], bounds, [0, 2], np.vstack([boxes, boxes[:1]]), [0, 4])
    with pytest.raises(ValueError):                          # a syllable beyond the transcript
        _native.syllable_boxes(ops, [0], [4], np.array([[0, 2], [2, 5]], np.int32), [0, 2], boxes, [0, 3])


@pytest.mark.gpu
def test_array_consumer_on_gpu_llocs_to_json(golden):
    """.llocs bytes in, JSON bytes out, alignment on the device, no CharBox anywhere: equal to the
    object path on the same pages (which equals the reference's process(), see above)."""
    import json
    pages = [p for p in golden['pages'] if p['params'] is None]
    arrays = []
    for p in pages:
        # one .llocs "file" per text line of the page, rebuilt from the golden boxes
        lines = {}
        for c, ul, lr in p['boxes']:
            lines.setdefault((ul[1], lr[1]), []).append((c, ul[0], lr[0]))
        llocs, strips = [], []
        for (y0, y1), recs in sorted(lines.items()):
            x0 = recs[0][1]
            llocs.append(u''.join(u'{}\t{}\n'.format(c, float(x1 - x0)) for c, _, x1 in recs).encode('utf-8'))
            strips.append((x0, y0, y1 - y0))
        ocr, boxes = atocr.page_from_llocs(llocs, strips)
        assert ocr == ''.join(c for c, _, _ in p['boxes'])
        assert boxes.tolist() == [[ul[0], ul[1], lr[0], lr[1]] for _, ul, lr in p['boxes']]
        arrays.append((p['transcript'], ocr, boxes))
    out = atocr.boxes_for_pages_arrays(arrays)
    assert [_arrays_as_tuples(o) for o in out] == [p['syl_boxes'] for p in pages]
    peaks = [200, 340, 480, 620]
    for (syls, boxes), p in zip(out, pages):
        want = json.dumps({'median_line_spacing': 140.0,
                           'syl_boxes': [{'syl': s, 'ul': ul, 'lr': lr} for s, ul, lr in p['syl_boxes']]})
        assert atocr.to_JSON_bytes(syls, boxes, peaks) == want.encode()


def test_native_syllabifier_equals_python(golden):
    """csrc/tanw_consumer.cu restates latinSyllabification on bytes: same syllables on the golden
    words (which come from the reference itself), on random letter salads, on whole texts; the same
    refusal of a word without a vowel; non-ASCII text is left to the Python syllabifier."""
    from text_alignment_b200 import _native
    for word, want in golden['syllables']:
        if not (word.isascii() and (word == '' or word.isalnum())):
            continue
        if want is None:
            with pytest.raises(ValueError):
                _native.syllable_bounds(word)
        else:
            assert [word[a:b] for a, b in _native.syllable_bounds(word).tolist()] == want, word
    rng = random.Random(1)
    for k in range(20000):
        w = ''.join(rng.choice('aeiouyqchpflrstbxt' if k % 2 else 'abcdefghijklmnopqrstuvwxyz') for _ in range(rng.randint(1, 12)))
        try:
            want = latsyl.syllabify_word(w)
        except ValueError:
            want = None
        if want is None:
            with pytest.raises(ValueError):
                _native.syllable_bounds(w)
        else:
            assert [w[a:b] for a, b in _native.syllable_bounds(w).tolist()] == want, w
    text = 'quaecumque ejus  michi antiphonum assistens alleluya dixit extra exhibeamus euouae cuius eius '
    assert [text[a:b] for a, b in _native.syllable_bounds(text).tolist()] == latsyl.syllabify_text(text)
    assert _native.syllable_bounds(u'dūs') is None and _native.syllable_bounds('glo.ria') is None
    assert _native.syllable_bounds('').shape == (0, 2)
