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
