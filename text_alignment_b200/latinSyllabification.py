# -*- coding: utf-8 -*-
"""Latin syllable splitter with the behaviour of the reference's ``latinSyllabification``
(/root/reference/latinSyllabification.py:5-19, :22-109, :170-174) -- the consumer of the
aligned transcript needs exactly these syllables (alignToOCR.py:277, :297-324).

Host-side glue, not a kernel: O(len) string rules.  Re-implemented, not copied; pinned against
the reference on word lists and the reference's own demo sentence by tests/test_consumer.py.

Rules (reference :22-109): a word is cut into units -- consonant clusters first, then
diphthongs, each class in its listed priority order, each occurrence taken left to right by
``str.split``; whatever is left falls apart into single letters.  Vowels and diphthongs seed
syllables.  Then, repeatedly, every non-seed unit directly before a seed is glued onto it, and
after that every non-seed unit directly after a seed; until only seeds remain.

Deviation (documented): a word without any vowel never terminates in the reference (:71 loops
forever); here it raises ``ValueError``.
"""

consonant_groups = ['qu', 'ch', 'ph', 'fl', 'fr', 'st', 'br', 'cr', 'cl', 'pr', 'tr', 'ct', 'th']
diphthongs = ['ae', 'au', 'ei', 'oe', 'ui', 'ya', 'ex', 'ix']
vowels = ['a', 'e', 'i', 'o', 'u', 'y']

# OCR abbreviation -> syllables it stands for (reference :9-19; insertion order matters to
# alignToOCR's expansion loop)
abbreviations = {
    u'dns': ['do', 'mi', 'nus'],
    u'dūs': ['do', 'mi', 'nus'],
    u'dne': ['do', 'mi', 'ne'],
    u'alla': ['al', 'le', 'lu', 'ia'],
    u'^': ['us'],
    u'ā': ['am'],
    u'ē': ['em'],
    u'ū': ['um'],
    u'ō': ['om']
}

_FIXED = {'euouae': ['e', 'u', 'o', 'u', 'ae'], 'cuius': ['cu', 'ius'], 'eius': ['e', 'ius']}   # :30-35
_MARK = '*'       # the reference tags finished units with '*', so a literal '*' in a word
#                   behaves like a tag there; keeping the same tag keeps that behaviour


def _cut_units(word):
    """Clusters, then diphthongs, then single letters (:37-63)."""
    parts = [word]
    for unit in consonant_groups + diphthongs:
        nxt = []
        for part in parts:
            if _MARK in part:
                nxt.append(part)
                continue
            pieces = part.split(unit)
            for idx, piece in enumerate(pieces):
                if piece:
                    nxt.append(piece)
                if idx + 1 < len(pieces):
                    nxt.append(unit + _MARK)
        parts = nxt
    units = []
    for part in parts:
        if _MARK in part:
            units.append(part.replace(_MARK, ''))
        else:
            units.extend(part)
    return units


def _glue(units, seed_first):
    """One sweep: join (consonant, seed) pairs when not seed_first, (seed, consonant) pairs when
    seed_first; a unit takes part in at most one join per sweep (:73-105)."""
    out = []
    i = 0
    while i < len(units):
        if i + 1 < len(units):
            a, b = units[i], units[i + 1]
            a_seed, b_seed = _MARK in a, _MARK in b
            if (a_seed and not b_seed) if seed_first else (b_seed and not a_seed):
                out.append(a + b)
                i += 2
                continue
        out.append(units[i])
        i += 1
    return out


def syllabify_word(inp):
    if inp in _FIXED:
        return list(_FIXED[inp])
    units = _cut_units(inp)
    seeds = set(vowels + diphthongs)
    units = [u + _MARK if u in seeds else u for u in units]      # :66-68
    if units and not any(_MARK in u for u in units):
        raise ValueError('cannot syllabify {!r}: no vowel (the reference loops forever here)'.format(inp))
    while not all(_MARK in u for u in units):                    # :71
        units = _glue(units, seed_first=False)
        units = _glue(units, seed_first=True)
    return [u.replace(_MARK, '') for u in units]


def syllabify_text(input):
    words = input.split(' ')                                     # :171
    return [syl for w in words for syl in syllabify_word(w)]
