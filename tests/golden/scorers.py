"""Named callable scorers used by the golden vectors (form [f(a,b), gox, goy, gex, gey],
textSeqCompare.py:27-29).  Shared by make_golden.py and the tests so that a fixture can
refer to a callable by name."""

VOWELS = set('aeiouy')


def vowel_aware(a, b):
    if a == b:
        return 9
    if a in VOWELS and b in VOWELS:
        return -1
    return -6


def confusable(a, b):
    if a == b:
        return 6
    pair = {a, b}
    for grp in ('il1', 'un', 'ce', 'rn', 'vy'):
        if pair <= set(grp):
            return 2
    return -5


def asymmetric(a, b):
    # deliberately not symmetric and with a positive mismatch for one ordering
    if a == b:
        return 4
    return 1 if a < b else -3


SCORERS = {'vowel_aware': vowel_aware, 'confusable': confusable, 'asymmetric': asymmetric}
