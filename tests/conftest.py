import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'slow: minutes-long CPU test')


def _device_count():
    try:
        from text_alignment_b200 import _native
        return _native.device_count()
    except Exception:                    # library not built, no driver, ...
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a device: on a box without one they are skipped with a reason instead of
    failing with NativeError (the driver's CPU run deselects them with -m "not gpu" anyway)."""
    gpu_items = [it for it in items if it.get_closest_marker('gpu')]
    if gpu_items and _device_count() == 0:
        skip = pytest.mark.skip(reason='no CUDA device on this box')
        for it in gpu_items:
            it.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def resolve_system(system):
    """Golden fixtures name callables by string (tests/golden/scorers.py)."""
    import scorers
    if system is not None and isinstance(system[0], str):
        return [scorers.SCORERS[system[0]]] + list(system[1:])
    return system


def golden_elems(rec):
    """Inputs of a golden record as two lists of elements."""
    T, O = rec['T'], rec['O']
    return (list(T), list(O))


def ops_string(tra, ocr):
    out = []
    for a, b in zip(tra, ocr):
        if b == '_' and a != '_':
            out.append('1')
        elif a == '_' and b != '_':
            out.append('2')
        else:
            out.append('0')
    return ''.join(out)


@pytest.fixture(scope='session')
def kats():
    return load_golden('kats.json')


@pytest.fixture(scope='session')
def random_pairs():
    return load_golden('random_pairs.json')


@pytest.fixture(scope='session')
def wide_pairs():
    """Pairs with more than 256 distinct elements; T / O are stored as code points."""
    recs = load_golden('wide_pairs.json')
    for r in recs:
        r['T'] = [chr(c) for c in r['T']]
        r['O'] = [chr(c) for c in r['O']]
    return recs


@pytest.fixture(scope='session')
def appendix_c():
    return load_golden('appendix_c.json')
