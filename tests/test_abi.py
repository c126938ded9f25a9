"""The C-ABI library loads and exports every symbol include/tanw.h declares (no GPU needed),
and the product path fails loudly -- never falls back to the CPU -- without a device."""
import os
import re

import pytest

import __graft_entry__ as entry
from text_alignment_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def built():
    entry.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'tanw.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tanw_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    declared = _declared_symbols()
    assert declared, 'no declarations found in tanw.h'
    assert sorted(_native.SIGNATURES) == declared


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.tanw_version() >= 100


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, 'text_alignment_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('CPU oracle', ''), f


def test_no_cpu_fallback_without_device():
    try:
        count = _native.device_count()
    except _native.NativeError:
        count = 0
    if count > 0:
        pytest.skip('a GPU is present')
    from text_alignment_b200 import textSeqCompare as tsc
    with pytest.raises(_native.NativeError):
        tsc.perform_alignment(list('abc'), list('abd'))


def test_null_context_is_an_error_not_a_crash():
    import ctypes
    lib = _native.load()
    assert lib.tanw_batch_run(None) == 1                      # TANW_E_INVALID
    assert lib.tanw_sync(None) == 1
    assert lib.tanw_destroy(None) == 0
    assert lib.tanw_set_arena_limit(None, 0) == 1
    t = _native.Timing()
    assert lib.tanw_last_timing(None, ctypes.byref(t)) == 1
    assert b'NULL' in lib.tanw_last_error(None)
    h = ctypes.c_void_p()
    rc = lib.tanw_create(10 ** 6, ctypes.byref(h))            # no such device anywhere
    assert rc == 4 and not h.value                            # TANW_E_NODEVICE
