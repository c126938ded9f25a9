"""Loader of oracle/_ref/textSeqCompare.py -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

``make -C oracle ref`` copies the reference's aligner, byte for byte, from
/root/reference/textSeqCompare.py into the git-ignored ``oracle/_ref/`` so that the UNMODIFIED
reference can run as the CPU arm on the GPU box, where /root/reference does not exist
(``bench.py --impl reference`` and the ``cpu_baseline`` leg; "kind": "reference").  The copy is
checked against the SHA-256 of the file this repository was developed against, so a modified
copy is refused rather than timed.  Its two unused imports (textSeqCompare.py:2-3) resolve to
the empty stub modules under ``_ref/stubs`` when the real packages are not installed.

Only bench.py's CPU legs and tests/ may import this module; the product never does.
"""
import hashlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')
REF_FILE = os.path.join(REF_DIR, 'textSeqCompare.py')
# sha256 of /root/reference/textSeqCompare.py (202 lines) as surveyed
EXPECTED_SHA256 = '1b3d0d41a776ce99b9d480e2a5d5f5536a16495b87c38925f5b2520f9746a37a'

_KEY = '_reference_copy_textSeqCompare'


def available():
    """True when the copy exists and is byte-identical to the surveyed reference file."""
    try:
        with open(REF_FILE, 'rb') as f:
            return hashlib.sha256(f.read()).hexdigest() == EXPECTED_SHA256
    except OSError:
        return False


def load():
    """The reference's own ``textSeqCompare`` module, executed from the copy."""
    if _KEY in sys.modules:
        return sys.modules[_KEY]
    if not available():
        raise RuntimeError('oracle/_ref/textSeqCompare.py is missing or differs from the reference '
                           '(run `make -C oracle ref` where /root/reference exists)')
    stubs = os.path.join(REF_DIR, 'stubs')
    for name in ('unidecode', 'matplotlib.pyplot'):
        try:
            importlib.import_module(name)
        except ImportError:
            if stubs not in sys.path:
                sys.path.append(stubs)          # after everything else: real packages win
            importlib.import_module(name)
    spec = importlib.util.spec_from_file_location(_KEY, REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_KEY] = mod
    spec.loader.exec_module(mod)
    return mod
