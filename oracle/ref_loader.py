"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  The reference is Python 2-era code with two unused imports
(``unidecode``, ``matplotlib.pyplot``; textSeqCompare.py:2-3) that are not installed here, and
``alignToOCR.py`` needs Gamera.  We register stub modules in ``sys.modules`` and leave the
reference files untouched.  /root/reference does not exist on the GPU
box, so nothing that runs there may call this.
"""
import builtins
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get('TEXT_ALIGNMENT_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'textSeqCompare.py'))


def _stub(name):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
    return sys.modules[name]


def load_textseqcompare():
    """Return the reference's own ``textSeqCompare`` module (cached under a private name)."""
    key = '_reference_textSeqCompare'
    if key in sys.modules:
        return sys.modules[key]
    if not available():
        raise RuntimeError('reference not present at %s' % REFERENCE_DIR)
    _stub('unidecode').unidecode = lambda s: s
    mpl = _stub('matplotlib')
    mpl.pyplot = _stub('matplotlib.pyplot')
    spec = importlib.util.spec_from_file_location(key, os.path.join(REFERENCE_DIR, 'textSeqCompare.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


class _ZerosRecorder(object):
    """numpy proxy whose zeros() records the six matrices in call order
    (mat, y_mat, x_mat, mat_ptr, y_mat_ptr, x_mat_ptr; textSeqCompare.py:45-50)."""

    def __init__(self, np):
        self._np = np
        self.arrays = []

    def zeros(self, *a, **k):
        arr = self._np.zeros(*a, **k)
        self.arrays.append(arr)
        return arr

    def __getattr__(self, name):
        return getattr(self._np, name)


def reference_align_full(transcript, ocr, scoring_system=None):
    """Run the reference and also capture its matrices.

    Returns (tra_align, ocr_align, dict(M, Y, X, PM, PY, PX)) with the full (n+1)x(m+1) arrays.
    """
    import numpy as np
    tsc = load_textseqcompare()
    rec = _ZerosRecorder(np)
    saved = tsc.np
    tsc.np = rec
    try:
        tra, ocr_al = tsc.perform_alignment(list(transcript), list(ocr), scoring_system=scoring_system)
    finally:
        tsc.np = saved
    M, Y, X, PM, PY, PX = rec.arrays[:6]
    return tra, ocr_al, dict(M=M, Y=Y, X=X, PM=PM, PY=PY, PX=PX)


def load_aligntoocr():
    """Return the reference's ``alignToOCR`` module with Gamera / OCRopus mocked out
    (SURVEY.md Appendix D recipe 3).  Its ``tsc`` attribute is the reference aligner."""
    key = 'alignToOCR'
    if key in sys.modules and getattr(sys.modules[key], '_is_reference', False):
        return sys.modules[key]
    if not available():
        raise RuntimeError('reference not present at %s' % REFERENCE_DIR)
    from unittest.mock import MagicMock

    class Point(object):
        def __init__(self, x, y):
            self.x = x
            self.y = y

    for name in ('gamera', 'gamera.core', 'gamera.plugins', 'gamera.plugins.image_utilities',
                 'gamera.toolkits', 'matplotlib', 'matplotlib.pyplot', 'unidecode'):
        if name not in sys.modules or not isinstance(sys.modules[name], MagicMock):
            sys.modules[name] = MagicMock()
    sys.modules['gamera.core'].Point = Point
    sys.modules['gamera'].core = sys.modules['gamera.core']
    if not hasattr(builtins, 'reload'):
        builtins.reload = importlib.reload
    if not hasattr(builtins, 'unicode'):
        builtins.unicode = str
    # the reference imports its siblings by bare name
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    for sib in ('textSeqCompare', 'latinSyllabification', 'textAlignPreprocessing', 'alignToOCR'):
        sys.modules.pop(sib, None)
    mod = importlib.import_module('alignToOCR')
    mod._is_reference = True
    return mod
