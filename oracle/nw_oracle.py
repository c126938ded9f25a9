"""ctypes front-end of oracle/nw_oracle.c -- TEST INFRASTRUCTURE ONLY.

Builds ``libnw_oracle.so`` with gcc on first use (``make -C oracle``).  Inputs are arbitrary
Python sequences; they are interned to uint8 codes per pair (equal elements -> equal codes),
and a callable scorer (textSeqCompare.py:27-29) is tabulated into a K x K float64 table.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Scoring(ctypes.Structure):
    _fields_ = [('match', ctypes.c_double), ('mismatch', ctypes.c_double),
                ('subst', ctypes.POINTER(ctypes.c_double)), ('k', ctypes.c_int),
                ('gox', ctypes.c_double), ('goy', ctypes.c_double),
                ('gex', ctypes.c_double), ('gey', ctypes.c_double),
                ('bgap', ctypes.c_double)]


def build(force=False):
    so = os.path.join(_HERE, 'libnw_oracle.so')
    src = os.path.join(_HERE, 'nw_oracle.c')
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libnw_oracle.so'],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.nwo_align.restype = ctypes.c_int
        L.nwo_align.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.POINTER(Scoring),
                                u8p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double), u8p]
        L.nwo_align_batch.restype = ctypes.c_int
        L.nwo_align_batch.argtypes = [u8p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.c_int64, ctypes.POINTER(Scoring), u8p,
                                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_double), ctypes.c_int]
        _LIB = L
    return _LIB


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def make_scoring(scoring_system, symbols=None, boundary_gap=-1):
    """(Scoring struct, keepalive) from a reference-style scoring system (textSeqCompare.py:24-42).
    `symbols`: list of the distinct elements in code order, needed for a callable scorer."""
    ss = [8, -4, -7, -7, -3, 0] if scoring_system is None else scoring_system
    sc = Scoring()
    keep = None
    if len(ss) == 5 and callable(ss[0]):
        k = len(symbols)
        tab = np.zeros((max(k, 1), max(k, 1)), dtype=np.float64)
        for a in range(k):
            for b in range(k):
                tab[a, b] = ss[0](symbols[a], symbols[b])
        keep = tab
        sc.subst = _ptr(tab, ctypes.c_double); sc.k = max(k, 1)
        sc.gox, sc.goy, sc.gex, sc.gey = [float(v) for v in ss[1:5]]
    elif len(ss) == 6:
        sc.match, sc.mismatch = float(ss[0]), float(ss[1])
        sc.gox, sc.goy, sc.gex, sc.gey = [float(v) for v in ss[2:6]]
    elif len(ss) == 4:
        sc.match, sc.mismatch = float(ss[0]), float(ss[1])
        sc.gox = sc.goy = float(ss[2]); sc.gex = sc.gey = float(ss[3])
    else:
        raise ValueError('scoring_system {} invalid'.format(ss))
    sc.bgap = float(boundary_gap)
    return sc, keep


def intern_pair(T, O):
    table = {}
    syms = []

    def code(e):
        c = table.get(e)
        if c is None:
            c = table[e] = len(syms)
            syms.append(e)
        return c
    tc = np.fromiter((code(e) for e in T), dtype=np.int64, count=len(T))
    oc = np.fromiter((code(e) for e in O), dtype=np.int64, count=len(O))
    if len(syms) > 256:
        raise ValueError('oracle handles at most 256 distinct symbols per pair')
    return tc.astype(np.uint8), oc.astype(np.uint8), syms


def align_codes(tc, oc, sc, want_ptr=False):
    """tc/oc uint8 arrays -> (ops uint8[L], (M,X,Y) end scores as floats, ptr or None)."""
    n, m = int(tc.size), int(oc.size)
    tc = np.ascontiguousarray(tc, dtype=np.uint8); oc = np.ascontiguousarray(oc, dtype=np.uint8)
    ops = np.zeros(max(n + m, 1), dtype=np.uint8)
    L = ctypes.c_int(0)
    end3 = np.zeros(3, dtype=np.float64)
    ptr = np.zeros((n, m), dtype=np.uint8) if want_ptr else None
    rc = lib().nwo_align(_ptr(tc, ctypes.c_uint8), n, _ptr(oc, ctypes.c_uint8), m, ctypes.byref(sc),
                         _ptr(ops, ctypes.c_uint8), ctypes.byref(L), _ptr(end3, ctypes.c_double),
                         _ptr(ptr, ctypes.c_uint8) if want_ptr and ptr.size else None)
    if rc:
        raise MemoryError('nwo_align failed')
    return ops[:L.value].copy(), tuple(end3.tolist()), ptr


def perform_alignment(transcript, ocr, scoring_system=None, boundary_gap=-1, full=False):
    """Reference-shaped entry (lists in, two lists out) computed by the C oracle."""
    tc, oc, syms = intern_pair(transcript, ocr)
    sc, keep = make_scoring(scoring_system, syms, boundary_gap)
    ops, end3, ptr = align_codes(tc, oc, sc, want_ptr=full)
    tra, oc_al = [], []
    x = y = 0
    for op in ops.tolist():
        if op == 0:
            tra.append(transcript[x]); oc_al.append(ocr[y]); x += 1; y += 1
        elif op == 1:
            tra.append(transcript[x]); oc_al.append('_'); x += 1
        else:
            tra.append('_'); oc_al.append(ocr[y]); y += 1
    if full:
        return tra, oc_al, dict(ops=ops, end=end3, ptr=ptr)
    return tra, oc_al


def align_batch_codes(sym, t_off, n, o_off, m, sc, threads=1, want_scores=True):
    """Packed batch (same layout as the C-ABI of the product) on `threads` host threads."""
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    t_off = np.ascontiguousarray(t_off, dtype=np.int64); o_off = np.ascontiguousarray(o_off, dtype=np.int64)
    n = np.ascontiguousarray(n, dtype=np.int32); m = np.ascontiguousarray(m, dtype=np.int32)
    P = int(n.size)
    cap = n.astype(np.int64) + m.astype(np.int64)
    ops_off = np.zeros(P, dtype=np.int64)
    if P:
        ops_off[1:] = np.cumsum(cap)[:-1]
    ops = np.zeros(max(int(cap.sum()), 1), dtype=np.uint8)
    ops_len = np.zeros(max(P, 1), dtype=np.int32)
    end3 = np.zeros((max(P, 1), 3), dtype=np.float64)
    rc = lib().nwo_align_batch(_ptr(sym, ctypes.c_uint8), _ptr(t_off, ctypes.c_int64), _ptr(n, ctypes.c_int32),
                               _ptr(o_off, ctypes.c_int64), _ptr(m, ctypes.c_int32), P, ctypes.byref(sc),
                               _ptr(ops, ctypes.c_uint8), _ptr(ops_off, ctypes.c_int64),
                               _ptr(ops_len, ctypes.c_int32),
                               _ptr(end3, ctypes.c_double) if want_scores else None, int(threads))
    if rc:
        raise MemoryError('nwo_align_batch failed')
    return ops, ops_off, ops_len[:P], end3[:P]
