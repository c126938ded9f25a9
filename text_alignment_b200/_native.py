"""ctypes binding of libtanw.so (include/tanw.h).  There is no CPU fallback: if the library
is missing or no sm_100 device is present, every entry raises."""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtanw.so')

NEG_INF = -1073741824          # TANW_NEG_INF

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


class Scoring(ctypes.Structure):
    _fields_ = [('match', ctypes.c_int32), ('mismatch', ctypes.c_int32),
                ('gap_open_x', ctypes.c_int32), ('gap_open_y', ctypes.c_int32),
                ('gap_extend_x', ctypes.c_int32), ('gap_extend_y', ctypes.c_int32),
                ('boundary_gap', ctypes.c_int32), ('subst_k', ctypes.c_int32),
                ('subst', _i32p)]


class DeviceInfo(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char * 128), ('cc_major', ctypes.c_int32), ('cc_minor', ctypes.c_int32),
                ('sm_count', ctypes.c_int32), ('clock_khz', ctypes.c_int32),
                ('total_mem_bytes', ctypes.c_int64), ('free_mem_bytes', ctypes.c_int64)]


class Timing(ctypes.Structure):
    _fields_ = [('h2d_ms', ctypes.c_float), ('kernel_ms', ctypes.c_float), ('d2h_ms', ctypes.c_float),
                ('kernel_launches', ctypes.c_int32), ('cells', ctypes.c_int64), ('ptr_bytes', ctypes.c_int64),
                ('h2d_bytes', ctypes.c_int64), ('d2h_bytes', ctypes.c_int64),
                ('host_prepare_ms', ctypes.c_float), ('host_run_ms', ctypes.c_float),
                ('host_fetch_ms', ctypes.c_float), ('chunks', ctypes.c_int32), ('table_launches', ctypes.c_int32)]


# every symbol include/tanw.h declares: (restype, argtypes)
_VOIDP = ctypes.c_void_p
SIGNATURES = {
    'tanw_version': (ctypes.c_int, []),
    'tanw_device_count': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    'tanw_device_query': (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(DeviceInfo)]),
    'tanw_last_error': (ctypes.c_char_p, [_VOIDP]),
    'tanw_create': (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_VOIDP)]),
    'tanw_destroy': (ctypes.c_int, [_VOIDP]),
    'tanw_set_arena_limit': (ctypes.c_int, [_VOIDP, ctypes.c_int64]),
    'tanw_set_long_threshold': (ctypes.c_int, [_VOIDP, ctypes.c_int64]),
    'tanw_set_long_band_rows': (ctypes.c_int, [_VOIDP, ctypes.c_int]),
    'tanw_set_symbol_bytes': (ctypes.c_int, [_VOIDP, ctypes.c_int]),
    'tanw_set_line_kernel': (ctypes.c_int, [_VOIDP, ctypes.c_int]),
    'tanw_set_packed_ops': (ctypes.c_int, [_VOIDP, ctypes.c_int]),
    'tanw_align_batch': (ctypes.c_int, [_VOIDP, _u8p, ctypes.c_int64, _i64p, _i32p, _i64p, _i32p, ctypes.c_int64,
                                        ctypes.POINTER(Scoring), _u8p, _i64p, ctypes.c_int64, _i32p, _i32p]),
    'tanw_align_batch_multi': (ctypes.c_int, [_VOIDP, _u8p, ctypes.c_int64, _i64p, _i32p, _i64p, _i32p, ctypes.c_int64,
                                              ctypes.POINTER(Scoring), ctypes.c_int32, _i32p,
                                              _u8p, _i64p, ctypes.c_int64, _i32p, _i32p]),
    'tanw_align_batch_sharded': (ctypes.c_int, [ctypes.POINTER(_VOIDP), ctypes.c_int32, _i64p,
                                                _u8p, ctypes.c_int64, _i64p, _i32p, _i64p, _i32p, ctypes.c_int64,
                                                ctypes.POINTER(Scoring), _u8p, _i64p, ctypes.c_int64, _i32p, _i32p]),
    'tanw_batch_prepare_multi': (ctypes.c_int, [_VOIDP, _u8p, ctypes.c_int64, _i64p, _i32p, _i64p, _i32p, ctypes.c_int64,
                                                ctypes.POINTER(Scoring), ctypes.c_int32, _i32p]),
    'tanw_batch_prepare': (ctypes.c_int, [_VOIDP, _u8p, ctypes.c_int64, _i64p, _i32p, _i64p, _i32p, ctypes.c_int64,
                                          ctypes.POINTER(Scoring)]),
    'tanw_batch_run': (ctypes.c_int, [_VOIDP]),
    'tanw_batch_rescore': (ctypes.c_int, [_VOIDP, ctypes.POINTER(Scoring)]),
    'tanw_batch_fetch': (ctypes.c_int, [_VOIDP, _u8p, _i64p, ctypes.c_int64, _i32p, _i32p]),
    'tanw_sync': (ctypes.c_int, [_VOIDP]),
    'tanw_last_timing': (ctypes.c_int, [_VOIDP, ctypes.POINTER(Timing)]),
    'tanw_stream_handle': (ctypes.c_int, [_VOIDP, ctypes.POINTER(ctypes.c_uint64)]),
    'tanw_consumer_last_error': (ctypes.c_char_p, []),
    'tanw_syllabify_text': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64, _i32p, ctypes.c_int64, _i64p]),
    'tanw_parse_llocs': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.POINTER(ctypes.c_uint32), _i32p, ctypes.c_int64, _i64p]),
    'tanw_syllable_boxes': (ctypes.c_int, [ctypes.c_int64, _u8p, _i64p, _i32p, _i32p, _i64p, _i32p, _i64p, _i32p, _u8p]),
    'tanw_boxes_to_json': (ctypes.c_int, [ctypes.c_char_p, _i64p, ctypes.c_int64, _i32p, _u8p, ctypes.c_char_p,
                                          ctypes.c_char_p, ctypes.c_int64, _i64p]),
    'tanw_measure_int32_peak': (ctypes.c_int, [_VOIDP, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
}

_lib = None
_lib_lock = threading.Lock()


class NativeError(RuntimeError):
    pass


def load(path=None):
    """Load libtanw.so (built in-tree by __graft_entry__.build()); fail loudly if absent.
    `path`: another build of the library (tools/: kernel variants, the TANW_CHECKED build); it must
    be given before anything else has loaded the library."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is not None and os.path.abspath(path) != _lib._name:
            raise NativeError('libtanw is already loaded from %s' % _lib._name)
        if _lib is None:
            path = os.path.abspath(path) if path else LIB_PATH
            if not os.path.exists(path):
                raise NativeError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                                  '(there is no CPU fallback)' % path)
            lib = ctypes.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)      # AttributeError if the symbol is missing
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def _ptr(a, ct):
    return a.ctypes.data_as(ct)


_pylist = None


def pylist():
    """The CPython-side list marshalling helper (csrc/tanw_pylist.c), or None when it was not
    built: it only speeds up list <-> code conversion, the numpy path gives the same results."""
    global _pylist
    if _pylist is None:
        path = os.path.join(_HERE, '_tanw_pylist.so')
        try:
            lib = ctypes.PyDLL(path)
            lib.tanw_pylist_codepoints.restype = ctypes.c_ssize_t
            lib.tanw_pylist_codepoints.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_ssize_t, ctypes.c_void_p]
            lib.tanw_pylist_expand.restype = ctypes.py_object
            lib.tanw_pylist_expand.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_ssize_t, ctypes.c_int,
                                               ctypes.py_object]
            lib.tanw_pylist_slices.restype = ctypes.py_object
            lib.tanw_pylist_slices.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_ssize_t, ctypes.c_void_p]
            _pylist = lib
        except (OSError, AttributeError):
            _pylist = False
    return _pylist or None


def device_count():
    c = ctypes.c_int(0)
    rc = load().tanw_device_count(ctypes.byref(c))
    if rc:
        raise NativeError(load().tanw_last_error(None).decode())
    return c.value


def device_info(device=0):
    info = DeviceInfo()
    rc = load().tanw_device_query(device, ctypes.byref(info))
    if rc:
        raise NativeError(load().tanw_last_error(None).decode())
    return dict(name=info.name.decode(), cc=(info.cc_major, info.cc_minor), sm_count=info.sm_count,
                clock_khz=info.clock_khz, total_mem_bytes=info.total_mem_bytes,
                free_mem_bytes=info.free_mem_bytes)


def _consumer_check(rc):
    if rc:
        msg = load().tanw_consumer_last_error().decode()
        if rc == 5:
            raise MemoryError(msg)
        if 'all_chars not same length' in msg:
            raise AssertionError(msg)          # the reference's own assertion (alignToOCR.py:291)
        raise ValueError(msg)


def text_slices(text, bounds, keep=None):
    """[text[a:b] for a, b in bounds] (only where ``keep`` is set, if given) -- through the CPython
    helper when it is built, else in Python."""
    bounds = np.ascontiguousarray(bounds, dtype=np.int32).reshape(-1, 2)
    lib = pylist()
    if lib is not None and type(text) is str:
        k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        return lib.tanw_pylist_slices(text, bounds.ctypes.data, bounds.shape[0], None if k is None else k.ctypes.data)
    rows = zip(bounds[:, 0].tolist(), bounds[:, 1].tolist())
    if keep is None:
        return [text[a:b] for a, b in rows]
    return [text[a:b] for (a, b), h in zip(rows, np.asarray(keep).tolist()) if h]


def syllable_bounds(text):
    """Syllables of an ASCII transcript as int32[S, 2] character ranges, natively; None when the
    text holds anything but ASCII letters, digits and spaces (the caller then uses the Python
    syllabifier); ValueError for a word without a vowel, as the Python syllabifier."""
    lib = load()
    try:
        raw = text.encode('ascii')
    except UnicodeEncodeError:
        return None
    cap = len(raw) + 1
    bounds = np.empty((cap, 2), dtype=np.int32)
    k = ctypes.c_int64(0)
    rc = lib.tanw_syllabify_text(raw, len(raw), _ptr(bounds, _i32p), cap, ctypes.byref(k))
    if rc == 6:
        return None
    _consumer_check(rc)
    return bounds[:k.value]


def parse_llocs(text, x_min, y_min, y_max):
    """One .llocs text line (bytes) -> (code points uint32[k], boxes int32[k, 4])."""
    lib = load()
    cap = text.count(b'\n') + 2
    chars = np.empty(cap, dtype=np.uint32)
    boxes = np.empty((cap, 4), dtype=np.int32)
    k = ctypes.c_int64(0)
    _consumer_check(lib.tanw_parse_llocs(text, len(text), int(x_min), int(y_min), int(y_max),
                                         chars.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), _ptr(boxes, _i32p), cap,
                                         ctypes.byref(k)))
    return chars[:k.value], boxes[:k.value]


def syllable_boxes(ops, ops_off, ops_len, syl_bounds, syl_off, boxes, box_off):
    """Many pages at once -> (boxes int32[S, 4], has_box bool[S]) for all S syllables."""
    lib = load()
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    ops_off = np.ascontiguousarray(ops_off, dtype=np.int64)
    ops_len = np.ascontiguousarray(ops_len, dtype=np.int32)
    syl_bounds = np.ascontiguousarray(syl_bounds, dtype=np.int32)
    syl_off = np.ascontiguousarray(syl_off, dtype=np.int64)
    boxes = np.ascontiguousarray(boxes, dtype=np.int32)
    box_off = np.ascontiguousarray(box_off, dtype=np.int64)
    S = int(syl_off[-1]) if syl_off.size else 0
    out = np.zeros((max(S, 1), 4), dtype=np.int32)
    has = np.zeros(max(S, 1), dtype=np.uint8)
    _consumer_check(lib.tanw_syllable_boxes(int(ops_len.size), _ptr(ops, _u8p), _ptr(ops_off, _i64p), _ptr(ops_len, _i32p),
                                            _ptr(syl_bounds, _i32p), _ptr(syl_off, _i64p), _ptr(boxes, _i32p),
                                            _ptr(box_off, _i64p), _ptr(out, _i32p), _ptr(has, _u8p)))
    return out[:S], has[:S].astype(bool)


def boxes_to_json(syllables, syl_boxes, has_box, median_line_spacing):
    """JSON bytes of alignToOCR.to_JSON_dict, written natively (no dict, no per-box objects)."""
    lib = load()
    enc = [s.encode('utf-8') for s in syllables]
    text = b''.join(enc)
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        np.cumsum(np.fromiter(map(len, enc), dtype=np.int64, count=len(enc)), out=off[1:])
    syl_boxes = np.ascontiguousarray(syl_boxes, dtype=np.int32)
    has = np.ascontiguousarray(has_box, dtype=np.uint8)
    cap = 96 * int(has.sum()) + 3 * len(text) + 128
    buf = ctypes.create_string_buffer(cap)
    k = ctypes.c_int64(0)
    _consumer_check(lib.tanw_boxes_to_json(text, _ptr(off, _i64p), len(enc), _ptr(syl_boxes, _i32p), _ptr(has, _u8p),
                                           repr(float(median_line_spacing)).encode(), buf, cap, ctypes.byref(k)))
    return buf.raw[:k.value]


class Context(object):
    """One per GPU.  The native context is single-caller (include/tanw.h), and ctypes releases
    the GIL for the duration of each native call, so every entry that touches the batch state
    holds ``self.lock`` (re-entrant): two Python threads may share a Context and are serialised,
    distinct contexts run concurrently.  Callers of the three-phase form (prepare / rescore / run
    / fetch) that share a Context between threads hold ``ctx.lock`` across the whole sequence."""

    def __init__(self, device=0):
        self.lock = threading.RLock()
        self._lib = load()
        h = _VOIDP()
        rc = self._lib.tanw_create(int(device), ctypes.byref(h))
        if rc:
            raise NativeError('tanw_create(device=%d) failed: %s' % (device, self._lib.tanw_last_error(None).decode()))
        self._h = h
        self.device = int(device)
        self._keep = None
        self._sym_bytes = 1
        self._packed = False

    def close(self):
        if getattr(self, '_h', None):
            self._lib.tanw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            msg = self._lib.tanw_last_error(self._h).decode()
            if rc == 2:
                raise OverflowError(msg)
            if rc == 1:
                raise ValueError(msg)
            if rc == 5:
                raise MemoryError(msg)
            if rc == 7:
                raise AssertionError(msg)
            raise NativeError('libtanw error %d: %s' % (rc, msg))

    def set_arena_limit(self, nbytes):
        self._check(self._lib.tanw_set_arena_limit(self._h, int(nbytes)))

    def set_long_threshold(self, cells):
        """Pairs with n*m >= cells use the chained-pass (whole-GPU) path; default 2**26."""
        self._check(self._lib.tanw_set_long_threshold(self._h, int(cells)))

    def set_long_band_rows(self, rows):
        """Cut chained-pass pairs into bands of `rows` rows (checkpoint + recompute); 0 = only
        when the pointer block would not fit the arena."""
        self._check(self._lib.tanw_set_long_band_rows(self._h, int(rows)))

    def _symbol_width(self, symbols):
        """Tell the library whether `symbols` holds uint8 or uint16 codes (only when it changes)."""
        width = symbols.dtype.itemsize
        if width != self._sym_bytes:
            self._check(self._lib.tanw_set_symbol_bytes(self._h, width))
            self._sym_bytes = width

    def set_line_kernel(self, mode):
        """Route of short pairs (m <= 128): 1 / True = the line kernels (two pairs per register when
        the scores fit 16 bits, else four pairs per warp in int32; default), 2 = the int32 line
        kernel only, 0 / False = the page kernel."""
        self._check(self._lib.tanw_set_line_kernel(self._h, int(mode)))

    def set_packed_ops(self, enabled):
        """Deliver op strings four to a byte (include/tanw.h, tanw_set_packed_ops); ``align_batch``
        then returns the packed buffer, ``packed_layout`` / ``unpack_ops`` give the byte view back."""
        with self.lock:
            self._check(self._lib.tanw_set_packed_ops(self._h, 1 if enabled else 0))
            self._packed = bool(enabled)

    @staticmethod
    def packed_layout(n, m):
        """(packed offsets, packed total bytes) of a batch: (ops_off >> 2) + p, sum(n+m)/4 + P + 1."""
        off, total = Context.canonical_ops_layout(n, m)
        return (off >> 2) + np.arange(n.size, dtype=np.int64), total // 4 + int(n.size) + 1

    @staticmethod
    def unpack_ops(packed, n, m, ops_len):
        """Packed op strings -> (ops, ops_off) in the canonical one-byte-per-op layout."""
        off, total = Context.canonical_ops_layout(n, m)
        poff, _ = Context.packed_layout(n, m)
        ops = np.zeros(max(total, 1), dtype=np.uint8)
        P = int(n.size)
        if P == 0:
            return ops, off
        lens = ops_len[:P].astype(np.int64)
        tot = int(lens.sum())
        start = np.cumsum(lens) - lens
        pair = np.repeat(np.arange(P), lens)
        q = np.arange(tot, dtype=np.int64) - np.repeat(start, lens)
        src = packed[np.repeat(poff, lens) + (q >> 2)]
        ops[np.repeat(off, lens) + q] = (src >> ((q & 3) * 2).astype(np.uint8)) & 3
        del pair
        return ops, off

    @staticmethod
    def _canon(symbols, t_off, n, o_off, m):
        # uint16 codes stay 16 bits wide (pairs with more than 256 distinct elements); anything else is bytes
        wide = isinstance(symbols, np.ndarray) and symbols.dtype == np.uint16
        symbols = np.ascontiguousarray(symbols, dtype=np.uint16 if wide else np.uint8)
        t_off = np.ascontiguousarray(t_off, dtype=np.int64)
        o_off = np.ascontiguousarray(o_off, dtype=np.int64)
        n = np.ascontiguousarray(n, dtype=np.int32)
        m = np.ascontiguousarray(m, dtype=np.int32)
        if not (t_off.size == n.size == o_off.size == m.size):
            raise ValueError('pair table arrays differ in length')
        return symbols, t_off, n, o_off, m

    @staticmethod
    def make_scoring(match, mismatch, gox, goy, gex, gey, boundary_gap, subst=None):
        sc = Scoring()
        sc.match, sc.mismatch = int(match), int(mismatch)
        sc.gap_open_x, sc.gap_open_y = int(gox), int(goy)
        sc.gap_extend_x, sc.gap_extend_y = int(gex), int(gey)
        sc.boundary_gap = int(boundary_gap)
        keep = None
        if subst is not None:
            keep = np.ascontiguousarray(subst, dtype=np.int32)
            if keep.ndim != 2 or keep.shape[0] != keep.shape[1]:
                raise ValueError('substitution table must be square')
            sc.subst_k = keep.shape[0]
            sc.subst = _ptr(keep, _i32p)
        return sc, keep

    @staticmethod
    def canonical_ops_layout(n, m):
        cap = n.astype(np.int64) + m.astype(np.int64)
        off = np.zeros(n.size, dtype=np.int64)
        if n.size:
            np.cumsum(cap[:-1], out=off[1:])
        return off, int(cap.sum())

    def align_batch(self, symbols, t_off, n, o_off, m, scoring, want_scores=True, out=None, layout=None):
        """One call = H2D + fill + traceback + D2H.  Returns (ops, ops_off, ops_len, scores).
        `out` = (ops, ops_len, scores) preallocated arrays (e.g. pinned) to receive the results;
        `layout` = canonical_ops_layout(n, m) computed earlier by a caller that reuses its output
        buffers (the prefix sums are host work of the same order as the call itself for 10^5
        short pairs)."""
        symbols, t_off, n, o_off, m = self._canon(symbols, t_off, n, o_off, m)
        sc, keep = scoring
        P = int(n.size)
        ops_off, total = layout if layout is not None else self.canonical_ops_layout(n, m)
        if self._packed:
            total = total // 4 + P + 1           # the packed buffer (ops_off stays the canonical byte layout)
        if out is not None:
            ops, ops_len, scores = out
            if (ops.size < total and total > 0) or ops_len.size < P or (want_scores and scores.size < 3 * P):
                raise ValueError('preallocated output buffers are too small')
        else:
            ops = np.empty(max(total, 1), dtype=np.uint8)
            ops_len = np.zeros(max(P, 1), dtype=np.int32)
            scores = np.zeros((max(P, 1), 3), dtype=np.int32) if want_scores else None
        with self.lock:
            self._symbol_width(symbols)
            rc = self._lib.tanw_align_batch(self._h, symbols.ctypes.data_as(_u8p), symbols.size, _ptr(t_off, _i64p),
                                            _ptr(n, _i32p), _ptr(o_off, _i64p), _ptr(m, _i32p), P, ctypes.byref(sc),
                                            _ptr(ops, _u8p), _ptr(ops_off, _i64p), ops.size, _ptr(ops_len, _i32p),
                                            _ptr(scores, _i32p) if want_scores else None)
            self._check(rc)
        return ops, ops_off, ops_len[:P], (scores[:P] if want_scores else None)

    @staticmethod
    def align_batch_sharded(contexts, bounds, symbols, t_off, n, o_off, m, scoring_params, subst=None,
                            want_scores=True, out=None, layout=None):
        """One batch over several contexts (devices): shard d = pairs bounds[d]..bounds[d+1]-1 on
        contexts[d], one native host thread per shard (tanw_align_batch_sharded), results written
        straight into the batch's arrays.  Returns (ops, ops_off, ops_len, scores)."""
        first = contexts[0]
        symbols, t_off, n, o_off, m = first._canon(symbols, t_off, n, o_off, m)
        bounds = np.ascontiguousarray(bounds, dtype=np.int64)
        if bounds.size != len(contexts) + 1:
            raise ValueError('bounds must have one more entry than there are contexts')
        P = int(n.size)
        ops_off, total = layout if layout is not None else Context.canonical_ops_layout(n, m)
        if out is not None:
            ops, ops_len, scores = out
            if (ops.size < total and total > 0) or ops_len.size < P or (want_scores and scores.size < 3 * P):
                raise ValueError('preallocated output buffers are too small')
        else:
            ops = np.empty(max(total, 1), dtype=np.uint8)
            ops_len = np.zeros(max(P, 1), dtype=np.int32)
            scores = np.zeros((max(P, 1), 3), dtype=np.int32) if want_scores else None
        sc, keep = first.make_scoring(*scoring_params, subst=subst)
        handles = (_VOIDP * len(contexts))(*[c._h for c in contexts])
        locks = sorted(set(contexts), key=id)
        for c in locks:
            c.lock.acquire()
        try:
            for c in contexts:
                c._symbol_width(symbols)
            rc = first._lib.tanw_align_batch_sharded(handles, len(contexts), _ptr(bounds, _i64p),
                                                     symbols.ctypes.data_as(_u8p), symbols.size, _ptr(t_off, _i64p),
                                                     _ptr(n, _i32p), _ptr(o_off, _i64p), _ptr(m, _i32p), P,
                                                     ctypes.byref(sc), _ptr(ops, _u8p), _ptr(ops_off, _i64p), ops.size,
                                                     _ptr(ops_len, _i32p), _ptr(scores, _i32p) if want_scores else None)
            first._check(rc)
        finally:
            for c in reversed(locks):
                c.lock.release()
        return ops, ops_off, ops_len[:P], (scores.reshape(-1, 3)[:P] if want_scores else None)

    def align_batch_multi(self, symbols, t_off, n, o_off, m, scorings, scoring_idx, want_scores=True):
        """Every pair under its own scoring system (the reference's parameter sweep as one launch):
        ``scorings`` = list of (match, mismatch, gox, goy, gex, gey, boundary_gap) tuples,
        ``scoring_idx[p]`` = which of them pair p uses.  Returns (ops, ops_off, ops_len, scores)."""
        symbols, t_off, n, o_off, m = self._canon(symbols, t_off, n, o_off, m)
        if symbols.dtype != np.uint8:
            raise ValueError('per-pair scoring systems need 8-bit symbol codes')
        sidx = np.ascontiguousarray(scoring_idx, dtype=np.int32)
        if sidx.size != n.size:
            raise ValueError('scoring_idx and the pair table differ in length')
        arr = (Scoring * max(len(scorings), 1))()
        for k, prm in enumerate(scorings):
            sc, _ = self.make_scoring(*prm)
            arr[k] = sc
        P = int(n.size)
        ops_off, total = self.canonical_ops_layout(n, m)
        ops = np.empty(max(total, 1), dtype=np.uint8)
        ops_len = np.zeros(max(P, 1), dtype=np.int32)
        scores = np.zeros((max(P, 1), 3), dtype=np.int32) if want_scores else None
        with self.lock:
            self._symbol_width(symbols)
            rc = self._lib.tanw_align_batch_multi(self._h, symbols.ctypes.data_as(_u8p), symbols.size, _ptr(t_off, _i64p),
                                                  _ptr(n, _i32p), _ptr(o_off, _i64p), _ptr(m, _i32p), P, arr,
                                                  len(scorings), _ptr(sidx, _i32p), _ptr(ops, _u8p),
                                                  _ptr(ops_off, _i64p), ops.size, _ptr(ops_len, _i32p),
                                                  _ptr(scores, _i32p) if want_scores else None)
            self._check(rc)
        return ops, ops_off, ops_len[:P], (scores[:P] if want_scores else None)

    # ---- three-phase form (bench.py times run() alone with the inputs resident in HBM) ----
    def prepare(self, symbols, t_off, n, o_off, m, scoring):
        symbols, t_off, n, o_off, m = self._canon(symbols, t_off, n, o_off, m)
        sc, keep = scoring
        with self.lock:
            self._symbol_width(symbols)
            self._keep = (symbols, t_off, n, o_off, m, sc, keep)
            self._check(self._lib.tanw_batch_prepare(self._h, symbols.ctypes.data_as(_u8p), symbols.size,
                                                     _ptr(t_off, _i64p), _ptr(n, _i32p), _ptr(o_off, _i64p),
                                                     _ptr(m, _i32p), int(n.size), ctypes.byref(sc)))

    def rescore(self, scoring):
        """New scoring system for the prepared batch (sequences stay resident in HBM)."""
        sc, keep = scoring
        with self.lock:
            self._keep = self._keep[:5] + (sc, keep)
            self._check(self._lib.tanw_batch_rescore(self._h, ctypes.byref(sc)))

    def run(self):
        with self.lock:
            self._check(self._lib.tanw_batch_run(self._h))

    def sync(self):
        with self.lock:
            self._check(self._lib.tanw_sync(self._h))

    def fetch(self, want_scores=True):
        _, _, n, _, m, _, _ = self._keep
        P = int(n.size)
        ops_off, total = self.canonical_ops_layout(n, m)
        ops = np.empty(max(total, 1), dtype=np.uint8)
        ops_len = np.zeros(max(P, 1), dtype=np.int32)
        scores = np.zeros((max(P, 1), 3), dtype=np.int32) if want_scores else None
        self.fetch_into(ops_off, (ops, ops_len, scores))
        return ops, ops_off, ops_len[:P], (scores[:P] if want_scores else None)

    def fetch_into(self, ops_off, out):
        """Results of the batch that ran into caller-owned (e.g. pinned) arrays
        ``out = (ops, ops_len, scores or None)``; ``ops_off`` as from canonical_ops_layout."""
        ops, ops_len, scores = out
        with self.lock:
            self._check(self._lib.tanw_batch_fetch(self._h, _ptr(ops, _u8p), _ptr(ops_off, _i64p), ops.size,
                                                   _ptr(ops_len, _i32p),
                                                   _ptr(scores, _i32p) if scores is not None else None))

    def timing(self):
        t = Timing()
        self._check(self._lib.tanw_last_timing(self._h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def stream_handle(self):
        v = ctypes.c_uint64(0)
        self._check(self._lib.tanw_stream_handle(self._h, ctypes.byref(v)))
        return v.value

    def measure_int32_peak(self, which=0):
        v = ctypes.c_double(0.0)
        self._check(self._lib.tanw_measure_int32_peak(self._h, int(which), ctypes.byref(v)))
        return v.value
