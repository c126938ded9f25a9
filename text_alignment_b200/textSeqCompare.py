"""Drop-in for DDMAL/text_alignment's ``textSeqCompare`` module, computed on B200.

Mirrors /root/reference/textSeqCompare.py:

* ``perform_alignment(transcript, ocr, scoring_system=None, verbose=False)`` keeps the
  signature, argument meaning, return value ``(tra_align, ocr_align)``, ``ValueError`` text
  (:42) and ``verbose`` print format (:172-175) of the reference (:13-177);
* the module attributes ``default_match, default_mismatch, gap_open, gap_extend,
  default_sys`` (:6-10) exist, are writable and are read at call time -- ``default_sys`` at
  :24-25 and ``gap_extend`` for the boundary rows at :54-59.

The O(n*m) work (boundary, fill, traceback; :45-164) runs in hand-written CUDA behind the C ABI
of ``libtanw.so`` (include/tanw.h); this file only converts Python lists to packed uint8 codes
and op strings back to lists.  There is no CPU fallback.

Additive API (not in the reference): ``perform_alignment_batch`` for many pairs per launch
and ``align_packed`` for callers that already hold packed buffers.
"""
import threading

import numpy as np

from . import _native

# scoring system (textSeqCompare.py:5-10)
default_match = 10
default_mismatch = -5
gap_open = -10
gap_extend = -1
default_sys = [8, -4, -7, -7, -3, 0]

GAP = '_'            # gap symbol of the reference (:130, :139)

_contexts = {}
_contexts_lock = threading.Lock()


def get_context(device=0):
    """Process-wide native context of a device (created on first use)."""
    with _contexts_lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _contexts[device] = _native.Context(device)
        return ctx


def close_contexts():
    with _contexts_lock:
        for ctx in _contexts.values():
            ctx.close()
        _contexts.clear()


# ---- scoring-system parsing (textSeqCompare.py:24-42) -----------------------------------------

def _as_int(v, what):
    """The device path is int32 fixed point.  Every scoring system the reference ships or
    sweeps is integral (default_sys :10; the grid of evaluate_text_alignment.py:181-188)."""
    if isinstance(v, (bool, np.bool_)):
        return int(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)) and float(v).is_integer():
        return int(v)
    raise TypeError('{} = {!r}: only integral scoring values are supported by the device path'.format(what, v))


def parse_scoring_system(scoring_system):
    """-> (callable or None, match, mismatch, gox, goy, gex, gey); raises the reference's
    ValueError for any other form (textSeqCompare.py:41-42)."""
    if scoring_system is None:
        scoring_system = default_sys                                   # :24-25
    if len(scoring_system) == 5 and callable(scoring_system[0]):      # :27-29
        gox, goy, gex, gey = scoring_system[-4:]
        return (scoring_system[0], 0, 0, _as_int(gox, 'gap_open_x'), _as_int(goy, 'gap_open_y'),
                _as_int(gex, 'gap_extend_x'), _as_int(gey, 'gap_extend_y'))
    elif len(scoring_system) == 6:                                     # :30-34
        gox, goy, gex, gey = scoring_system[-4:]
        return (None, _as_int(scoring_system[0], 'match'), _as_int(scoring_system[1], 'mismatch'),
                _as_int(gox, 'gap_open_x'), _as_int(goy, 'gap_open_y'),
                _as_int(gex, 'gap_extend_x'), _as_int(gey, 'gap_extend_y'))
    elif len(scoring_system) == 4:                                     # :35-40
        go = _as_int(scoring_system[2], 'gap_open')
        ge = _as_int(scoring_system[3], 'gap_extend')
        return (None, _as_int(scoring_system[0], 'match'), _as_int(scoring_system[1], 'mismatch'),
                go, go, ge, ge)
    raise ValueError('scoring_system {} invalid'.format(scoring_system))   # :41-42


# ---- list <-> packed code conversion ------------------------------------------------------------

class _Encoded(object):
    """One pair as uint8 codes (uint16 when it has more than 256 distinct elements).
    ``symbols``: distinct elements in code order (None when the codes are the elements' own code
    points, which needs no table)."""
    __slots__ = ('t_codes', 'o_codes', 'symbols', 't_cp', 'o_cp', 'reflexive')


def _code_points(seq):
    """uint32 code points of a list of 1-character strings, or None if the list holds anything
    else (other types, longer or empty strings).  Everything runs at C speed: joining with a
    NUL separator fails for non-strings, and the result has the separators at exactly the odd
    positions iff every element is one character long."""
    k = len(seq)
    if k == 0:
        return np.zeros(0, dtype=np.uint32)
    try:
        joined = '\x00'.join(seq)
    except TypeError:
        return None
    chars = joined[0::2]
    if len(joined) != 2 * k - 1 or joined[1::2] != '\x00' * (k - 1) or '\x00' in chars:
        return None
    try:
        return np.frombuffer(chars.encode('utf-32-le', 'surrogatepass'), dtype=np.uint32)
    except UnicodeError:
        return None


MAX_SYMBOLS = 65536          # distinct elements per pair (uint16 codes)
MAX_TABLE_SYMBOLS = 2048     # ... when the scorer is a callable (K x K table, include/tanw.h)


def _code_dtype(distinct, tabulated):
    """uint8 codes while they suffice; uint16 for pairs with more distinct elements (the device
    then runs them on the page kernel only, tanw_set_symbol_bytes)."""
    if distinct <= 256:
        return np.uint8
    limit = MAX_TABLE_SYMBOLS if tabulated else MAX_SYMBOLS
    if distinct > limit:
        raise ValueError('more than {} distinct symbols in one pair ({}); the device path cannot '
                         'represent them'.format(limit, distinct))
    return np.uint16


def _encode_pair(transcript, ocr, need_dense):
    enc = _Encoded()
    enc.reflexive = True
    enc.t_cp = enc.o_cp = None
    # production case (alignToOCR.py:273: list(transcript), list(ocr)): single characters
    t_cp = _code_points(transcript)
    o_cp = _code_points(ocr) if t_cp is not None else None
    if o_cp is not None:
        enc.t_cp, enc.o_cp = t_cp, o_cp
        top = int(max(t_cp.max() if t_cp.size else 0, o_cp.max() if o_cp.size else 0))
        if top < 256 and not need_dense:
            enc.t_codes = t_cp.astype(np.uint8)
            enc.o_codes = o_cp.astype(np.uint8)
            enc.symbols = None
            return enc
        both = np.concatenate([t_cp, o_cp])
        uniq, inv = np.unique(both, return_inverse=True)
        code_t = _code_dtype(uniq.size, need_dense)
        enc.t_codes = inv[:t_cp.size].astype(code_t)
        enc.o_codes = inv[t_cp.size:].astype(code_t)
        enc.symbols = [chr(c) for c in uniq.tolist()]
        return enc
    # general elements (e.g. the 2-character strings of the reference's demo, :185-186)
    table = {}
    symbols = []
    loose = []          # unhashable elements: compared with == against known representatives

    def code_of(e):
        try:
            c = table.get(e)
            if c is None:
                c = table[e] = len(symbols)
                symbols.append(e)
            return c
        except TypeError:
            for c, rep in loose:
                if rep == e:
                    return c
            c = len(symbols)
            symbols.append(e)
            loose.append((c, e))
            return c
    t_codes = [code_of(e) for e in transcript]
    o_codes = [code_of(e) for e in ocr]
    code_t = _code_dtype(len(symbols), need_dense)
    enc.t_codes = np.asarray(t_codes, dtype=code_t)
    enc.o_codes = np.asarray(o_codes, dtype=code_t)
    enc.symbols = symbols
    # a == b must mean "same code"; objects with a non-reflexive == (NaN) break that
    for s in symbols:
        if type(s) is not str:
            try:
                if not (s == s):
                    enc.reflexive = False
            except Exception:
                enc.reflexive = False
    return enc


def _tabulate(enc, fn, match, mismatch):
    """K x K int32 substitution table for the symbols of one pair.  Only (transcript symbol,
    OCR symbol) combinations are evaluated -- the reference never calls the scorer on any
    other combination (textSeqCompare.py:67)."""
    k = len(enc.symbols)
    tab = np.zeros((max(k, 1), max(k, 1)), dtype=np.int32)
    t_present = np.unique(enc.t_codes).tolist()
    o_present = np.unique(enc.o_codes).tolist()
    for a in t_present:
        sa = enc.symbols[a]
        for b in o_present:
            sb = enc.symbols[b]
            if fn is not None:
                tab[a, b] = _as_int(fn(sa, sb), 'scoring function value')
            else:
                tab[a, b] = match if sa == sb else mismatch
    return tab


def _decode(transcript, ocr, ops, enc):
    """ops (uint8, left to right) -> (tra_align, ocr_align) lists (textSeqCompare.py:116-117,
    :129-130, :139-140 after the reversal of :167-168)."""
    if enc is not None and enc.t_cp is not None:
        L = ops.size
        tra = np.full(L, ord(GAP), dtype=np.uint32)
        oc = np.full(L, ord(GAP), dtype=np.uint32)
        tra[ops != 2] = enc.t_cp
        oc[ops != 1] = enc.o_cp
        try:        # bytes -> str -> list of 1-character strings: three times faster than view('<U1').tolist()
            return (list(tra.tobytes().decode('utf-32-le', 'surrogatepass')),
                    list(oc.tobytes().decode('utf-32-le', 'surrogatepass')))
        except (ValueError, UnicodeError):
            pass
    it_t = iter(transcript)
    it_o = iter(ocr)
    ops_l = ops.tolist()
    tra = [GAP if op == 2 else next(it_t) for op in ops_l]
    oc = [GAP if op == 1 else next(it_o) for op in ops_l]
    return tra, oc


def _align_record(transcript, ocr, ops):
    """'O' equal diagonal, '~' unequal diagonal, ' ' gap (textSeqCompare.py:107, :121, :133, :143)."""
    rec = []
    x = y = 0
    for op in ops.tolist():
        if op == 0:
            rec.append('O' if transcript[x] == ocr[y] else '~')
            x += 1
            y += 1
        elif op == 1:
            rec.append(' ')
            x += 1
        else:
            rec.append(' ')
            y += 1
    return rec


def _check_list(seq, name):
    """The reference pads with ``seq + [' ']`` (:21-22), so anything that is not a list raises
    there; raise the same error here instead of silently accepting it."""
    if not isinstance(seq, list):
        seq + [' ']          # raises TypeError for str / tuple exactly as the reference does
        raise TypeError('{} must be a list'.format(name))


# ---- public API ---------------------------------------------------------------------------------

def perform_alignment(transcript, ocr, scoring_system=None, verbose=False, device=0, return_scores=False):
    '''
    @scoring_system must be array-like, of one of the following forms:
    [match_func(a,b), gap_open_x, gap_open_y, gap_extend_x, gap_extend_y]
    [match, mismatch, gap_open_x, gap_open_y, gap_extend_x, gap_extend_y]
    [match, mismatch, gap_open, gap_extend]

    Same contract as textSeqCompare.perform_alignment (textSeqCompare.py:13-177).
    `device` and `return_scores` are additive keyword arguments.
    '''
    res = perform_alignment_batch([(transcript, ocr)], scoring_system=scoring_system, devices=[device],
                                  return_scores=return_scores, _keep_ops=verbose)
    if return_scores:
        (tra, oc, sc), ops = (res[0][0], res[0][1], res[0][2]), (res[0][3] if verbose else None)
    else:
        (tra, oc), ops = (res[0][0], res[0][1]), (res[0][2] if verbose else None)
    if verbose:                                                        # :172-175
        rec = _align_record(transcript, ocr, ops)
        for k in range(len(tra)):
            line = '{} {} {}'
            print(line.format(tra[k], oc[k], rec[k]))
    if return_scores:
        return tra, oc, sc
    return (tra, oc)


def perform_alignment_batch(pairs, scoring_system=None, devices=None, return_scores=False, _keep_ops=False):
    """Align many (transcript, ocr) list pairs in one launch per device.

    Returns a list of ``(tra_align, ocr_align)`` (plus ``(M, X, Y)[n][m]`` when
    ``return_scores``), in input order.  Each pair gets exactly the result
    ``perform_alignment`` would give it."""
    fn, match, mismatch, gox, goy, gex, gey = parse_scoring_system(scoring_system)
    boundary = _as_int(gap_extend, 'gap_extend')          # module attribute, read at call time (:54-59)
    pairs = list(pairs)
    for t, o in pairs:
        _check_list(t, 'transcript')
        _check_list(o, 'ocr')
    encs = [_encode_pair(t, o, need_dense=fn is not None) for t, o in pairs]
    need_table = [fn is not None or not e.reflexive for e in encs]
    results = [None] * len(pairs)
    plain = [k for k in range(len(pairs)) if not need_table[k]]
    # pairs with 16-bit codes run on the page kernel only: keep them out of the others' launch
    for group in ([k for k in plain if encs[k].t_codes.dtype == np.uint8],
                  [k for k in plain if encs[k].t_codes.dtype != np.uint8]):
        if group:
            out = _run_group([encs[k] for k in group], (match, mismatch, gox, goy, gex, gey, boundary), None, devices)
            for k, r in zip(group, out):
                results[k] = r
    for k in range(len(pairs)):
        if need_table[k]:
            # a substitution table is specific to the pair's symbol set: one launch per pair
            if encs[k].symbols is None:
                encs[k] = _encode_pair(pairs[k][0], pairs[k][1], need_dense=True)
            tab = _tabulate(encs[k], fn, match, mismatch)
            results[k] = _run_group([encs[k]], (0, 0, gox, goy, gex, gey, boundary), tab, devices)[0]
    final = []
    for k, (ops, score) in enumerate(results):
        tra, oc = _decode(pairs[k][0], pairs[k][1], ops, encs[k])
        item = (tra, oc)
        if return_scores:
            item = item + (score,)
        if _keep_ops:
            item = item + (ops,)
        final.append(item)
    return final


def perform_alignment_sweep(pairs, scoring_systems, device=0, return_scores=False):
    """The reference's parameter sweep (evaluate_text_alignment.py:134-198 re-aligns the same
    pages under 729 scoring vectors): the pairs are encoded and uploaded once, then every
    numeric scoring system costs one launch.  Returns one result list (as
    ``perform_alignment_batch``) per scoring system."""
    pairs = list(pairs)
    for t, o in pairs:
        _check_list(t, 'transcript')
        _check_list(o, 'ocr')
    parsed = [parse_scoring_system(s) for s in scoring_systems]
    if any(p[0] is not None for p in parsed):
        return [perform_alignment_batch(pairs, s, devices=[device], return_scores=return_scores)
                for s in scoring_systems]
    boundary = _as_int(gap_extend, 'gap_extend')
    encs = [_encode_pair(t, o, need_dense=False) for t, o in pairs]
    if not all(e.reflexive for e in encs):
        return [perform_alignment_batch(pairs, s, devices=[device], return_scores=return_scores)
                for s in scoring_systems]
    symbols, t_off, n, o_off, m = _pack_encoded(encs)
    ctx = get_context(device)
    out = []
    for k, (_, match, mismatch, gox, goy, gex, gey) in enumerate(parsed):
        scoring = ctx.make_scoring(match, mismatch, gox, goy, gex, gey, boundary)
        if k == 0:
            ctx.prepare(symbols, t_off, n, o_off, m, scoring)
        else:
            ctx.rescore(scoring)
        ctx.run()
        ops, ops_off, ops_len, scores = ctx.fetch()
        res = []
        for i in range(len(pairs)):
            tra, oc = _decode(pairs[i][0], pairs[i][1], ops[ops_off[i]:ops_off[i] + ops_len[i]], encs[i])
            item = (tra, oc)
            if return_scores:
                item += (tuple(None if v == _native.NEG_INF else int(v) for v in scores[i].tolist()),)
            res.append(item)
        out.append(res)
    return out


def _pack_encoded(encs):
    n = np.asarray([e.t_codes.size for e in encs], dtype=np.int32)
    m = np.asarray([e.o_codes.size for e in encs], dtype=np.int32)
    parts = []
    for e in encs:
        parts.append(e.t_codes)
        parts.append(e.o_codes)
    symbols = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    lens = n.astype(np.int64) + m.astype(np.int64)
    t_off = np.zeros(len(encs), dtype=np.int64)
    if len(encs):
        np.cumsum(lens[:-1], out=t_off[1:])
    return symbols, t_off, n, t_off + n, m


def _run_group(encs, params, subst, devices):
    n = np.asarray([e.t_codes.size for e in encs], dtype=np.int32)
    m = np.asarray([e.o_codes.size for e in encs], dtype=np.int32)
    parts = []
    for e in encs:
        parts.append(e.t_codes)
        parts.append(e.o_codes)
    symbols = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    lens = n.astype(np.int64) + m.astype(np.int64)
    t_off = np.zeros(len(encs), dtype=np.int64)
    if len(encs):
        np.cumsum(lens[:-1], out=t_off[1:])
    o_off = t_off + n
    ops, ops_off, ops_len, scores = align_packed(symbols, t_off, n, o_off, m, params, subst=subst, devices=devices)
    out = []
    for k in range(len(encs)):
        sc = tuple(None if v == _native.NEG_INF else int(v) for v in scores[k].tolist())
        out.append((ops[ops_off[k]:ops_off[k] + ops_len[k]], sc))
    return out


def split_by_cells(n, m, parts):
    """Contiguous ranges of pairs with balanced sum(n*m): pairs are
    independent, so multi-GPU is a partition with no data-path collective.  Returns
    `parts`+1 boundaries."""
    P = int(len(n))
    parts = max(1, int(parts))
    cells = n.astype(np.int64) * m.astype(np.int64) + 1          # +1 so that empty pairs spread too
    csum = np.concatenate([[0], np.cumsum(cells)])
    targets = csum[-1] * np.arange(1, parts, dtype=np.float64) / parts
    cuts = np.searchsorted(csum, targets, side='left')
    bounds = np.concatenate([[0], cuts, [P]]).astype(np.int64)
    return np.maximum.accumulate(np.clip(bounds, 0, P))


def align_packed(symbols, t_off, n, o_off, m, params, subst=None, devices=None, want_scores=True):
    """Packed-buffer entry: uint8 (or uint16) codes + offsets in, op strings out (include/tanw.h layout).

    params = (match, mismatch, gap_open_x, gap_open_y, gap_extend_x, gap_extend_y, boundary_gap).
    devices: list of CUDA device indices; the batch is cut into contiguous, cell-balanced
    shards, one host thread + context + stream per device, results gathered by pair index."""
    if devices is None:
        devices = [0]
    devices = list(devices)
    wide = isinstance(symbols, np.ndarray) and symbols.dtype == np.uint16
    symbols = np.ascontiguousarray(symbols, dtype=np.uint16 if wide else np.uint8)
    t_off = np.ascontiguousarray(t_off, dtype=np.int64)
    o_off = np.ascontiguousarray(o_off, dtype=np.int64)
    n = np.ascontiguousarray(n, dtype=np.int32)
    m = np.ascontiguousarray(m, dtype=np.int32)
    if len(devices) == 1 or n.size < 2:
        ctx = get_context(devices[0])
        return ctx.align_batch(symbols, t_off, n, o_off, m, ctx.make_scoring(*params, subst=subst),
                               want_scores=want_scores)
    bounds = split_by_cells(n, m, len(devices))
    outs = [None] * len(devices)
    errs = [None] * len(devices)

    def work(d):
        lo, hi = int(bounds[d]), int(bounds[d + 1])
        try:
            ctx = get_context(devices[d])
            sub_sym, sub_t, sub_o = _rebase(symbols, t_off[lo:hi], n[lo:hi], o_off[lo:hi], m[lo:hi])
            outs[d] = ctx.align_batch(sub_sym, sub_t, n[lo:hi], sub_o, m[lo:hi],
                                      ctx.make_scoring(*params, subst=subst), want_scores=want_scores)
        except BaseException as e:       # re-raised on the caller's thread
            errs[d] = e
    threads = [threading.Thread(target=work, args=(d,)) for d in range(len(devices))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for e in errs:
        if e is not None:
            raise e
    return gather_shards(outs, n, m, want_scores, bounds)


def _rebase(symbols, t_off, n, o_off, m):
    """Smallest contiguous slice of `symbols` that holds a shard, and offsets into it."""
    if n.size == 0:
        return symbols[:0], t_off, o_off
    lo = int(min(t_off.min(), o_off.min()))
    hi = int(max((t_off + n).max(), (o_off + m).max()))
    return symbols[lo:hi], t_off - lo, o_off - lo


def gather_shards(outs, n, m, want_scores=True, bounds=None):
    """Host-side gather of per-device results back into one canonical layout.  Shards are
    contiguous ranges of pairs and each shard's op buffer is canonical for the shard
    (capacity sum(n+m)), so the concatenation is canonical for the batch."""
    ops_off, total = _native.Context.canonical_ops_layout(n, m)
    if bounds is None:
        bounds = np.concatenate([[0], np.cumsum([o[2].size for o in outs])]).astype(np.int64)
    cap = n.astype(np.int64) + m.astype(np.int64)
    pieces = []
    for d, o in enumerate(outs):
        shard_total = int(cap[int(bounds[d]):int(bounds[d + 1])].sum())
        pieces.append(o[0][:shard_total])
    ops = np.concatenate(pieces) if pieces else np.zeros(0, np.uint8)
    if ops.size == 0:
        ops = np.zeros(1, np.uint8)
    ops_len = np.concatenate([o[2] for o in outs]) if outs else np.zeros(0, np.int32)
    scores = np.concatenate([o[3] for o in outs]) if (outs and want_scores) else None
    return ops, ops_off, ops_len, scores


if __name__ == '__main__':
    # the reference's demo (textSeqCompare.py:180-190)
    seq1 = 'Lorem ipsum dolor sit amet, consectetur adipiscing elit '
    seq2 = 'LoLorem fipsudolor ..... sit eamet, c.nnr adizisdcing eelitellit'
    seq1 = [seq1[2 * x] + seq1[2 * x + 1] for x in range(len(seq1) // 2)]
    seq2 = [seq2[2 * x] + seq2[2 * x + 1] for x in range(len(seq2) // 2)]
    a, b = perform_alignment(seq1, seq2, scoring_system=[10, -5, -7, -7])
    print('|'.join(a))
    print('|'.join(b))
