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


def get_context(device=0, replica=0):
    """Process-wide native context of a device (created on first use).  A context serialises its
    callers (``Context.lock``); ``replica`` > 0 names further contexts on the same device, which
    ``align_packed`` uses when its device list names one device more than once."""
    key = (int(device), int(replica))
    with _contexts_lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = _contexts[key] = _native.Context(int(device))
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
#
# The device sees uint8 (or uint16) codes with "equal code <=> elements compare equal".  A whole
# batch is interned at once, so that one launch -- and, for a callable scorer, one K x K table --
# serves all its pairs.  Two routes:
#   * every element is a 1-character str (the production call, alignToOCR.py:273): code points,
#     read with the CPython helper (csrc/tanw_pylist.c) or one join + encode per sequence;
#   * anything else (the 2-character strings of the reference's demo :185-186, tuples, ...):
#     a dict-based interner over the batch.

MAX_SYMBOLS = 65536          # distinct elements per launch (uint16 codes)
MAX_TABLE_SYMBOLS = 2048     # ... when the scorer needs a K x K table (include/tanw.h)


class _Encoded(object):
    """A batch of pairs as packed codes.  ``alphabet``: the element each code stands for, or None
    when the codes are the elements' own code points."""
    __slots__ = ('symbols', 't_off', 'n', 'o_off', 'm', 'alphabet', 'reflexive', 'chars')


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


def _batch_code_points(pairs, n, m):
    """Code points of every sequence of the batch, concatenated T0 O0 T1 O1 ..., or None when some
    element is not a 1-character str."""
    total = int(n.sum() + m.sum())
    cp = np.empty(max(total, 1), dtype=np.uint32)
    helper = _native.pylist()
    if helper is not None:
        top = np.zeros(1, dtype=np.uint32)
        base, ptop, fn = cp.ctypes.data, top.ctypes.data, helper.tanw_pylist_codepoints
        at = 0
        for t, o in pairs:
            k = fn(t, base + 4 * at, total - at, ptop)
            if k < 0:
                return None
            at += k
            k = fn(o, base + 4 * at, total - at, ptop)
            if k < 0:
                return None
            at += k
        return cp[:total]
    at = 0
    for t, o in pairs:
        for seq in (t, o):
            c = _code_points(seq)
            if c is None:
                return None
            cp[at:at + c.size] = c
            at += c.size
    return cp[:total]


def _layout(n, m):
    lens = n.astype(np.int64) + m
    t_off = np.zeros(n.size, dtype=np.int64)
    if n.size:
        np.cumsum(lens[:-1], out=t_off[1:])
    return t_off, t_off + n


def _code_dtype(distinct, tabulated):
    """uint8 codes while they suffice, uint16 beyond; None when one launch cannot hold them."""
    if distinct <= 256:
        return np.uint8
    if distinct <= (MAX_TABLE_SYMBOLS if tabulated else MAX_SYMBOLS):
        return np.uint16
    return None


def _encode_batch(pairs, tabulated):
    """-> _Encoded for the whole batch, or None when its distinct elements do not fit one launch
    (the caller then encodes pair by pair)."""
    enc = _Encoded()
    enc.reflexive = True
    enc.n = np.fromiter((len(t) for t, _ in pairs), dtype=np.int32, count=len(pairs))
    enc.m = np.fromiter((len(o) for _, o in pairs), dtype=np.int32, count=len(pairs))
    enc.t_off, enc.o_off = _layout(enc.n, enc.m)
    cp = _batch_code_points(pairs, enc.n, enc.m)
    enc.chars = cp is not None
    if cp is not None:
        if not tabulated and (cp.size == 0 or int(cp.max()) < 256):
            enc.symbols = cp.astype(np.uint8)
            enc.alphabet = None
            return enc
        uniq, inv = np.unique(cp, return_inverse=True)
        dt = _code_dtype(uniq.size, tabulated)
        if dt is None:
            return None
        enc.symbols = inv.astype(dt)
        enc.alphabet = [chr(c) for c in uniq.tolist()]
        return enc
    # general elements: equal under == (and hash) -> same code; unhashable ones are compared
    # with == against the representatives seen so far
    table, alphabet, loose = {}, [], []

    def code_of(e):
        try:
            c = table.get(e)
            if c is None:
                c = table[e] = len(alphabet)
                alphabet.append(e)
            return c
        except TypeError:
            for c, rep in loose:
                if rep == e:
                    return c
            c = len(alphabet)
            alphabet.append(e)
            loose.append((c, e))
            return c
    codes = []
    for t, o in pairs:
        codes.extend(code_of(e) for e in t)
        codes.extend(code_of(e) for e in o)
    dt = _code_dtype(len(alphabet), True)      # a non-reflexive element would need the table
    if dt is None:
        return None
    enc.symbols = np.asarray(codes, dtype=dt) if codes else np.zeros(0, dtype=dt)
    enc.alphabet = alphabet
    # a == b must mean "same code"; objects with a non-reflexive == (NaN) break that
    for e in alphabet:
        if type(e) is not str:
            try:
                if not (e == e):
                    enc.reflexive = False
            except Exception:
                enc.reflexive = False
    if enc.reflexive and _code_dtype(len(alphabet), tabulated) is None:
        return None
    return enc


def _tabulate(enc, fn, match, mismatch):
    """K x K int32 substitution table for the alphabet of a batch.  Only (transcript symbol,
    OCR symbol) combinations that meet in some pair are evaluated -- the reference never calls
    the scorer on any other combination (textSeqCompare.py:67)."""
    k = max(len(enc.alphabet), 1)
    tab = np.zeros((k, k), dtype=np.int32)
    P = enc.n.size
    if P == 0 or enc.symbols.size == 0:
        return tab
    # which symbols occur in the transcript / the OCR of which pair
    pair_of_t = np.repeat(np.arange(P), enc.n)
    pair_of_o = np.repeat(np.arange(P), enc.m)
    t_idx = np.repeat(enc.t_off, enc.n) + (np.arange(pair_of_t.size) - np.repeat(np.cumsum(enc.n) - enc.n, enc.n))
    o_idx = np.repeat(enc.o_off, enc.m) + (np.arange(pair_of_o.size) - np.repeat(np.cumsum(enc.m) - enc.m, enc.m))
    in_t = np.zeros((P, k), dtype=bool)
    in_o = np.zeros((P, k), dtype=bool)
    in_t[pair_of_t, enc.symbols[t_idx]] = True
    in_o[pair_of_o, enc.symbols[o_idx]] = True
    meet = (in_t.astype(np.float32).T @ in_o.astype(np.float32)) > 0
    for a, b in zip(*np.nonzero(meet)):
        sa, sb = enc.alphabet[a], enc.alphabet[b]
        if fn is not None:
            tab[a, b] = _as_int(fn(sa, sb), 'scoring function value')
        else:
            tab[a, b] = match if sa == sb else mismatch
    return tab


def _expand(seq, ops, gap_op):
    """One aligned sequence (textSeqCompare.py:116-117, :129-130, :139-140 after the reversal of
    :167-168): GAP where the op is `gap_op`, else the caller's own next element."""
    helper = _native.pylist()
    if helper is not None:
        ops = np.ascontiguousarray(ops)
        return helper.tanw_pylist_expand(seq, ops.ctypes.data, ops.size, gap_op, GAP)
    it = iter(seq)
    return [GAP if op == gap_op else next(it) for op in ops.tolist()]


def _decode(transcript, ocr, ops):
    """ops (uint8, left to right) -> (tra_align, ocr_align) lists."""
    return _expand(transcript, ops, 2), _expand(ocr, ops, 1)


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


def _score_tuple(row):
    return tuple(None if v == _native.NEG_INF else int(v) for v in row.tolist())


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


def _align_encoded(enc, fn, numeric, boundary, devices):
    """One launch per device for an encoded batch -> (ops, ops_off, ops_len, scores)."""
    match, mismatch, gox, goy, gex, gey = numeric
    if fn is not None or not enc.reflexive:
        tab = _tabulate(enc, fn, match, mismatch)
        params, subst = (0, 0, gox, goy, gex, gey, boundary), tab
    else:
        params, subst = (match, mismatch, gox, goy, gex, gey, boundary), None
    return align_packed(enc.symbols, enc.t_off, enc.n, enc.o_off, enc.m, params, subst=subst, devices=devices)


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
    if not pairs:
        return []
    numeric = (match, mismatch, gox, goy, gex, gey)
    enc = _encode_batch(pairs, tabulated=fn is not None)
    if enc is not None:
        groups = [(range(len(pairs)), _align_encoded(enc, fn, numeric, boundary, devices))]
    else:
        # more distinct elements than one launch can code: pair by pair (each has its own alphabet)
        groups = []
        for k, pr in enumerate(pairs):
            one = _encode_batch([pr], tabulated=fn is not None)
            if one is None:
                raise ValueError('more than {} distinct symbols in one pair; the device path cannot '
                                 'represent them'.format(MAX_TABLE_SYMBOLS if fn is not None else MAX_SYMBOLS))
            groups.append(([k], _align_encoded(one, fn, numeric, boundary, devices)))
    final = [None] * len(pairs)
    for members, (ops, ops_off, ops_len, scores) in groups:
        for j, k in enumerate(members):
            o = ops[ops_off[j]:ops_off[j] + ops_len[j]]
            item = _decode(pairs[k][0], pairs[k][1], o)
            if return_scores:
                item = item + (_score_tuple(scores[j]),)
            if _keep_ops:
                item = item + (o,)
            final[k] = item
    return final


def align_strings(pairs, scoring_system=None, devices=None):
    """(transcript str, OCR str) pairs -> (ops, ops_off, ops_len): the alignment of
    ``perform_alignment(list(t), list(o))`` as op strings, for callers that stay on arrays
    (alignToOCR.boxes_for_pages_arrays).  One encode for the whole batch, one launch per device."""
    fn, match, mismatch, gox, goy, gex, gey = parse_scoring_system(scoring_system)
    boundary = _as_int(gap_extend, 'gap_extend')
    enc = _Encoded()
    enc.reflexive, enc.chars = True, True
    enc.n = np.fromiter((len(t) for t, _ in pairs), dtype=np.int32, count=len(pairs))
    enc.m = np.fromiter((len(o) for _, o in pairs), dtype=np.int32, count=len(pairs))
    enc.t_off, enc.o_off = _layout(enc.n, enc.m)
    text = ''.join(t + o for t, o in pairs)
    try:
        if fn is not None:
            raise UnicodeEncodeError('latin-1', '', 0, 0, 'a table needs dense codes')
        enc.symbols = np.frombuffer(text.encode('latin-1'), dtype=np.uint8)
        enc.alphabet = None
    except UnicodeEncodeError:
        cp = np.frombuffer(text.encode('utf-32-le', 'surrogatepass'), dtype=np.uint32)
        uniq, inv = np.unique(cp, return_inverse=True)
        dt = _code_dtype(uniq.size, fn is not None)
        if dt is None:
            raise ValueError('more distinct characters in one batch ({}) than the device path can code'.format(uniq.size))
        enc.symbols = inv.astype(dt)
        enc.alphabet = [chr(c) for c in uniq.tolist()]
    ops, ops_off, ops_len, _ = _align_encoded(enc, fn, (match, mismatch, gox, goy, gex, gey), boundary, devices)
    return ops, ops_off, ops_len


def perform_alignment_sweep(pairs, scoring_systems, device=0, return_scores=False):
    """The reference's parameter sweep (evaluate_text_alignment.py:134-198 re-aligns the same
    pages under 729 scoring vectors) as ONE launch: the pairs are encoded and uploaded once and
    every (scoring system, pair) combination is a pair of the batch with its own parameters
    (tanw_align_batch_multi).  Returns one result list (as ``perform_alignment_batch``) per
    scoring system."""
    pairs = list(pairs)
    for t, o in pairs:
        _check_list(t, 'transcript')
        _check_list(o, 'ocr')
    scoring_systems = list(scoring_systems)
    parsed = [parse_scoring_system(s) for s in scoring_systems]
    boundary = _as_int(gap_extend, 'gap_extend')
    enc = _encode_batch(pairs, tabulated=False) if pairs and not any(p[0] is not None for p in parsed) else None
    if enc is None or not enc.reflexive or enc.symbols.dtype != np.uint8:
        # callables, or alphabets beyond 8-bit codes: one batch per scoring system
        return [perform_alignment_batch(pairs, s, devices=[device], return_scores=return_scores)
                for s in scoring_systems]
    P, S = len(pairs), len(parsed)
    systems = [(mt, mi, gox, goy, gex, gey, boundary) for _, mt, mi, gox, goy, gex, gey in parsed]
    ctx = get_context(device)
    ops, ops_off, ops_len, scores = ctx.align_batch_multi(
        enc.symbols, np.tile(enc.t_off, S), np.tile(enc.n, S), np.tile(enc.o_off, S), np.tile(enc.m, S),
        systems, np.repeat(np.arange(S, dtype=np.int32), P))
    out = []
    for k in range(S):
        res = []
        for i in range(P):
            j = k * P + i
            item = _decode(pairs[i][0], pairs[i][1], ops[ops_off[j]:ops_off[j] + ops_len[j]])
            if return_scores:
                item += (_score_tuple(scores[j]),)
            res.append(item)
        out.append(res)
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


def align_packed(symbols, t_off, n, o_off, m, params, subst=None, devices=None, want_scores=True, out=None):
    """Packed-buffer entry: uint8 (or uint16) codes + offsets in, op strings out (include/tanw.h layout).

    params = (match, mismatch, gap_open_x, gap_open_y, gap_extend_x, gap_extend_y, boundary_gap).
    devices: list of CUDA device indices; the batch is cut into contiguous, cell-balanced
    shards, one host thread + context + stream per device.  Shards are contiguous ranges of pairs,
    so every device writes its results straight into its slice of the batch's output arrays: the
    "host-side gather" of SURVEY.md 8(e) is the layout itself, no copy.
    out = (ops, ops_len, scores) preallocated (e.g. page-locked) arrays for the whole batch."""
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
                               want_scores=want_scores, out=out)
    bounds = split_by_cells(n, m, len(devices))
    contexts = [get_context(dev, replica=devices[:d].count(dev)) for d, dev in enumerate(devices)]
    # one native host thread per shard (tanw_align_batch_sharded): Python threads would take turns
    # at the interpreter lock for their share of the call's host work before any device starts
    return _native.Context.align_batch_sharded(contexts, bounds, symbols, t_off, n, o_off, m, params, subst=subst,
                                               want_scores=want_scores, out=out)


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
