"""Pure-Python restatement of the reference aligner -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Restates /root/reference/textSeqCompare.py:13-177 with the same data structures the
reference uses (six float64 numpy matrices addressed one scalar at a time, builtin ``max``,
``list.index``), so that it costs what the reference costs per cell (~10 us) and produces the
same values, but is laid out as separate phases so it can travel to the GPU box where
/root/reference does not exist.  It is what ``bench.py --impl reference`` and the
``cpu_baseline`` leg time ("kind": "port").  Pinned against the live reference and the golden
vectors by tests/test_oracle.py.
"""
import numpy as np

NEG = -1e100                      # textSeqCompare.py:55, :60
DEFAULT_SYS = [8, -4, -7, -7, -3, 0]   # textSeqCompare.py:10
BOUNDARY_GAP = -1                 # module-level gap_extend, textSeqCompare.py:9


def _parse(scoring_system):
    """textSeqCompare.py:24-42 -> (score_fn, gox, goy, gex, gey)."""
    ss = DEFAULT_SYS if scoring_system is None else scoring_system
    if len(ss) == 5 and callable(ss[0]):
        return ss[0], ss[1], ss[2], ss[3], ss[4]
    if len(ss) == 6:
        return (lambda a, b: ss[0] if a == b else ss[1]), ss[2], ss[3], ss[4], ss[5]
    if len(ss) == 4:
        return (lambda a, b: ss[0] if a == b else ss[1]), ss[2], ss[2], ss[3], ss[3]
    raise ValueError('scoring_system {} invalid'.format(ss))


def fill(T, O, scoring_system=None, boundary_gap=BOUNDARY_GAP):
    """Boundary + recurrence (textSeqCompare.py:45-88).  Returns the six matrices."""
    score, gox, goy, gex, gey = _parse(scoring_system)
    rows, cols = len(T) + 1, len(O) + 1
    M = np.zeros((rows, cols)); Y = np.zeros((rows, cols)); X = np.zeros((rows, cols))
    PM = np.zeros((rows, cols)); PY = np.zeros((rows, cols)); PX = np.zeros((rows, cols))
    for i in range(rows):                         # :53-56
        M[i][0] = boundary_gap * i
        X[i][0] = NEG
        Y[i][0] = boundary_gap * i
    for j in range(cols):                         # :57-60 (runs second: X[0][0] ends up 0)
        M[0][j] = boundary_gap * j
        X[0][j] = boundary_gap * j
        Y[0][j] = NEG
    for i in range(1, rows):                      # :62-88
        a = T[i - 1]
        for j in range(1, cols):
            diag = [M[i - 1][j - 1], X[i - 1][j - 1], Y[i - 1][j - 1]]
            top = max(diag)
            M[i][j] = top + score(a, O[j - 1])
            PM[i][j] = diag.index(top)
            left = [M[i][j - 1] + goy + gey, X[i][j - 1] + goy + gey, Y[i][j - 1] + gey]
            top = max(left)
            Y[i][j] = top
            PY[i][j] = left.index(top)
            up = [M[i - 1][j] + gox + gex, X[i - 1][j] + gex, Y[i - 1][j] + gox + gex]
            top = max(up)
            X[i][j] = top
            PX[i][j] = up.index(top)
    return M, X, Y, PM, PX, PY


def walk(T, O, PM, PX, PY):
    """Traceback, tail flush, reversal (textSeqCompare.py:96-170) as op codes
    0 = (T,O), 1 = (T,'_'), 2 = ('_',O), left to right."""
    x, y = len(T), len(O)
    state = int(PM[x][y])                          # :102
    ops = []
    while x > 0 and y > 0:                         # :110-145
        if state == 0:
            ops.append(0); state = int(PM[x][y]); x -= 1; y -= 1
        elif state == 1:
            ops.append(1); state = int(PX[x][y]); x -= 1
        else:
            ops.append(2); state = int(PY[x][y]); y -= 1
    ops.extend([2] * y)                            # :154-158 (OCR remainder first)
    ops.extend([1] * x)                            # :160-164
    ops.reverse()                                  # :167-170
    return ops


def ops_to_alignment(T, O, ops, gap='_'):
    tra, ocr = [], []
    x = y = 0
    for op in ops:
        if op == 0:
            tra.append(T[x]); ocr.append(O[y]); x += 1; y += 1
        elif op == 1:
            tra.append(T[x]); ocr.append(gap); x += 1
        else:
            tra.append(gap); ocr.append(O[y]); y += 1
    return tra, ocr


def perform_alignment(transcript, ocr, scoring_system=None, boundary_gap=BOUNDARY_GAP, full=False):
    """Same inputs/outputs as textSeqCompare.perform_alignment (:13, :177)."""
    M, X, Y, PM, PX, PY = fill(transcript, ocr, scoring_system, boundary_gap)
    ops = walk(transcript, ocr, PM, PX, PY)
    tra, oc = ops_to_alignment(transcript, ocr, ops)
    if full:
        return tra, oc, dict(M=M, X=X, Y=Y, PM=PM, PX=PX, PY=PY, ops=ops)
    return tra, oc
