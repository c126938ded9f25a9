/* tanw_pylist.c -- CPython-side marshalling for the drop-in signature.
 *
 * textSeqCompare.perform_alignment takes and returns Python LISTS of elements
 * (/root/reference/textSeqCompare.py:13-22, :167-177; the call site passes list(str),
 * alignToOCR.py:273).  Once the alignment itself takes microseconds, turning a list of 1500
 * one-character strings into codes and an op string back into two lists is the cost of a call;
 * ''.join / list(str) alone are ~30 us per page in pure Python.  These two helpers do it with the
 * CPython C API (3-4 ns per element).  Loaded with ctypes.PyDLL (the GIL is held); no device code.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* Code points of a list of one-character strings.  Returns the number of elements written, or
 * -1 if `seq` is not a list, holds anything but 1-character str objects, or exceeds `cap`
 * (then the caller takes the general interning path).  *maxcp receives the largest code point. */
Py_ssize_t tanw_pylist_codepoints(PyObject *seq, uint32_t *out, Py_ssize_t cap, uint32_t *maxcp)
{
    if (!PyList_CheckExact(seq)) return -1;
    const Py_ssize_t k = PyList_GET_SIZE(seq);
    if (k > cap) return -1;
    uint32_t top = *maxcp;
    for (Py_ssize_t i = 0; i < k; ++i) {
        PyObject *e = PyList_GET_ITEM(seq, i);
        if (!PyUnicode_CheckExact(e) || PyUnicode_GET_LENGTH(e) != 1) return -1;
        const uint32_t c = (uint32_t)PyUnicode_READ_CHAR(e, 0);
        out[i] = c;
        if (c > top) top = c;
    }
    *maxcp = top;
    return k;
}

/* One aligned sequence (textSeqCompare.py:116-117, :129-130, :139-140 after the reversal of
 * :167-168): for every op, `gap` where op == gap_op, else the next element of `src` -- the
 * caller's own objects, as in the reference.  Returns a new list, or NULL (with an exception set)
 * if the ops consume more or fewer elements than `src` holds. */
PyObject *tanw_pylist_expand(PyObject *src, const uint8_t *ops, Py_ssize_t L, int gap_op, PyObject *gap)
{
    if (!PyList_CheckExact(src)) {
        PyErr_SetString(PyExc_TypeError, "tanw_pylist_expand: src must be a list");
        return NULL;
    }
    const Py_ssize_t k = PyList_GET_SIZE(src);
    PyObject *out = PyList_New(L);
    if (!out) return NULL;
    Py_ssize_t x = 0;
    for (Py_ssize_t i = 0; i < L; ++i) {
        PyObject *e;
        if (ops[i] == (uint8_t)gap_op) {
            e = gap;
        } else {
            if (x >= k) {
                Py_DECREF(out);
                PyErr_SetString(PyExc_ValueError, "tanw_pylist_expand: op string consumes more elements than the sequence holds");
                return NULL;
            }
            e = PyList_GET_ITEM(src, x++);
        }
        Py_INCREF(e);
        PyList_SET_ITEM(out, i, e);
    }
    if (x != k) {
        Py_DECREF(out);
        PyErr_SetString(PyExc_ValueError, "tanw_pylist_expand: op string consumes fewer elements than the sequence holds");
        return NULL;
    }
    return out;
}

/* Substrings text[bounds[2i] : bounds[2i+1]] as a new list (the syllables of a transcript,
 * latinSyllabification.syllabify_text -> alignToOCR.py:277, from the native syllabifier's ranges).
 * `keep`, when not NULL, selects the ranges to take.  NULL with an exception set on bad ranges. */
PyObject *tanw_pylist_slices(PyObject *text, const int32_t *bounds, Py_ssize_t count, const uint8_t *keep)
{
    if (!PyUnicode_CheckExact(text)) {
        PyErr_SetString(PyExc_TypeError, "tanw_pylist_slices: text must be a str");
        return NULL;
    }
    const Py_ssize_t len = PyUnicode_GET_LENGTH(text);
    Py_ssize_t kept = count;
    if (keep) {
        kept = 0;
        for (Py_ssize_t i = 0; i < count; ++i) kept += keep[i] != 0;
    }
    PyObject *out = PyList_New(kept);
    if (!out) return NULL;
    Py_ssize_t at = 0;
    for (Py_ssize_t i = 0; i < count; ++i) {
        if (keep && !keep[i]) continue;
        const Py_ssize_t a = bounds[2 * i], b = bounds[2 * i + 1];
        if (a < 0 || b < a || b > len) {
            Py_DECREF(out);
            PyErr_SetString(PyExc_ValueError, "tanw_pylist_slices: range outside the text");
            return NULL;
        }
        PyObject *piece = PyUnicode_Substring(text, a, b);
        if (!piece) {
            Py_DECREF(out);
            return NULL;
        }
        PyList_SET_ITEM(out, at++, piece);
    }
    return out;
}
