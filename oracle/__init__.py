"""CPU parity oracle for the fuzzysearch hot path -- TEST INFRASTRUCTURE ONLY.

ctypes binding over ``oracle/libfzoracle.so`` (built from ``oracle/fzoracle.c`` by
``oracle/Makefile``), a statement-by-statement C restatement of the reference's pure-Python
algorithms (configuration "P").  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; nothing under ``fuzzysearch_b200/`` does.

The oracle is pinned (tests/test_oracle_golden.py) against fixtures produced by importing the
real reference in the build container (tests/golden/gen_golden.py).

Reference citations are relative to /root/reference/src/fuzzysearch/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfzoracle.so")


def build(force=False):
    """Compile oracle/libfzoracle.so with gcc (no-op if up to date)."""
    src = os.path.join(_HERE, "fzoracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "libfzoracle.so"])
    return _LIB_PATH


_lib = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i64pp = ctypes.POINTER(_i64p)
_vpp = ctypes.POINTER(ctypes.c_void_p)


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.fzo_free.argtypes = [ctypes.c_void_p]
        _lib.fzo_free.restype = None
        for name in ("fzo_search_exact", "fzo_levenshtein_ngrams", "fzo_levenshtein_lp",
                     "fzo_generic_lp", "fzo_generic_ngrams", "fzo_subs_lp", "fzo_subs_ngrams",
                     "fzo_consolidate"):
            getattr(_lib, name).restype = ctypes.c_int64
        _lib.fzo_expand.restype = ctypes.c_int
    return _lib


def _buf(b):
    """bytes-like -> (keepalive, void* address, length). Zero-copy for numpy/bytes."""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
        return a, ctypes.c_void_p(a.ctypes.data), a.size
    if isinstance(b, bytearray):
        a = np.frombuffer(b, dtype=np.uint8)
        return a, ctypes.c_void_p(a.ctypes.data if a.size else 0), a.size
    b = bytes(b)
    return b, ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b)


def _take_matches(ptr, count):
    """Copy `count` fzo_match records (3 x int64) out of a malloc'ed block and free it."""
    lib = _load()
    if count <= 0 or not ptr:
        if ptr:
            lib.fzo_free(ptr)
        return np.zeros((0, 3), dtype=np.int64)
    arr = np.ctypeslib.as_array(ctypes.cast(ptr, _i64p), shape=(count, 3)).copy()
    lib.fzo_free(ptr)
    return arr


def _take_i64(ptr, count):
    lib = _load()
    if count <= 0 or not ptr:
        if ptr:
            lib.fzo_free(ptr)
        return np.zeros((0,), dtype=np.int64)
    arr = np.ctypeslib.as_array(ctypes.cast(ptr, _i64p), shape=(count,)).copy()
    lib.fzo_free(ptr)
    return arr


def _as_tuples(arr):
    return [tuple(int(x) for x in row) for row in arr]


def synth(global_offset, n, alphabet, seed):
    """Bytes [global_offset, global_offset+n) of the counter-based synthetic benchmark corpus
    (fzo_synth; same definition as the device generator) -> numpy uint8."""
    lib = _load()
    ka, pa, la = _buf(alphabet)
    out = np.empty(n, dtype=np.uint8)
    lib.fzo_synth.restype = None
    lib.fzo_synth(ctypes.c_void_p(out.ctypes.data), ctypes.c_uint64(global_offset), ctypes.c_uint64(n), pa,
                  ctypes.c_uint32(la), ctypes.c_uint64(seed))
    return out


# ---------------------------------------------------------------------------------------------
# expansion DP (levenshtein_ngram.py:8-143)
# ---------------------------------------------------------------------------------------------
def expand(sub, seq, max_l_dist, which="auto"):
    """(dist, length) or (None, None).  which in {"auto","short","long"}."""
    lib = _load()
    code = {"auto": 0, "short": 1, "long": 2}[which]
    ks, ps, ls = _buf(sub)
    kq, pq, lq = _buf(seq)
    out = (ctypes.c_int * 2)()
    r = lib.fzo_expand(code, ps, ctypes.c_int(ls), pq, ctypes.c_int(lq),
                       ctypes.c_int(max_l_dist), out)
    if r < 0:
        raise ValueError("pattern too long for the oracle")
    return (out[0], out[1]) if r == 1 else (None, None)


# ---------------------------------------------------------------------------------------------
# raw streams
# ---------------------------------------------------------------------------------------------
def search_exact(sub, seq, start=0, end=None):
    """search_exact.py:22-56 -> list[int]."""
    lib = _load()
    ks, ps, ls = _buf(sub)
    kq, pq, lq = _buf(seq)
    if end is None:
        end = lq
    out = ctypes.c_void_p()
    n = lib.fzo_search_exact(ps, ctypes.c_int64(ls), pq, ctypes.c_int64(lq),
                             ctypes.c_int64(start), ctypes.c_int64(end), ctypes.byref(out))
    if n == -1:
        raise ValueError("subsequence must not be empty")
    return [int(x) for x in _take_i64(out.value, n)]


def levenshtein_ngrams_raw(pattern, haystack, k, with_anchor=False):
    """levenshtein_ngram.py:159-198 raw stream -> int64 array [N,3] (start,end,dist),
    in generation order (n-gram major, index ascending).  with_anchor -> also (ngram, idx)."""
    lib = _load()
    kp, pp, m = _buf(pattern)
    kh, ph, n = _buf(haystack)
    out = ctypes.c_void_p()
    ong = ctypes.c_void_p()
    oix = ctypes.c_void_p()
    cnt = lib.fzo_levenshtein_ngrams(pp, ctypes.c_int(m), ph, ctypes.c_int64(n), ctypes.c_int(k),
                                     ctypes.byref(out), ctypes.byref(ong), ctypes.byref(oix))
    if cnt == -2:
        raise ValueError("the subsequence length must be greater than max_l_dist")
    if cnt < 0:
        raise ValueError("oracle error %d" % cnt)
    raw = _take_matches(out.value, cnt)
    ng = _take_i64(ong.value, cnt)
    ix = _take_i64(oix.value, cnt)
    if with_anchor:
        return raw, ng, ix
    return raw


def levenshtein_lp_raw(pattern, haystack, k):
    """levenshtein.py:52-148 raw stream."""
    lib = _load()
    kp, pp, m = _buf(pattern)
    kh, ph, n = _buf(haystack)
    out = ctypes.c_void_p()
    cnt = lib.fzo_levenshtein_lp(pp, ctypes.c_int(m), ph, ctypes.c_int64(n), ctypes.c_int(k),
                                 ctypes.byref(out))
    if cnt < 0:
        raise ValueError("Given subsequence is empty!")
    return _take_matches(out.value, cnt)


def generic_lp_raw(pattern, haystack, max_subs, max_ins, max_dels, max_l):
    """generic_search.py:57-177 raw stream."""
    lib = _load()
    kp, pp, m = _buf(pattern)
    kh, ph, n = _buf(haystack)
    out = ctypes.c_void_p()
    cnt = lib.fzo_generic_lp(pp, ctypes.c_int(m), ph, ctypes.c_int64(n), ctypes.c_int(max_subs),
                             ctypes.c_int(max_ins), ctypes.c_int(max_dels), ctypes.c_int(max_l),
                             ctypes.byref(out))
    if cnt < 0:
        raise ValueError("Given subsequence is empty!")
    return _take_matches(out.value, cnt)


def generic_ngrams_raw(pattern, haystack, max_subs, max_ins, max_dels, max_l):
    """generic_search.py:198-237 raw stream."""
    lib = _load()
    kp, pp, m = _buf(pattern)
    kh, ph, n = _buf(haystack)
    out = ctypes.c_void_p()
    cnt = lib.fzo_generic_ngrams(pp, ctypes.c_int(m), ph, ctypes.c_int64(n),
                                 ctypes.c_int(max_subs), ctypes.c_int(max_ins),
                                 ctypes.c_int(max_dels), ctypes.c_int(max_l), ctypes.byref(out))
    if cnt == -2:
        raise ValueError("the subsequence length must be greater than max_l_dist")
    if cnt < 0:
        raise ValueError("Given subsequence is empty!")
    return _take_matches(out.value, cnt)


def subs_lp(pattern, haystack, k):
    """substitutions_only.py:82-136."""
    lib = _load()
    kp, pp, m = _buf(pattern)
    kh, ph, n = _buf(haystack)
    out = ctypes.c_void_p()
    cnt = lib.fzo_subs_lp(pp, ctypes.c_int(m), ph, ctypes.c_int64(n), ctypes.c_int(k),
                          ctypes.byref(out))
    if cnt < 0:
        raise ValueError("Given subsequence is empty!")
    return _take_matches(out.value, cnt)


def subs_ngrams(pattern, haystack, k):
    """substitutions_only.py:148-215 (de-duplicated, sorted by start)."""
    lib = _load()
    kp, pp, m = _buf(pattern)
    kh, ph, n = _buf(haystack)
    out = ctypes.c_void_p()
    cnt = lib.fzo_subs_ngrams(pp, ctypes.c_int(m), ph, ctypes.c_int64(n), ctypes.c_int(k),
                              ctypes.byref(out))
    if cnt == -2:
        raise ValueError("The subsequence's length must be greater than max_substitutions!")
    if cnt < 0:
        raise ValueError("Given subsequence is empty!")
    return _take_matches(out.value, cnt)


def consolidate(raw, with_groups=False):
    """common.py:145-189 (literal group_matches; ties -> smallest (start,end))."""
    lib = _load()
    raw = np.ascontiguousarray(np.asarray(raw, dtype=np.int64).reshape(-1, 3))
    n = raw.shape[0]
    out = ctypes.c_void_p()
    groups = np.zeros((max(n, 1),), dtype=np.int64)
    cnt = lib.fzo_consolidate(ctypes.c_void_p(raw.ctypes.data), ctypes.c_int64(n),
                              ctypes.byref(out), ctypes.c_void_p(groups.ctypes.data))
    res = _take_matches(out.value, cnt)
    if with_groups:
        return res, groups[:n]
    return res


# ---------------------------------------------------------------------------------------------
# routers + top-level dispatch, restated from levenshtein.py:9-38, substitutions_only.py:37-63,
# generic_search.py:25-54, __init__.py:35-83, common.py:61-116.
# ---------------------------------------------------------------------------------------------
def _exact_matches(pattern, haystack):
    m = len(pattern)
    idx = search_exact(pattern, haystack)
    return np.array([(i, i + m, 0) for i in idx], dtype=np.int64).reshape(-1, 3)


def levenshtein_raw(pattern, haystack, k):
    """find_near_matches_levenshtein (levenshtein.py:9-38)."""
    if len(pattern) == 0:
        raise ValueError("Given subsequence is empty!")
    if k == 0:
        return _exact_matches(pattern, haystack)
    if len(pattern) // (k + 1) >= 3:
        return levenshtein_ngrams_raw(pattern, haystack, k)
    return levenshtein_lp_raw(pattern, haystack, k)


def substitutions(pattern, haystack, k):
    """find_near_matches_substitutions (substitutions_only.py:37-63)."""
    if len(pattern) == 0:
        raise ValueError("Given subsequence is empty!")
    if k == 0:
        return _exact_matches(pattern, haystack)
    if len(pattern) // (k + 1) >= 3:
        return subs_ngrams(pattern, haystack, k)
    return subs_lp(pattern, haystack, k)


def generic_raw(pattern, haystack, max_subs, max_ins, max_dels, max_l):
    """find_near_matches_generic (generic_search.py:25-54)."""
    if len(pattern) == 0:
        raise ValueError("Given subsequence is empty!")
    if max_l == 0:
        return _exact_matches(pattern, haystack)
    if len(pattern) // (max_l + 1) >= 3:
        return generic_ngrams_raw(pattern, haystack, max_subs, max_ins, max_dels, max_l)
    return generic_lp_raw(pattern, haystack, max_subs, max_ins, max_dels, max_l)


def normalize_params(max_substitutions=None, max_insertions=None, max_deletions=None,
                     max_l_dist=None):
    """LevenshteinSearchParams (common.py:61-116) -> (subs, ins, dels, l_dist)."""
    vals = [max_substitutions, max_insertions, max_deletions, max_l_dist]
    if not all(x is None or (isinstance(x, int) and x >= 0) for x in vals):
        raise TypeError("All limits must be positive integers or None.")
    if max_l_dist is None:
        n_limits = sum(1 for x in vals[:3] if x is not None)
        if n_limits < 3:
            if n_limits == 0:
                raise ValueError("No limitations given!")
            elif max_substitutions is None:
                raise ValueError("# substitutions must be limited!")
            elif max_insertions is None:
                raise ValueError("# insertions must be limited!")
            else:
                raise ValueError("# deletions must be limited!")
    maxes_sum = sum(x if x is not None else 1 << 29 for x in vals[:3])
    if max_l_dist is None:
        return (max_substitutions, max_insertions, max_deletions, maxes_sum)
    norm = lambda p: min(p, max_l_dist) if p is not None else max_l_dist  # noqa: E731
    return (norm(max_substitutions), norm(max_insertions), norm(max_deletions),
            min(max_l_dist, maxes_sum))


def find_near_matches(pattern, haystack, max_substitutions=None, max_insertions=None,
                      max_deletions=None, max_l_dist=None, return_raw=False):
    """Top-level dispatch (__init__.py:35-83) -> list[(start,end,dist)].

    Ties inside a group of overlapping matches go to the smallest (start,end) (the reference's
    choice there is hash-seed dependent)."""
    subs, ins, dels, l = normalize_params(max_substitutions, max_insertions, max_deletions,
                                          max_l_dist)
    if l == 0:  # ExactSearch  (__init__.py:65-66)
        if len(pattern) == 0:
            raise ValueError("subsequence must not be empty")
        raw = _exact_matches(pattern, haystack)
        final = raw
    elif ins == 0 and dels == 0:  # SubstitutionsOnlySearch (:69-70)
        raw = substitutions(pattern, haystack, min(l, subs))  # substitutions_only.py:291-295
        final = raw
    elif l <= min(subs, ins, dels):  # LevenshteinSearch (:74-79)
        raw = levenshtein_raw(pattern, haystack, l)
        final = consolidate(raw)
    else:  # GenericSearch (:82-83)
        raw = generic_raw(pattern, haystack, subs, ins, dels, l)
        final = consolidate(raw)
    if return_raw:
        return _as_tuples(final), _as_tuples(raw)
    return _as_tuples(final)
