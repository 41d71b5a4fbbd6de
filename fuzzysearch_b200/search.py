"""The four search classes -- the reference's plugin seam.

``fuzzysearch.choose_search_class`` (__init__.py:60-83) picks one of ExactSearch /
SubstitutionsOnlySearch / LevenshteinSearch / GenericSearch, each a ``FuzzySearchBase`` with
``search(subsequence, sequence, search_params)`` and ``consolidate_matches(matches)``
(common.py:192-209).  Here each ``search`` is ONE call into libfuzzb200.so (sm_100a kernels);
there is no CPU implementation behind them.
"""
import threading
from collections.abc import Sequence

import numpy as np

from . import _native
from .common import FuzzySearchBase, Match, consolidate_overlapping_matches

__all__ = ["DeviceSequence", "ExactSearch", "SubstitutionsOnlySearch", "LevenshteinSearch",
           "GenericSearch", "RawMatches"]


class DeviceSequence(object):
    """A haystack kept resident in HBM so that several searches reuse one upload.

    ``find_near_matches(pattern, DeviceSequence(data), ...)`` behaves like
    ``find_near_matches(pattern, data, ...)``."""

    def __init__(self, data=None, device=0, _haystack=None, _host=None):
        if _haystack is not None:
            self.haystack, self._host = _haystack, _host
        else:
            self._host, self._is_str = _coerce(data)
            self.haystack = _native.Haystack.from_host(self._host, device=device)
        self._is_str = getattr(self, "_is_str", False)

    def __len__(self):
        return len(self.haystack)

    def slice(self, start, end):
        if self._host is not None:
            b = bytes(memoryview(self._host)[start:end])
        else:
            b = self.haystack.read(start, end - start)
        return b.decode("latin-1") if self._is_str else b

    def close(self):
        self.haystack.close()


def _coerce(seq):
    """-> (uint8 view, is_str).  bytes-like stay zero-copy; latin-1 encodable str are encoded."""
    if isinstance(seq, str):
        try:
            return np.frombuffer(seq.encode("latin-1"), dtype=np.uint8), True
        except UnicodeEncodeError:
            raise TypeError("str sequences must be latin-1 encodable (single-byte symbols); "
                            "fuzzysearch_b200 has no CPU fallback for general Unicode")
    if isinstance(seq, (list, tuple)):
        raise TypeError("unsupported sequence type: %s (byte-like sequences only)" % type(seq))
    return _native.as_u8(seq), False


class RawMatches(Sequence):
    """The raw match stream of one search, materialised LAZILY: ``find_near_matches`` only needs the
    consolidated list (``.final``, built from the device-side consolidation), so the raw ``Match``
    objects -- one Python object and one slice per raw record -- are only built if somebody iterates,
    indexes or compares this object (the search-class ``search()`` contract of the reference,
    common.py:192-197, is "an iterable of Match")."""

    def __init__(self, result, slicer, consolidated):
        self._result = result
        self._slicer = slicer
        self._list = None
        self.stats = result.stats()
        self.final = _to_matches(result, _native.FINAL, slicer) if consolidated else None
        self._n = result.count(_native.RAW)

    def materialize(self):
        if self._list is None:
            self._list = _to_matches(self._result, _native.RAW, self._slicer)
            self._result.close()
            self._result = None
        return self._list

    def __len__(self):
        return self._n

    def __iter__(self):
        return iter(self.materialize())

    def __getitem__(self, i):
        return self.materialize()[i]

    def __eq__(self, other):
        if isinstance(other, RawMatches):
            other = other.materialize()
        return self.materialize() == other

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return repr(self.materialize())

    def __del__(self):
        if getattr(self, "_result", None) is not None:
            self._result.close()


def _prepare(subsequence, sequence):
    """-> (pattern u8, haystack handle, slicer, owns_handle)"""
    pat, pat_is_str = _coerce(subsequence)
    if isinstance(sequence, DeviceSequence):
        return pat, sequence.haystack, sequence.slice, False
    host, is_str = _coerce(sequence)
    if is_str != pat_is_str:
        raise TypeError("subsequence and sequence must both be str or both be byte-like")
    hay = _workspace(host.size)
    hay.upload(host)
    if is_str:
        text = sequence

        def slicer(s, e):
            return text[s:e]
    elif isinstance(sequence, (bytes, bytearray)):
        def slicer(s, e):
            return sequence[s:e]
    else:
        mv = memoryview(host)

        def slicer(s, e):
            return bytes(mv[s:e])
    return pat, hay, slicer, False


_WORKSPACE = {}
# One lock per process around upload + search + copy-out of the SHARED workspace: ctypes drops the GIL
# during the native calls, so without it two threads calling find_near_matches() would overwrite each
# other's haystack mid-search.  (The reference is serialised by the GIL; DeviceSequence handles are
# serialised per handle inside the library.)
_WORKSPACE_LOCK = threading.RLock()


def _workspace(nbytes, device=0):
    """A cached device buffer for plain (host) sequences: like the reference's reusable chunk buffer
    (__init__.py:141-145), a search then costs one H2D copy instead of allocations."""
    ws = _WORKSPACE.get(device)
    if ws is None or ws[1] < nbytes:
        if ws is not None:
            ws[0].close()
        cap = max(nbytes + nbytes // 8, 1 << 20)
        ws = (_native.Haystack.alloc(cap, device=device), cap)
        _WORKSPACE[device] = ws
    return ws[0]


def release_workspace():
    for ws in _WORKSPACE.values():
        ws[0].close()
    _WORKSPACE.clear()
    _native.lib().fzb_release_workspace()


def _to_matches(result, which, slicer):
    s, e, d = result.arrays(which)
    return [Match(a, b, c, matched=slicer(a, b)) for a, b, c in zip(s.tolist(), e.tolist(), d.tolist())]


def _run(subsequence, sequence, call, consolidated):
    shared = not isinstance(sequence, DeviceSequence)
    if shared:
        _WORKSPACE_LOCK.acquire()
    try:
        pat, hay, slicer, _ = _prepare(subsequence, sequence)
        res = call(hay, pat)
        try:
            return RawMatches(res, slicer, consolidated)
        except BaseException:
            res.close()
            raise
    finally:
        if shared:
            _WORKSPACE_LOCK.release()


class ExactSearch(FuzzySearchBase):
    """search_exact.py:80-89."""

    @classmethod
    def search(cls, subsequence, sequence, search_params=None):
        if len(subsequence) == 0:
            raise ValueError("subsequence must not be empty")
        return _run(subsequence, sequence, lambda h, p: h.search_exact(p), False)

    @classmethod
    def extra_items_for_chunked_search(cls, subsequence, search_params):
        return 0


class SubstitutionsOnlySearch(FuzzySearchBase):
    """substitutions_only.py:288-301."""

    @classmethod
    def search(cls, subsequence, sequence, search_params):
        if len(subsequence) == 0:
            raise ValueError("Given subsequence is empty!")
        actual_max_subs = min(x for x in [search_params.max_l_dist, search_params.max_substitutions]
                              if x is not None)
        return _run(subsequence, sequence, lambda h, p: h.search_hamming(p, actual_max_subs), False)

    @classmethod
    def extra_items_for_chunked_search(cls, subsequence, search_params):
        return 0


class LevenshteinSearch(FuzzySearchBase):
    """levenshtein.py:151-164."""

    @classmethod
    def search(cls, subsequence, sequence, search_params):
        if len(subsequence) == 0:
            raise ValueError("Given subsequence is empty!")
        k = search_params.max_l_dist
        return _run(subsequence, sequence, lambda h, p: h.search_levenshtein(p, k), True)

    @classmethod
    def consolidate_matches(cls, matches):
        return consolidate_overlapping_matches(matches)

    @classmethod
    def extra_items_for_chunked_search(cls, subsequence, search_params):
        return search_params.max_l_dist


class GenericSearch(FuzzySearchBase):
    """generic_search.py:256-273."""

    @classmethod
    def search(cls, subsequence, sequence, search_params):
        if len(subsequence) == 0:
            raise ValueError("Given subsequence is empty!")
        subs, ins, dels, l = search_params.unpacked
        return _run(subsequence, sequence, lambda h, p: h.search_generic(p, subs, ins, dels, l), True)

    @classmethod
    def consolidate_matches(cls, matches):
        return consolidate_overlapping_matches(matches)

    @classmethod
    def extra_items_for_chunked_search(cls, subsequence, search_params):
        return max(x for x in [search_params.max_l_dist, search_params.max_insertions] if x is not None)
