"""TEST DOUBLE for the ctypes layer -- CPU suite only.

``fuzzysearch_b200`` has no CPU search path (``_native.Haystack`` needs libfuzzb200.so AND a CUDA device), so
on a box without a GPU nothing above the C-ABI can run.  The ``fake_device`` fixture swaps ``_native.Haystack``
/ ``_native.PinnedBuffer`` for the stand-ins below, which answer each C-ABI call from the CPU oracle
(``oracle/``, itself pinned to the reference), so that the HOST logic of the package -- dispatch, limits,
type handling, the reduction of wide symbols, slicing of ``matched``, the file chunk loops, locking -- is
exercised by ``pytest -m "not gpu"``.  Nothing here ships or is importable from the product; the real kernels
are tested by the ``-m gpu`` files through the real library.
"""
import numpy as np
import pytest

import oracle
from fuzzysearch_b200 import _native as F


class FakeResult(object):
    def __init__(self, raw, final):
        self.raw = [tuple(int(x) for x in r) for r in raw]
        self.final = [tuple(int(x) for x in r) for r in final]

    def close(self):
        pass

    def _list(self, which):
        return self.raw if which == F.RAW else self.final

    def count(self, which=F.FINAL):
        return len(self._list(which))

    def arrays(self, which=F.FINAL, anchors=False):
        a = np.array(self._list(which), dtype=np.int64).reshape(-1, 3)
        return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].astype(np.int32)

    def triples(self, which=F.FINAL):
        return list(self._list(which))

    def stats(self):
        return {"gpu_ms": 0.0, "filter_ms": 0.0, "bytes_scanned": 0, "n_candidates": 0, "n_launches": 0,
                "route": "fake"}


class FakeHaystack(object):
    """What fzb_haystack_* / fzb_search_* do, as far as the Python layer can tell."""

    def __init__(self, data, capacity):
        self.data, self.capacity = bytes(data), capacity

    @classmethod
    def from_host(cls, data, device=0, **kw):
        a = F.as_u8(data)
        return cls(a.tobytes(), a.size)

    @classmethod
    def alloc(cls, n, device=0, **kw):
        return cls(bytes(n), n)

    def close(self):
        pass

    def __len__(self):
        return len(self.data)

    def upload(self, data):
        a = F.as_u8(data)
        assert a.size <= self.capacity, "upload larger than the handle's capacity"
        self.data = a.tobytes()

    def upload_symbols(self, units, alphabet):  # k_reduce_symbols, in numpy
        units = np.ascontiguousarray(units)
        assert units.dtype in (np.uint16, np.uint32) and units.size <= self.capacity
        alpha = np.ascontiguousarray(alphabet, dtype=np.uint32)
        assert alpha.size <= F.FZB_MAX_PATTERN and np.all(alpha[1:] > alpha[:-1]), "alphabet must ascend strictly"
        if alpha.size == 0:
            self.data = bytes(units.size)
            return
        pos = np.minimum(np.searchsorted(alpha, units.astype(np.uint32)), alpha.size - 1)
        self.data = np.where(alpha[pos] == units, pos + 1, 0).astype(np.uint8).tobytes()

    def read(self, off, n):
        return self.data[off:off + n]

    def write(self, off, data):
        b = F.as_u8(data).tobytes()
        self.data = self.data[:off] + b + self.data[off + len(b):]

    @staticmethod
    def _pat(p):
        p = F.as_u8(p).tobytes()
        if len(p) > F.FZB_MAX_PATTERN:
            raise F.UnsupportedError("pattern longer than %d bytes" % F.FZB_MAX_PATTERN)
        return p

    def search_levenshtein(self, p, k, flags=0):
        if flags & F.F_FORCE_NGRAMS:
            if len(self._pat(p)) // (k + 1) == 0:
                raise ValueError("the subsequence length must be greater than max_l_dist")
            raw = oracle.levenshtein_ngrams_raw(self._pat(p), self.data, k)
        elif flags & F.F_FORCE_LP:
            raw = oracle.levenshtein_lp_raw(self._pat(p), self.data, k)
        else:
            raw = oracle.levenshtein_raw(self._pat(p), self.data, k)
        return FakeResult(raw, oracle.consolidate(raw))

    def search_hamming(self, p, k, flags=0):
        raw = oracle.substitutions(self._pat(p), self.data, k)
        return FakeResult(raw, raw)

    def search_generic(self, p, subs, ins, dels, l, flags=0):
        if flags & F.F_FORCE_NGRAMS:
            raw = oracle.generic_ngrams_raw(self._pat(p), self.data, subs, ins, dels, l)
        elif flags & F.F_FORCE_LP:
            raw = oracle.generic_lp_raw(self._pat(p), self.data, subs, ins, dels, l)
        else:
            raw = oracle.generic_raw(self._pat(p), self.data, subs, ins, dels, l)
        return FakeResult(raw, oracle.consolidate(raw))

    def search_exact(self, p, flags=0, start=None, end=None):
        p = self._pat(p)
        idx = oracle.search_exact(p, self.data, 0 if start is None else start, end)
        raw = [(i, i + len(p), 0) for i in idx]
        return FakeResult(raw, raw)

    def search_levenshtein_batch(self, pats, ks, flags=0):
        return [self.search_levenshtein(p, int(k)) for p, k in zip(pats, ks)], {}

    def has_near_match(self, p, subs, ins, dels, l):
        big = 1 << 29
        lim = [None if x >= big else x for x in (subs, ins, dels)]
        return len(oracle.find_near_matches(self._pat(p), self.data, lim[0], lim[1], lim[2], l)) > 0


class FakePinnedBuffer(object):
    def __init__(self, n):
        self.array = np.zeros(n, dtype=np.uint8)

    def close(self):
        self.array = None


@pytest.fixture()
def fake_device(monkeypatch):
    from fuzzysearch_b200 import search
    monkeypatch.setattr(F, "Haystack", FakeHaystack)
    monkeypatch.setattr(F, "PinnedBuffer", FakePinnedBuffer)
    monkeypatch.setattr(F, "device_count", lambda: 1)
    saved = dict(search._WORKSPACE)
    search._WORKSPACE.clear()
    yield
    search._WORKSPACE.clear()
    search._WORKSPACE.update(saved)
