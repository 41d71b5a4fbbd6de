"""find_near_matches_in_file against an oracle-driven emulation of the reference's
chunk loop (__init__.py:129-171), incl. matches straddling the chunk boundary at the deltas of
tests/test_find_near_matches_in_file.py:84-86."""
import io

import numpy as np
import pytest

import oracle
from fuzzysearch_b200 import find_near_matches, find_near_matches_in_file
from parity import tup

pytestmark = pytest.mark.gpu


def _emulate(pattern, data, k, chunk_size, binary=True):
    """The reference's _search_binary_file (binary=True, __init__.py:129-171) or _search_unicode_file
    (any other file object, :174-200) with the oracle as the per-chunk search."""
    keep = len(pattern) - 1 + k
    raw = []
    f = io.BytesIO(data)
    chunk = f.read(chunk_size)
    offset = 0
    while True:
        raw += [(s + offset, e + offset, d) for s, e, d in tup(oracle.levenshtein_raw(pattern, chunk, k))]
        n_keep = min(keep, len(chunk))
        offset += len(chunk) - n_keep
        block = f.read(max(0, chunk_size - n_keep) if binary else chunk_size)
        if not block:
            break
        chunk = chunk[len(chunk) - n_keep:] + block
    return tup(oracle.consolidate(np.array(raw, dtype=np.int64).reshape(-1, 3)))


def _t(ms):
    return [(m.start, m.end, m.dist) for m in ms]


@pytest.mark.parametrize("chunk_size", [100, 4096, 1 << 16])
def test_match_split_between_chunks(cuda_device, chunk_size, tmp_path):
    needle = b"PATTERNXYZ12"
    for k in (1, 2):
        for variant in (needle, b"PATERNXYZ12", b"PATTERNxYZ12"):
            for delta in (-len(needle), -len(needle) + 1, -4, -2, -1, 0, 1):
                hay = bytearray(chunk_size + 100)
                hay[chunk_size + delta:chunk_size + delta + len(variant)] = variant
                hay = bytes(hay)
                path = tmp_path / "hay.bin"
                path.write_bytes(hay)
                with open(path, "rb") as f:   # a real binary file: the fixed-buffer loop
                    got = _t(find_near_matches_in_file(needle, f, max_l_dist=k, _chunk_size=chunk_size))
                assert got == _emulate(needle, hay, k, chunk_size), (k, variant, delta)
                assert got == _t(find_near_matches(needle, hay, max_l_dist=k))
                # BytesIO has no binary mode flag: the reference (and we) take the generic-file loop
                got2 = _t(find_near_matches_in_file(needle, io.BytesIO(hay), max_l_dist=k, _chunk_size=chunk_size))
                assert got2 == _emulate(needle, hay, k, chunk_size, binary=False) == got
                if chunk_size <= 4096:  # text-mode file object, same result
                    txt = _t(find_near_matches_in_file(needle.decode(), io.StringIO(hay.decode("latin-1")),
                                                       max_l_dist=k, _chunk_size=chunk_size))
                    assert txt == got


def test_file_random_corpus(cuda_device, tmp_path):
    rng = np.random.default_rng(9)
    pat = bytes(rng.integers(97, 101, size=12, dtype=np.uint8))
    data = bytearray(rng.integers(97, 101, size=200_000, dtype=np.uint8).tobytes())
    for pos in range(5000, 195000, 7919):
        data[pos:pos + 12] = pat
    path = tmp_path / "corpus.bin"
    path.write_bytes(bytes(data))
    for chunk in (1000, 65536):
        with open(path, "rb") as f:
            got = _t(find_near_matches_in_file(pat, f, max_l_dist=2, _chunk_size=chunk))
        assert got == _emulate(pat, bytes(data), 2, chunk)
    with open(path, "rb") as f:
        ms = find_near_matches_in_file(pat, f, max_substitutions=1, max_insertions=0, max_deletions=0)
    assert _t(ms) == tup(oracle.substitutions(pat, bytes(data), 1))
    assert all(m.matched == bytes(data[m.start:m.end]) for m in ms)
    with pytest.raises(ValueError):
        find_near_matches_in_file(b"", io.BytesIO(b"abc"), max_l_dist=1)
