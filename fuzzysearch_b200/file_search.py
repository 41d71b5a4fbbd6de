"""find_near_matches_in_file -- the reference's long-input API (fuzzysearch/__init__.py:86-200).

Semantics mirror the reference exactly: the file is searched in chunks of `_chunk_size` items; each chunk keeps
the last ``len(subsequence) - 1 + extra_items_for_chunked_search`` items of the previous one (__init__.py:135-138,
164-171), every chunk is searched with the chosen search class as an independent sequence (so window clipping
applies at chunk ends, as in the reference), offsets are re-based, and ONE ``consolidate_matches`` runs over all
chunks' matches at the end (:126).

Binary files stream through two PAGE-LOCKED chunk buffers (the reference's single reusable bytearray,
__init__.py:141-145, doubled): while the GPU searches chunk i straight out of one pinned buffer (DMA, no staging
copy), a reader thread fills the other with the kept tail + the next block (file reads drop the GIL).
Other file objects (text mode, BytesIO/StringIO without a mode) take the reference's generic loop (:174-200).
"""
import io
from concurrent.futures import ThreadPoolExecutor

from .common import LevenshteinSearchParams, Match

__all__ = ["find_near_matches_in_file"]


def _rebased(match, offset):
    return Match(match.start + offset, match.end + offset, match.dist, matched=match.matched)


def _search_binary_file(pattern, sequence_file, search_params, search_class, chunk_size, keep):
    """__init__.py:129-171 with two pinned buffers and read-ahead."""
    from . import _native
    bufs = [_native.PinnedBuffer(chunk_size), _native.PinnedBuffer(chunk_size)]
    arrays = [b.array for b in bufs]
    matches = []
    pool = ThreadPoolExecutor(max_workers=1)
    try:
        cur = 0
        n_read = sequence_file.readinto(memoryview(arrays[0]))
        offset = 0
        chunk_len = n_read
        while n_read:
            chunk = arrays[cur][:chunk_len]
            n_to_keep = min(keep, chunk_len) if keep > 0 else 0
            nxt = arrays[1 - cur]
            nxt[:n_to_keep] = chunk[chunk_len - n_to_keep:chunk_len]
            pending = pool.submit(sequence_file.readinto, memoryview(nxt)[n_to_keep:])  # read-ahead
            for match in search_class.search(pattern, chunk, search_params):
                matches.append(_rebased(match, offset))
            n_read = pending.result() or 0
            offset += chunk_len - n_to_keep
            chunk_len = n_to_keep + n_read
            cur = 1 - cur
    finally:
        pool.shutdown(wait=True)
        for b in bufs:
            b.close()
    return matches


def _search_generic_file(pattern, sequence_file, search_params, search_class, chunk_size, keep):
    """__init__.py:174-200."""
    matches = []
    chunk = sequence_file.read(chunk_size)
    offset = 0
    while chunk:
        for match in search_class.search(pattern, chunk, search_params):
            matches.append(_rebased(match, offset))
        n_to_keep = min(keep, len(chunk))
        offset += len(chunk) - n_to_keep
        if n_to_keep:
            chunk = chunk[-n_to_keep:] + sequence_file.read(chunk_size)
            if len(chunk) == n_to_keep:
                break
        else:
            chunk = sequence_file.read(chunk_size)
    return matches


def find_near_matches_in_file(subsequence, sequence_file, max_substitutions=None, max_insertions=None,
                              max_deletions=None, max_l_dist=None, _chunk_size=2 ** 20):
    from . import choose_search_class
    search_params = LevenshteinSearchParams(max_substitutions, max_insertions, max_deletions, max_l_dist)
    search_class = choose_search_class(search_params)
    if not subsequence:
        raise ValueError("subsequence must not be empty")
    binary = "b" in getattr(sequence_file, "mode", "") or isinstance(sequence_file, io.RawIOBase)
    keep = len(subsequence) - 1 + search_class.extra_items_for_chunked_search(subsequence, search_params)
    if binary and hasattr(sequence_file, "readinto"):
        pattern = bytes(bytearray(subsequence))
        matches = _search_binary_file(pattern, sequence_file, search_params, search_class, _chunk_size, keep)
    else:
        pattern = bytes(bytearray(subsequence)) if binary else subsequence
        matches = _search_generic_file(pattern, sequence_file, search_params, search_class, _chunk_size, keep)
    return search_class.consolidate_matches(matches)
