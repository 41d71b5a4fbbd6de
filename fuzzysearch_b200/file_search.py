"""find_near_matches_in_file -- the reference's long-input API (fuzzysearch/__init__.py:86-200).

Semantics mirror the reference exactly: the file is searched in chunks of `_chunk_size` items; each
chunk keeps the last ``len(subsequence) - 1 + extra_items_for_chunked_search`` items of the previous
one (__init__.py:135-138, 164-171), every chunk is searched with the chosen search class as an
independent sequence (so window clipping applies at chunk ends, as in the reference), offsets are
re-based, and ONE ``consolidate_matches`` runs over all chunks' matches at the end (:126).

Each chunk search is one upload into the cached device workspace plus the kernels.
"""
import io

from .common import LevenshteinSearchParams, Match

__all__ = ["find_near_matches_in_file"]


def find_near_matches_in_file(subsequence, sequence_file, max_substitutions=None, max_insertions=None,
                              max_deletions=None, max_l_dist=None, _chunk_size=2 ** 20):
    from . import choose_search_class
    search_params = LevenshteinSearchParams(max_substitutions, max_insertions, max_deletions, max_l_dist)
    search_class = choose_search_class(search_params)
    if not subsequence:
        raise ValueError("subsequence must not be empty")
    binary = "b" in getattr(sequence_file, "mode", "") or isinstance(sequence_file, io.RawIOBase)
    keep = len(subsequence) - 1 + search_class.extra_items_for_chunked_search(subsequence, search_params)
    matches = []
    pattern = bytes(bytearray(subsequence)) if binary else subsequence
    tail = None           # items carried over from the previous chunk (same type as the file's items)
    offset = 0            # global index of tail[0]
    while True:
        # binary files refill the fixed-size buffer behind the kept tail (__init__.py:170); any other
        # file object reads a whole new chunk after it (__init__.py:195)
        first = tail is None
        block = sequence_file.read(_chunk_size if (first or not binary) else max(0, _chunk_size - len(tail)))
        if not block:
            break
        chunk = block if first else tail + block
        for match in search_class.search(pattern, chunk, search_params):
            matches.append(Match(match.start + offset, match.end + offset, match.dist, matched=match.matched))
        n_keep = min(keep, len(chunk)) if keep > 0 else 0
        offset += len(chunk) - n_keep
        tail = chunk[len(chunk) - n_keep:]
    return search_class.consolidate_matches(matches)
