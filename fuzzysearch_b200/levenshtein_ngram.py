"""fuzzysearch.levenshtein_ngram (levenshtein_ngram.py:159-198), same name: the n-gram route's raw stream."""
from . import _native
from .search import _run

__all__ = ["find_near_matches_levenshtein_ngrams"]


def _as_byte_pair(subsequence, sequence):
    """Both sides as bytes: byte-like input as it is, anything else through the pattern-alphabet reduction every
    algorithm of the path is invariant under (search._make_alphabet)."""
    from .search import _kind, _make_alphabet, _narrow, _rename, _text
    subsequence, sequence = _text(subsequence), _text(sequence)
    kind = _kind(subsequence)
    if _kind(sequence) != kind:
        raise TypeError("subsequence and sequence must both be str or both be byte-like")
    a, b = _narrow(subsequence, kind), _narrow(sequence, kind)
    if a is not None and b is not None:
        return a.tobytes(), b.tobytes()
    alphabet = _make_alphabet([subsequence], kind)
    sub = _rename(subsequence, kind, alphabet).tobytes()
    if kind == "str":
        rank = {c: i + 1 for i, c in enumerate(alphabet)}
        return sub, bytes(rank.get(ord(c), 0) for c in sequence)
    return sub, bytes(alphabet.get(x, 0) for x in sequence)


def _expand_on_device(subsequence, sequence, max_l_dist, variant):
    """One expansion through the device routines of the verify kernels (fzb_debug_expand, a test hook: the search
    itself never expands one candidate per call)."""
    sub, seq = _as_byte_pair(subsequence, sequence)
    row = _native.debug_expand([(sub, seq, int(max_l_dist), variant)])[0].tolist()
    dist, length = row[0:2] if row[0] != -2 else row[4:6]  # bit-parallel path, else the cell-by-cell one
    return (None, None) if dist == -1 else (dist, length)


def _expand(subsequence, sequence, max_l_dist):
    """levenshtein_ngram.py:8-19."""
    return _expand_on_device(subsequence, sequence, max_l_dist, 0)


def _py_expand_short(subsequence, sequence, max_l_dist):
    """levenshtein_ngram.py:22-74 (with its early-break quirk)."""
    return _expand_on_device(subsequence, sequence, max_l_dist, 1)


def _py_expand_long(subsequence, sequence, max_l_dist):
    """levenshtein_ngram.py:77-143."""
    return _expand_on_device(subsequence, sequence, max_l_dist, 2)


_expand_short = _py_expand_short  # levenshtein_ngram.py:146-156 (no C twin here)
_expand_long = _py_expand_long


def find_near_matches_levenshtein_ngrams(subsequence, sequence, max_l_dist):
    """Raw matches in the reference's generation order (duplicates across n-grams included); ValueError when
    len(subsequence) // (max_l_dist + 1) == 0 (levenshtein_ngram.py:163-165, FZB_E_NGRAM_ZERO)."""
    if not len(subsequence):
        raise ValueError("Given subsequence is empty!")
    if len(subsequence) // (max_l_dist + 1) == 0:
        raise ValueError("the subsequence length must be greater than max_l_dist")
    flags = _native.F_FORCE_NGRAMS | _native.F_NO_FINAL
    return list(_run(subsequence, sequence, lambda h, p: h.search_levenshtein(p, max_l_dist, flags), False))
