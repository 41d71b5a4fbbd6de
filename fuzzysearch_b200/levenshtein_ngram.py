"""fuzzysearch.levenshtein_ngram (levenshtein_ngram.py:159-198), same name: the n-gram route's raw stream."""
from . import _native
from .search import _run

__all__ = ["find_near_matches_levenshtein_ngrams"]


def find_near_matches_levenshtein_ngrams(subsequence, sequence, max_l_dist):
    """Raw matches in the reference's generation order (duplicates across n-grams included); ValueError when
    len(subsequence) // (max_l_dist + 1) == 0 (levenshtein_ngram.py:163-165, FZB_E_NGRAM_ZERO)."""
    if not len(subsequence):
        raise ValueError("Given subsequence is empty!")
    if len(subsequence) // (max_l_dist + 1) == 0:
        raise ValueError("the subsequence length must be greater than max_l_dist")
    flags = _native.F_FORCE_NGRAMS | _native.F_NO_FINAL
    return list(_run(subsequence, sequence, lambda h, p: h.search_levenshtein(p, max_l_dist, flags), False))
