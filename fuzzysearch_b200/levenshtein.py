"""fuzzysearch.levenshtein (levenshtein.py:9-164), same names: every function is ONE call into libfuzzb200.so.

The functions return the RAW (unconsolidated) match stream of their route, like the reference's; ``LevenshteinSearch``
is the search class ``find_near_matches`` uses.  Order: the n-gram route yields in the reference's generation order
(n-gram major, hit index ascending); the LP route yields the same multiset (the reference's order there depends on its
candidate lists)."""
from . import _native
from .common import LevenshteinSearchParams
from .levenshtein_ngram import find_near_matches_levenshtein_ngrams
from .search import LevenshteinSearch, _run

__all__ = ["find_near_matches_levenshtein", "find_near_matches_levenshtein_linear_programming",
           "find_near_matches_levenshtein_ngrams", "LevenshteinSearch"]


def _check(subsequence, max_l_dist):
    if not len(subsequence):
        raise ValueError("Given subsequence is empty!")
    if max_l_dist < 0:
        raise ValueError("Maximum Levenshtein distance must be >= 0!")


def find_near_matches_levenshtein(subsequence, sequence, max_l_dist):
    """levenshtein.py:9-38: exact search for 0, n-grams if len(subsequence) // (max_l_dist + 1) >= 3, else the
    "linear programming" search -- the router lives in fzb_search_levenshtein."""
    _check(subsequence, max_l_dist)
    return list(LevenshteinSearch.search(subsequence, sequence, LevenshteinSearchParams(max_l_dist=max_l_dist)))


def find_near_matches_levenshtein_linear_programming(subsequence, sequence, max_l_dist):
    """levenshtein.py:52-148 (the candidate automaton), whatever the lengths: FZB_F_FORCE_LP."""
    _check(subsequence, max_l_dist)
    flags = _native.F_FORCE_LP | _native.F_NO_FINAL
    return list(_run(subsequence, sequence, lambda h, p: h.search_levenshtein(p, max_l_dist, flags), False))
