"""fuzzysearch.generic_search (generic_search.py:25-273), same names: raw match streams of the per-operation-limit
search, ONE call into libfuzzb200.so each."""
from . import _native
from .search import GenericSearch, _run

__all__ = ["find_near_matches_generic", "find_near_matches_generic_linear_programming",
           "find_near_matches_generic_ngrams", "has_near_match_generic_ngrams", "GenericSearch"]


def _limits(subsequence, search_params):
    if not len(subsequence):
        raise ValueError("Given subsequence is empty!")
    return search_params.unpacked


def find_near_matches_generic(subsequence, sequence, search_params):
    """generic_search.py:25-54: exact search for max_l_dist 0, n-grams if the n-gram length is >= 3, else LP."""
    _limits(subsequence, search_params)
    return list(GenericSearch.search(subsequence, sequence, search_params))


def find_near_matches_generic_linear_programming(subsequence, sequence, search_params):
    """generic_search.py:57-177, whatever the lengths (FZB_F_FORCE_LP)."""
    subs, ins, dels, l = _limits(subsequence, search_params)
    flags = _native.F_FORCE_LP | _native.F_NO_FINAL
    return list(_run(subsequence, sequence, lambda h, p: h.search_generic(p, subs, ins, dels, l, flags), False))


_find_near_matches_generic_linear_programming = find_near_matches_generic_linear_programming  # generic_search.py:57,180-185


def find_near_matches_generic_ngrams(subsequence, sequence, search_params):
    """generic_search.py:198-237; ValueError when len(subsequence) // (max_l_dist + 1) == 0 (:213-215)."""
    subs, ins, dels, l = _limits(subsequence, search_params)
    if len(subsequence) // (l + 1) == 0:
        raise ValueError("the subsequence length must be greater than max_l_dist")
    flags = _native.F_FORCE_NGRAMS | _native.F_NO_FINAL
    return list(_run(subsequence, sequence, lambda h, p: h.search_generic(p, subs, ins, dels, l, flags), False))


def has_near_match_generic_ngrams(subsequence, sequence, search_params):
    """generic_search.py:240-253."""
    return len(find_near_matches_generic_ngrams(subsequence, sequence, search_params)) > 0
