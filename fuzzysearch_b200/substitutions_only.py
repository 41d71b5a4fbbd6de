"""fuzzysearch.substitutions_only (substitutions_only.py:10-301), same names.

The reference's three variants (router, "_lp", "_ngrams") return the same list -- every start whose Hamming distance
is within the limit, once, ascending -- so all of them are the one counting-filter search of the library
(fzb_search_hamming); the n-gram variant keeps its own argument check."""
from .search import SubstitutionsOnlySearch, _run

__all__ = ["find_near_matches_substitutions", "find_near_matches_substitutions_lp",
           "find_near_matches_substitutions_ngrams", "has_near_match_substitutions",
           "has_near_match_substitutions_lp", "has_near_match_substitutions_ngrams", "SubstitutionsOnlySearch"]


def _check_arguments(subsequence, sequence, max_substitutions):
    # substitutions_only.py:10-15
    if not len(subsequence):
        raise ValueError("Given subsequence is empty!")
    if max_substitutions is None or max_substitutions < 0:
        raise ValueError("Maximum number of substitutions must be >= 0!")


def find_near_matches_substitutions(subsequence, sequence, max_substitutions):
    _check_arguments(subsequence, sequence, max_substitutions)
    return list(_run(subsequence, sequence, lambda h, p: h.search_hamming(p, max_substitutions), False))


def find_near_matches_substitutions_lp(subsequence, sequence, max_substitutions):
    return find_near_matches_substitutions(subsequence, sequence, max_substitutions)


def find_near_matches_substitutions_ngrams(subsequence, sequence, max_substitutions):
    _check_arguments(subsequence, sequence, max_substitutions)
    if len(subsequence) // (max_substitutions + 1) == 0:  # substitutions_only.py:176-179
        raise ValueError("The subsequence's length must be greater than max_substitutions!")
    return find_near_matches_substitutions(subsequence, sequence, max_substitutions)


def has_near_match_substitutions(subsequence, sequence, max_substitutions):
    """substitutions_only.py:18-34: any match at all?  (chunked early termination, fzb_has_near_match)"""
    from . import has_near_match
    _check_arguments(subsequence, sequence, max_substitutions)
    return has_near_match(subsequence, sequence, max_substitutions=max_substitutions, max_insertions=0,
                          max_deletions=0)


def has_near_match_substitutions_lp(subsequence, sequence, max_substitutions):
    return has_near_match_substitutions(subsequence, sequence, max_substitutions)


def has_near_match_substitutions_ngrams(subsequence, sequence, max_substitutions):
    _check_arguments(subsequence, sequence, max_substitutions)
    if len(subsequence) // (max_substitutions + 1) == 0:
        raise ValueError("The subsequence's length must be greater than max_substitutions!")
    return has_near_match_substitutions(subsequence, sequence, max_substitutions)
