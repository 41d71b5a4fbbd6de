"""Result / parameter types with the reference's semantics (fuzzysearch/common.py).

``Match`` mirrors fuzzysearch.common.Match (common.py:15-32): a frozen record
``(start, end, dist, matched)`` whose equality, hash and ordering use ``(start, end, dist)`` only.
``LevenshteinSearchParams`` mirrors common.py:35-116 (validation + normalisation).
"""

__all__ = ["Match", "LevenshteinSearchParams", "FuzzySearchBase", "consolidate_overlapping_matches", "group_matches",
           "GroupOfMatches", "get_best_match_in_group", "count_differences_with_maximum"]


def _check_match_fields(start, end, dist, matched):  # common.py:21-32
    if not (isinstance(start, int) and start >= 0):
        raise ValueError("start must be a non-negative integer")
    if not (isinstance(end, int) and end >= start):
        raise ValueError("end must be an integer no smaller than start")
    if not (isinstance(dist, int) and dist >= 0):
        raise ValueError("dist must be a non-negative integer")
    if matched is None:
        raise ValueError("matched must be supplied")


class _PlainMatch(object):
    """fuzzysearch.common.Match (common.py:15-32) without the attrs dependency."""
    __slots__ = ("start", "end", "dist", "matched")

    def __init__(self, start, end, dist, matched=None):
        if __debug__:
            _check_match_fields(start, end, dist, matched)
        object.__setattr__(self, "start", start)
        object.__setattr__(self, "end", end)
        object.__setattr__(self, "dist", dist)
        object.__setattr__(self, "matched", matched)

    def __setattr__(self, name, value):
        raise AttributeError("Match is frozen")

    __delattr__ = __setattr__

    def _key(self):
        return (self.start, self.end, self.dist)

    def __eq__(self, other):
        if other.__class__ is not self.__class__:
            return NotImplemented
        return self._key() == other._key()

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __lt__(self, other):
        if other.__class__ is not self.__class__:
            return NotImplemented
        return self._key() < other._key()

    def __le__(self, other):
        if other.__class__ is not self.__class__:
            return NotImplemented
        return self._key() <= other._key()

    def __gt__(self, other):
        if other.__class__ is not self.__class__:
            return NotImplemented
        return self._key() > other._key()

    def __ge__(self, other):
        if other.__class__ is not self.__class__:
            return NotImplemented
        return self._key() >= other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return "Match(start=%r, end=%r, dist=%r, matched=%r)" % (self.start, self.end, self.dist,
                                                                 self.matched)

    def __getstate__(self):
        return (self.start, self.end, self.dist, self.matched)

    def __setstate__(self, state):
        for name, value in zip(self.__slots__, state):
            object.__setattr__(self, name, value)


def _make_attrs_match(attr):
    """The same record as an attrs class, as in the reference (common.py:15-32: frozen, slots; eq / hash / order
    over (start, end, dist) only), so that code written against the reference keeps working on our results:
    ``attr.evolve(match, start=...)`` (the reference's own file search does this, __init__.py:160-162),
    ``attr.asdict``, ``attr.fields(Match)``."""
    def post_init(self):
        _check_match_fields(self.start, self.end, self.dist, self.matched)

    body = {
        "start": attr.ib(type=int, eq=True, hash=True),
        "end": attr.ib(type=int, eq=True, hash=True),
        "dist": attr.ib(type=int, eq=True, hash=True),
        "matched": attr.ib(eq=False, hash=False),
    }
    cls = type("Match", (object,), dict(body, **({"__attrs_post_init__": post_init} if __debug__ else {})))
    cls.__doc__ = "fuzzysearch.common.Match (common.py:15-32)."
    cls.__module__ = __name__
    cls.__qualname__ = "Match"
    return attr.s(frozen=True, slots=True)(cls)


try:  # attrs is the reference's one runtime dependency (setup.py:131): present wherever the reference runs
    import attr as _attr
except ImportError:  # pragma: no cover
    Match = _PlainMatch
    Match.__name__ = Match.__qualname__ = "Match"
else:
    Match = _make_attrs_match(_attr)


class LevenshteinSearchParams(object):
    """Parameter record for Levenshtein-distance searches (common.py:35-116)."""
    __slots__ = ("max_substitutions", "max_insertions", "max_deletions", "max_l_dist")

    def __init__(self, max_substitutions=None, max_insertions=None, max_deletions=None, max_l_dist=None):
        self._check_params_valid(max_substitutions, max_insertions, max_deletions, max_l_dist)
        subs, ins, dels, l = self._normalize_params(max_substitutions, max_insertions, max_deletions,
                                                    max_l_dist)
        object.__setattr__(self, "max_substitutions", subs)
        object.__setattr__(self, "max_insertions", ins)
        object.__setattr__(self, "max_deletions", dels)
        object.__setattr__(self, "max_l_dist", l)

    def __setattr__(self, name, value):
        raise AttributeError("LevenshteinSearchParams is frozen")

    @property
    def unpacked(self):
        return (self.max_substitutions, self.max_insertions, self.max_deletions, self.max_l_dist)

    def __eq__(self, other):
        return other.__class__ is self.__class__ and self.unpacked == other.unpacked

    def __hash__(self):
        return hash(self.unpacked)

    def __repr__(self):
        return ("LevenshteinSearchParams(max_substitutions=%r, max_insertions=%r, max_deletions=%r, "
                "max_l_dist=%r)" % self.unpacked)

    @staticmethod
    def _check_params_valid(max_substitutions, max_insertions, max_deletions, max_l_dist):
        # common.py:61-85
        if not all(x is None or (isinstance(x, int) and x >= 0)
                   for x in [max_substitutions, max_insertions, max_deletions, max_l_dist]):
            raise TypeError("All limits must be positive integers or None.")
        if max_l_dist is None:
            n_limits = ((1 if max_substitutions is not None else 0) +
                        (1 if max_insertions is not None else 0) +
                        (1 if max_deletions is not None else 0))
            if n_limits < 3:
                if n_limits == 0:
                    raise ValueError("No limitations given!")
                elif max_substitutions is None:
                    raise ValueError("# substitutions must be limited!")
                elif max_insertions is None:
                    raise ValueError("# insertions must be limited!")
                elif max_deletions is None:
                    raise ValueError("# deletions must be limited!")

    @staticmethod
    def _normalize_params(max_substitutions, max_insertions, max_deletions, max_l_dist):
        # common.py:87-116
        maxes_sum = sum(x if x is not None else 1 << 29
                        for x in [max_substitutions, max_insertions, max_deletions])
        if max_l_dist is None:
            return (max_substitutions, max_insertions, max_deletions, maxes_sum)

        def _normalize(param):
            return min(param, max_l_dist) if param is not None else max_l_dist
        return (_normalize(max_substitutions), _normalize(max_insertions), _normalize(max_deletions),
                min(max_l_dist, maxes_sum))


class FuzzySearchBase(object):
    """Abstract base class of the search classes (common.py:192-209)."""

    @classmethod
    def search(cls, subsequence, sequence, search_params):
        raise NotImplementedError

    @classmethod
    def consolidate_matches(cls, matches):
        materialize = getattr(matches, "materialize", None)  # search.RawMatches -> a plain list
        if materialize is not None:
            return materialize()
        try:
            len(matches)
        except TypeError:
            return list(matches)
        else:
            return matches

    @classmethod
    def extra_items_for_chunked_search(cls, subsequence, search_params):
        raise NotImplementedError


def consolidate_overlapping_matches(matches):
    """common.py:185-189 via the native O(N log N) sweep (fzb_consolidate).

    Groups are the connected components of interval overlap; the winner of a group minimises
    (dist, -(end-start)); ties -- which the reference leaves to set-iteration order -- go to the
    smallest (start, end).  Returns a list sorted by (start, end, dist)."""
    from . import _native
    import numpy as np
    precomputed = getattr(matches, "final", None)  # RawMatches: the device already consolidated
    if precomputed is not None:
        return list(precomputed)
    matches = list(matches)
    if not matches:
        return []
    start = np.fromiter((m.start for m in matches), dtype=np.int64, count=len(matches))
    end = np.fromiter((m.end for m in matches), dtype=np.int64, count=len(matches))
    dist = np.fromiter((m.dist for m in matches), dtype=np.int32, count=len(matches))
    os_, oe, od = _native.consolidate(start, end, dist)
    by_key = {}
    for m in matches:
        by_key.setdefault((m.start, m.end, m.dist), m)
    return [by_key[(int(s), int(e), int(d))] for s, e, d in zip(os_, oe, od)]


# ---- small host-side helpers of fuzzysearch.common that callers (and the reference's tests) import by name ----------
def count_differences_with_maximum(sequence1, sequence2, max_differences):
    """common.py:119-126: the number of positions at which the two sequences differ, counting no further than
    `max_differences` (a limit that is never REACHED -- zero or negative -- stops nothing, as in the reference)."""
    total = sum(1 for a, b in zip(sequence1, sequence2) if a != b)
    return min(total, max_differences) if max_differences > 0 else total


class GroupOfMatches(object):
    """common.py:145-158: a set of matches and the hull [start, end) they span."""

    def __init__(self, match):
        assert match.start <= match.end
        self.start, self.end, self.matches = match.start, match.end, {match}

    def is_match_in_group(self, match):
        return match.end > self.start and match.start < self.end

    def add_match(self, match):
        self.matches.add(match)
        self.start, self.end = min(self.start, match.start), max(self.end, match.end)


def group_matches(matches):
    """common.py:161-177: the groups (sets) of transitively overlapping matches, in order of first appearance.
    Same incremental definition as the reference -- a match joins the groups whose CURRENT hull it overlaps and
    fuses them -- which is what consolidate_overlapping_matches's native sweep computes for whole lists."""
    groups = []
    for match in matches:
        touched = [g for g in groups if g.is_match_in_group(match)]
        if len(touched) == 1:
            touched[0].add_match(match)
            continue
        fused = GroupOfMatches(match)  # a group of its own, or the union of every group it bridges
        for g in touched:
            for m in g.matches:
                fused.add_match(m)
        groups = [g for g in groups if all(g is not t for t in touched)]
        groups.append(fused)
    return [g.matches for g in groups]


def get_best_match_in_group(group):
    """common.py:180-182: smallest distance, then longest; further ties go to the smallest (start, end) here."""
    return min(group, key=lambda m: (m.dist, -(m.end - m.start), m.start, m.end))
