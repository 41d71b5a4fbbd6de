"""Tie-aware parity helpers (finding F5).

The reference picks the winner of a group of overlapping matches with ``min()`` over a ``set``
(common.py:180-189); ties on (dist, -(end-start)) are broken by set iteration order, which depends
on PYTHONHASHSEED.  So a final list is *admissible* iff it contains exactly one member per group
and that member minimises (dist, -(end-start)) within its group.
"""
import json
import os

import numpy as np

import oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)["records"]


def tup(rows):
    return [tuple(int(x) for x in r) for r in rows]


def admissible_sets(raw):
    """raw triples -> list of sets of admissible winners, one per group (literal grouping)."""
    raw = np.asarray(raw, dtype=np.int64).reshape(-1, 3)
    if raw.shape[0] == 0:
        return []
    _, groups = oracle.consolidate(raw, with_groups=True)
    by_group = {}
    for row, g in zip(tup(raw), groups):
        by_group.setdefault(int(g), []).append(row)
    out = []
    for members in by_group.values():
        best = min((d, -(e - s)) for s, e, d in members)
        out.append({(s, e, d) for s, e, d in members if (d, -(e - s)) == best})
    return out


def is_admissible_final(final, raw):
    """True iff `final` has exactly one admissible winner per group of `raw`, sorted."""
    final = tup(final)
    sets = admissible_sets(raw)
    if len(final) != len(sets):
        return False
    if final != sorted(final):
        return False
    used = [False] * len(sets)
    for x in final:
        for i, s in enumerate(sets):
            if not used[i] and x in s:
                used[i] = True
                break
        else:
            return False
    return True


def assert_final_parity(ours, ref_final, raw, ctx=""):
    """ours and the reference's final list must both be admissible for `raw`; where every group
    has a unique minimiser they must be identical."""
    ours, ref_final = tup(ours), tup(ref_final)
    assert is_admissible_final(ref_final, raw), "oracle raw stream does not explain reference: " + ctx
    assert is_admissible_final(ours, raw), "not admissible: %s\nours=%r\nref=%r" % (ctx, ours, ref_final)
    if all(len(s) == 1 for s in admissible_sets(raw)):
        assert ours == ref_final, ctx
