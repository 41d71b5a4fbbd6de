"""CPU-only tests: the host-side mirror of the reference interface, the C-ABI surface, and the
native consolidation (pure host code inside libfuzzb200.so) against the oracle."""
import os
import re

import numpy as np
import pytest

import oracle
from fuzzysearch_b200 import (ExactSearch, GenericSearch, LevenshteinSearch, LevenshteinSearchParams,
                              Match, SubstitutionsOnlySearch, _native, choose_search_class,
                              find_near_matches)
from parity import assert_final_parity, load_golden, tup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fuzzb200.h")).read()
    declared = set(re.findall(r"\b(fzb_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 20
    lib = _native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    assert lib.fzb_version() == 100


def test_header_is_plain_c(tmp_path):
    """include/fuzzb200.h must be consumable from C (the boundary is a C-ABI)."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "fuzzb200.h"\nint main(void) { fzb_stats s; (void)s; return fzb_version() == FZB_VERSION ? 0 : 1; }\n')
    exe = tmp_path / "use_header"
    lib_dir = os.path.join(ROOT, "fuzzysearch_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", lib_dir, "-l:libfuzzb200.so", "-Wl,-rpath," + lib_dir])
    assert subprocess.call([str(exe)]) == 0


def test_match_semantics():
    # common.py:15-32: eq/hash/order on (start,end,dist); matched excluded; frozen
    a, b = Match(1, 3, 0, matched=b"xy"), Match(1, 3, 0, matched=b"zz")
    assert a == b and hash(a) == hash(b) and len({a, b}) == 1
    assert Match(1, 3, 0, b"a") < Match(1, 3, 1, b"a") < Match(1, 4, 0, b"a") < Match(2, 2, 0, b"")
    assert sorted([Match(5, 6, 0, b"a"), Match(1, 9, 2, b"a")])[0].start == 1
    assert repr(a) == "Match(start=1, end=3, dist=0, matched=b'xy')"
    with pytest.raises(AttributeError):
        a.start = 4
    for bad in [(-1, 2, 0), (3, 2, 0), (1, 2, -1)]:
        with pytest.raises(ValueError):
            Match(*bad, matched=b"")
    with pytest.raises(TypeError):  # `matched` has no default in the reference (common.py:20)
        Match(1, 2, 0)
    with pytest.raises(ValueError):
        Match(1, 2, 0, matched=None)
    # the reference's Match is an attrs class and its own file search relies on it (__init__.py:160-162)
    import attr
    import pickle
    from fuzzysearch_b200.common import _PlainMatch
    assert attr.evolve(a, start=0, end=2) == Match(0, 2, 0, b"xy") and attr.evolve(a, start=0).matched == b"xy"
    assert attr.asdict(a) == {"start": 1, "end": 3, "dist": 0, "matched": b"xy"}
    assert [f.name for f in attr.fields(Match)] == ["start", "end", "dist", "matched"]
    assert pickle.loads(pickle.dumps(a)) == a and pickle.loads(pickle.dumps(a)).matched == b"xy"
    # the attrs-free twin (used only where attrs is not installed) behaves the same
    pa, pb = _PlainMatch(1, 3, 0, matched=b"xy"), _PlainMatch(1, 3, 0, matched=b"zz")
    assert pa == pb and hash(pa) == hash(pb) and repr(pa) == repr(a) and pa < _PlainMatch(1, 3, 1, b"")
    with pytest.raises(AttributeError):
        pa.start = 4
    with pytest.raises(ValueError):
        _PlainMatch(1, 2, 0)


def test_params_normalisation_matches_oracle_restatement():
    vals = [None, 0, 1, 2, 5]
    for s in vals:
        for i in vals:
            for d in vals:
                for l in vals:
                    try:
                        exp = oracle.normalize_params(s, i, d, l)
                    except (ValueError, TypeError) as e:
                        with pytest.raises(type(e)):
                            LevenshteinSearchParams(s, i, d, l)
                        continue
                    assert LevenshteinSearchParams(s, i, d, l).unpacked == exp
    for bad in [(-1, None, None, 2), ("1", 1, 1, 1), (1.5, 1, 1, None)]:
        with pytest.raises(TypeError):
            LevenshteinSearchParams(*bad)
    with pytest.raises(ValueError):
        LevenshteinSearchParams()
    with pytest.raises(ValueError):
        LevenshteinSearchParams(max_substitutions=1)


def test_choose_search_class():
    # tests/test_find_near_matches.py:53-199 (dispatch rules of __init__.py:60-83)
    c = lambda *a: choose_search_class(LevenshteinSearchParams(*a))  # noqa: E731
    assert c(None, None, None, 0) is ExactSearch
    assert c(0, 0, 0, None) is ExactSearch
    assert c(1, 0, 0, None) is SubstitutionsOnlySearch
    assert c(3, 0, 0, 2) is SubstitutionsOnlySearch
    assert c(None, None, None, 2) is LevenshteinSearch
    assert c(2, 2, 2, 2) is LevenshteinSearch
    assert c(3, 4, 5, 2) is LevenshteinSearch
    assert c(1, 2, 2, 2) is GenericSearch
    assert c(2, 1, 1, None) is GenericSearch
    assert c(2, 0, 1, 3) is GenericSearch


def test_no_gpu_fails_loudly():
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_native.CudaError):
        find_near_matches(b"abc", b"xxabcxx", max_l_dist=1)


def test_argument_errors_before_any_device_work():
    with pytest.raises(ValueError):
        find_near_matches(b"", b"TEXT", max_l_dist=1)
    with pytest.raises(TypeError):  # str / byte-like / item sequences do not mix (str.find(bytes) in the reference)
        find_near_matches(b"a", "aሴ", max_l_dist=1)
    with pytest.raises(TypeError):
        find_near_matches(["a"], "ab", max_l_dist=1)
    with pytest.raises(TypeError):  # items are numbered through a dict: they must be hashable
        find_near_matches([[1]], [[1], [2]], max_l_dist=1)
    from fuzzysearch_b200 import search_exact
    with pytest.raises(ValueError):
        search_exact(b"", b"abc")


def _random_raw(rng, n, span, with_empty):
    out = []
    for _ in range(n):
        s = int(rng.integers(0, span))
        ln = int(rng.integers(0 if with_empty else 1, 9))
        out.append((s, s + ln, int(rng.integers(0, 4))))
    return out


def test_native_consolidate_equals_literal_grouping():
    """fzb_consolidate (sort + sweep) vs the oracle's literal group_matches (common.py:145-189),
    incl. empty matches, duplicates and dense chains."""
    rng = np.random.default_rng(123)
    for trial in range(400):
        raw = _random_raw(rng, int(rng.integers(0, 40)), int(rng.integers(5, 120)), trial % 2 == 0)
        if trial % 7 == 0:
            raw = raw + raw[:3]
        a = np.array(raw, dtype=np.int64).reshape(-1, 3)
        s, e, d = _native.consolidate(a[:, 0], a[:, 1], a[:, 2].astype(np.int32))
        ours = list(zip(s.tolist(), e.tolist(), d.tolist()))
        assert ours == tup(oracle.consolidate(a)), raw


def test_native_consolidate_on_reference_records():
    n = 0
    for rec in load_golden("ref_suite_calls.json"):
        if rec["fn"] != "consolidate":
            continue
        raw = np.array(rec["args"][0], dtype=np.int64).reshape(-1, 3)
        s, e, d = _native.consolidate(raw[:, 0], raw[:, 1], raw[:, 2].astype(np.int32))
        assert_final_parity(list(zip(s.tolist(), e.tolist(), d.tolist())), rec["result"], raw)
        n += 1
    assert n > 50


def test_synth_host_is_counter_based():
    a = _native.synth_host(0, 1000, b"ACGT", 42)
    b = _native.synth_host(333, 100, b"ACGT", 42)
    assert a[333:433].tobytes() == b.tobytes()
    assert set(a.tolist()) == set(b"ACGT")
    assert _native.synth_host(0, 64, b"ACGT", 43).tobytes() != a[:64].tobytes()


def test_merge_groups_equals_whole_consolidation():
    """fzb_consolidate_groups per shard + fzb_merge_groups == consolidation of the whole raw list
    (random intervals incl. empty ones, random anchor partitions into 3 shards)."""
    rng = np.random.default_rng(5)
    for trial in range(400):
        n = int(rng.integers(0, 60))
        span = int(rng.integers(10, 200))
        raw = []
        for _ in range(n):
            s = int(rng.integers(0, span))
            ln = int(rng.integers(0 if trial % 2 else 1, 9))
            raw.append((s, s + ln, int(rng.integers(0, 4))))
        raw = np.array(raw, dtype=np.int64).reshape(-1, 3)
        whole = tup(oracle.consolidate(raw))
        cuts = sorted(rng.integers(0, span + 1, size=2).tolist())
        parts = []
        for lo, hi in zip([0] + cuts, cuts + [span + 10]):
            sub = raw[(raw[:, 0] >= lo) & (raw[:, 0] < hi)]
            parts.append(_native.consolidate_groups(sub[:, 0], sub[:, 1], sub[:, 2].astype(np.int32)))
        assert _native.merge_groups(np.concatenate(parts, axis=0)) == whole, (raw.tolist(), cuts)


def test_consolidate_uses_the_device_side_final_list():
    """A RawMatches-like object carrying `.final` short-circuits the host consolidation."""
    from fuzzysearch_b200.common import consolidate_overlapping_matches

    class Carrier(list):
        final = None

    c = Carrier([Match(0, 3, 1, b"abc"), Match(1, 4, 0, b"bcd")])
    sentinel = [Match(7, 9, 0, b"zz")]
    c.final = sentinel
    assert consolidate_overlapping_matches(c) == sentinel
    assert consolidate_overlapping_matches(list(c)) == [Match(1, 4, 0, b"bcd")]
