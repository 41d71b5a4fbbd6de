"""GPU parity against the committed golden fixtures (outputs of the real reference), through the
C-ABI of libfuzzb200.so.  Every record of tests/golden/*.json that names a hot-path function is
replayed on the device; results must be bit-exact (raw streams) / tie-aware equal (final lists)."""
import pytest

import oracle
from fuzzysearch_b200 import _native, find_near_matches, search_exact
from parity import assert_final_parity, load_golden, tup

pytestmark = pytest.mark.gpu

F = _native


def _b(h):
    return bytes.fromhex(h)


def _replay(records, cuda_device):
    counts = {}
    windowed = 0
    for rec in records:
        fn, a = rec["fn"], rec["args"]
        ctx = "%s%r" % (fn, a)
        if fn in ("expand", "expand_short", "expand_long", "consolidate", "lev_raw", "generic_raw", "subs"):
            continue  # internal helpers: covered through the search entry points below
        if fn == "find_near_matches":
            pat, hay = _b(a[0]), _b(a[1])
            if "exc" in rec:
                with pytest.raises((ValueError, TypeError)):
                    find_near_matches(pat, hay, *a[2:6])
                counts[fn] = counts.get(fn, 0) + 1
                continue
            ours = [(m.start, m.end, m.dist) for m in find_near_matches(pat, hay, *a[2:6])]
            subs, ins, dels, l = oracle.normalize_params(*a[2:6])
            if l == 0 or (ins == 0 and dels == 0):
                assert ours == tup(rec["result"]), ctx
            else:
                _, raw = oracle.find_near_matches(pat, hay, *a[2:6], return_raw=True)
                assert_final_parity(ours, rec["result"], raw, ctx)
            counts[fn] = counts.get(fn, 0) + 1
            continue
        if "exc" in rec:
            continue
        pat, hay = _b(a[0]), _b(a[1])
        if len(pat) == 0 or len(pat) > F.FZB_MAX_PATTERN:
            continue
        exp = rec["result"]
        hs = F.Haystack.from_host(hay)
        try:
            if fn == "lev_ngrams_raw":
                for flags in (F.F_FORCE_NGRAMS | F.F_NO_FINAL,
                              F.F_FORCE_NGRAMS | F.F_NO_FINAL | F.F_FORCE_DENSE):
                    r = hs.search_levenshtein(pat, a[2], flags)
                    assert r.triples(F.RAW) == tup(exp), ctx  # generation order
                    r.close()
            elif fn == "lev_lp_raw":
                r = hs.search_levenshtein(pat, a[2], F.F_FORCE_LP | F.F_NO_FINAL)
                assert sorted(r.triples(F.RAW)) == sorted(tup(exp)), ctx
                r.close()
            elif fn == "generic_lp_raw":
                if a[5] > 63:
                    continue
                r = hs.search_generic(pat, a[2], a[3], a[4], a[5], F.F_FORCE_LP | F.F_NO_FINAL)
                assert sorted(r.triples(F.RAW)) == sorted(tup(exp)), ctx
                r.close()
            elif fn == "generic_ngrams_raw":
                if a[5] > 63:
                    continue
                for flags in (F.F_FORCE_NGRAMS | F.F_NO_FINAL,
                              F.F_FORCE_NGRAMS | F.F_NO_FINAL | F.F_FORCE_DENSE):
                    r = hs.search_generic(pat, a[2], a[3], a[4], a[5], flags)
                    assert sorted(r.triples(F.RAW)) == sorted(tup(exp)), ctx
                    r.close()
            elif fn in ("subs_lp", "subs_ngrams"):
                r = hs.search_hamming(pat, a[2])
                assert r.triples(F.RAW) == tup(exp), ctx
                r.close()
            elif fn == "search_exact":
                if a[2] == 0 and a[3] is None:
                    r = hs.search_exact(pat)
                else:  # the window form (search_exact.py:22-56): a view of the resident buffer
                    r = hs.search_exact(pat, start=a[2], end=a[3])
                    assert search_exact(pat, hay, a[2], a[3]) == list(exp), ctx  # host sequence: sliced first
                    windowed += 1
                assert [s for s, _, _ in r.triples(F.RAW)] == list(exp), ctx
                r.close()
            else:
                raise KeyError(fn)
            counts[fn] = counts.get(fn, 0) + 1
        finally:
            hs.close()
    counts["search_exact_windowed"] = windowed
    return counts


def test_gpu_replays_reference_suite_calls(cuda_device):
    counts = _replay(load_golden("ref_suite_calls.json"), cuda_device)
    for fn in ("lev_ngrams_raw", "lev_lp_raw", "generic_lp_raw", "generic_ngrams_raw", "subs_lp",
               "subs_ngrams", "search_exact", "find_near_matches"):
        assert counts.get(fn, 0) > 0, fn
    assert counts["search_exact_windowed"] > 300


def test_gpu_replays_reference_fuzz(cuda_device):
    counts = _replay(load_golden("ref_fuzz.json"), cuda_device)
    assert sum(counts.values()) > 3000


def replay_route_functions(records):
    """The reference's module-level route functions, under their own names (fuzzysearch_b200.levenshtein,
    .levenshtein_ngram, .substitutions_only, .generic_search): every golden record of a route function is replayed
    through the public Python function and must give the reference's raw stream."""
    from fuzzysearch_b200 import LevenshteinSearchParams
    from fuzzysearch_b200.generic_search import (find_near_matches_generic_linear_programming,
                                                 find_near_matches_generic_ngrams)
    from fuzzysearch_b200.levenshtein import find_near_matches_levenshtein_linear_programming
    from fuzzysearch_b200.levenshtein_ngram import find_near_matches_levenshtein_ngrams
    from fuzzysearch_b200.substitutions_only import (find_near_matches_substitutions_lp,
                                                     find_near_matches_substitutions_ngrams)
    t = lambda ms: [(m.start, m.end, m.dist) for m in ms]  # noqa: E731
    n = 0
    for rec in records:
        fn, a = rec["fn"], rec["args"]
        if "exc" in rec or fn not in ("lev_ngrams_raw", "lev_lp_raw", "generic_lp_raw", "generic_ngrams_raw", "subs_lp",
                                      "subs_ngrams"):
            continue
        pat, hay = _b(a[0]), _b(a[1])
        if len(pat) == 0 or len(pat) > F.FZB_MAX_PATTERN:
            continue
        exp, ctx = tup(rec["result"]), "%s%r" % (fn, a)
        if fn == "lev_ngrams_raw":
            ms = find_near_matches_levenshtein_ngrams(pat, hay, a[2])
            assert t(ms) == exp, ctx  # generation order
            assert all(m.matched == hay[m.start:m.end] for m in ms), ctx
        elif fn == "lev_lp_raw":
            assert sorted(t(find_near_matches_levenshtein_linear_programming(pat, hay, a[2]))) == sorted(exp), ctx
        elif fn in ("generic_lp_raw", "generic_ngrams_raw"):
            if a[5] > 63:
                continue
            params = LevenshteinSearchParams(a[2], a[3], a[4], a[5])
            if params.unpacked != (a[2], a[3], a[4], a[5]):
                continue  # recorded with limits the public parameter object would normalise differently
            call = find_near_matches_generic_linear_programming if fn == "generic_lp_raw" else \
                find_near_matches_generic_ngrams
            assert sorted(t(call(pat, hay, params))) == sorted(exp), ctx
        else:
            call = find_near_matches_substitutions_lp if fn == "subs_lp" else find_near_matches_substitutions_ngrams
            if fn == "subs_ngrams" and len(pat) // (a[2] + 1) == 0:
                continue
            assert t(call(pat, hay, a[2])) == exp, ctx
        n += 1
    return n


def test_gpu_module_level_route_functions(cuda_device):
    n = replay_route_functions(load_golden("ref_suite_calls.json")) + \
        replay_route_functions(load_golden("ref_fuzz.json")[::7])
    assert n > 400
