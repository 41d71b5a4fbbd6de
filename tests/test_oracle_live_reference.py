"""Live fuzz of the CPU oracle against the REAL reference, only where /root/reference exists (the
build container).  Complements the committed golden fixtures with fresh random cases every run."""
import os
import random
import sys

import pytest

import oracle
from parity import assert_final_parity, tup

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference sources not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF_SRC)
    try:
        import fuzzysearch
        from fuzzysearch import generic_search, levenshtein, levenshtein_ngram, substitutions_only
        from fuzzysearch.common import LevenshteinSearchParams
        yield dict(fz=fuzzysearch, gen=generic_search, lev=levenshtein, ngr=levenshtein_ngram,
                   subs=substitutions_only, params=LevenshteinSearchParams)
    finally:
        sys.path.remove(REF_SRC)
        for name in [n for n in sys.modules if n == "fuzzysearch" or n.startswith("fuzzysearch.")]:
            del sys.modules[name]


def _case(rng, mmax=28, nmax=220):
    alphabet = rng.choice([b"ab", b"ACGT", b"abcdefgh", bytes(range(32, 127))])
    m = rng.randint(1, mmax)
    pat = bytes(rng.choice(alphabet) for _ in range(m))
    hay = bytearray(rng.choice(alphabet) for _ in range(rng.randint(0, nmax)))
    for _ in range(rng.randint(0, 4)):
        v = bytearray(pat)
        for _ in range(rng.randint(0, 4)):
            op = rng.randrange(3)
            if op == 0 and v:
                v[rng.randrange(len(v))] = rng.choice(alphabet)
            elif op == 1:
                v.insert(rng.randrange(len(v) + 1), rng.choice(alphabet))
            elif v:
                del v[rng.randrange(len(v))]
        pos = rng.randint(0, len(hay))
        hay[pos:pos + len(v)] = v
    return pat, bytes(hay)


def _t(ms):
    return [(m.start, m.end, m.dist) for m in ms]


def test_live_fuzz_against_reference(ref):
    rng = random.Random(int.from_bytes(os.urandom(4), "little"))
    for _ in range(1500):
        pat, hay = _case(rng)
        k = rng.randint(0, 5)
        ctx = (pat, hay, k)
        if len(pat) // (k + 1) >= 1:
            assert tup(oracle.levenshtein_ngrams_raw(pat, hay, k)) == \
                _t(ref["ngr"].find_near_matches_levenshtein_ngrams(pat, hay, k)), ctx
        if len(pat) <= 12 and k <= 3:
            assert tup(oracle.levenshtein_lp_raw(pat, hay, k)) == \
                _t(ref["lev"].find_near_matches_levenshtein_linear_programming(pat, hay, k)), ctx
        assert tup(oracle.subs_lp(pat, hay, k)) == _t(ref["subs"].find_near_matches_substitutions_lp(pat, hay, k)), ctx
        if len(pat) // (k + 1) >= 1:
            assert tup(oracle.subs_ngrams(pat, hay, k)) == \
                _t(ref["subs"].find_near_matches_substitutions_ngrams(pat, hay, k)), ctx
        if k > 0:
            final, raw = oracle.find_near_matches(pat, hay, max_l_dist=k, return_raw=True)
            assert_final_parity(final, _t(ref["fz"].find_near_matches(pat, hay, max_l_dist=k)), raw, repr(ctx))
    for _ in range(400):
        pat, hay = _case(rng, 16, 80)
        subs, ins, dels = rng.randint(0, 3), rng.randint(0, 3), rng.randint(0, 3)
        l = rng.choice([None, rng.randint(0, 4)])
        params = ref["params"](subs, ins, dels, l)
        up = params.unpacked
        assert tup(oracle.generic_lp_raw(pat, hay, *up)) == \
            _t(ref["gen"]._find_near_matches_generic_linear_programming(pat, hay, params)), (pat, hay, up)
        if len(pat) // (up[3] + 1) >= 1:
            assert tup(oracle.generic_ngrams_raw(pat, hay, *up)) == \
                _t(ref["gen"].find_near_matches_generic_ngrams(pat, hay, params)), (pat, hay, up)
