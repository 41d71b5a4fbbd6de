#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the REAL reference.

Run in the build container only (``/root/reference`` is absent on the GPU box):

    PYTHONHASHSEED=0 python tests/golden/gen_golden.py

Two fixture files are written next to this script:

* ``ref_suite_calls.json`` -- every call the reference's OWN test-suite
  (/root/reference/tests, 584 tests) makes into the hot-path functions, recorded by wrapping
  those functions before the test modules import them; inputs + the reference's output (or the
  exception type).  Only recorded if the whole suite passes, so each record is an input chosen
  by the reference's authors with the output their assertions accepted.
* ``ref_fuzz.json`` -- seeded random cases run through the reference's pure-Python functions
  (configuration "P": no native extension is importable from the read-only source tree).

str inputs are stored latin-1 encoded (hex); list/tuple inputs (non byte-like) are skipped.
"""
import json
import os
import random
import sys
import unittest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, REF)

import fuzzysearch  # noqa: E402
from fuzzysearch import common, generic_search, levenshtein, levenshtein_ngram  # noqa: E402
from fuzzysearch import search_exact as search_exact_mod  # noqa: E402
from fuzzysearch import substitutions_only  # noqa: E402
from fuzzysearch.common import LevenshteinSearchParams, Match  # noqa: E402

assert levenshtein_ngram._expand_short is levenshtein_ngram._py_expand_short, "natives present?"

_REAL_CLASSES = {n: getattr(fuzzysearch, n) for n in
                 ("ExactSearch", "LevenshteinSearch", "SubstitutionsOnlySearch", "GenericSearch")}
RECORDS = []
_SEEN = set()
MAX_RECORD_CHARS = 1500


def enc(x):
    """bytes-like / latin-1 str -> hex, else None."""
    if isinstance(x, (bytes, bytearray)):
        return bytes(x).hex()
    if isinstance(x, str):
        try:
            return x.encode("latin-1").hex()
        except UnicodeEncodeError:
            return None
    return None


def enc_result(r):
    if isinstance(r, Match):
        return [r.start, r.end, r.dist]
    if isinstance(r, (list, tuple)):
        return [enc_result(x) for x in r]
    if isinstance(r, (set, frozenset)):
        return sorted(enc_result(x) for x in r)
    if isinstance(r, (int, bool)) or r is None:
        return r
    raise TypeError(type(r))


def record(fn, args, result=None, exc=None):
    rec = {"fn": fn, "args": args}
    if exc is not None:
        rec["exc"] = exc
    else:
        rec["result"] = result
    key = json.dumps(rec, sort_keys=True)
    if len(key) > MAX_RECORD_CHARS:   # keeps the fixture small (file-API tests use 1 MiB chunks)
        return
    if key not in _SEEN:
        _SEEN.add(key)
        RECORDS.append(rec)


def wrap(module, name, fn_label, argconv, materialize=True):
    orig = getattr(module, name)

    def wrapper(*a, **kw):
        conv = argconv(*a, **kw)
        try:
            res = orig(*a, **kw)
            if materialize and not isinstance(res, (list, tuple, bool, int, type(None))):
                res = list(res)
        except (ValueError, TypeError) as e:
            if conv is not None:
                record(fn_label, conv, exc=type(e).__name__)
            raise
        if conv is not None:
            try:
                record(fn_label, conv, result=enc_result(res))
            except TypeError:
                pass
        return res

    wrapper.__wrapped__ = orig
    setattr(module, name, wrapper)
    return orig


def conv_expand(sub, seq, k):
    s, q = enc(sub), enc(seq)
    if s is None or q is None:
        return None
    return [s, q, k]


def conv_psk(sub, seq, k=None, **kw):
    if k is None:
        k = kw.get("max_l_dist", kw.get("max_substitutions"))
    s, q = enc(sub), enc(seq)
    if s is None or q is None or not isinstance(k, int):
        return None
    return [s, q, k]


def conv_generic(sub, seq, params):
    s, q = enc(sub), enc(seq)
    if s is None or q is None:
        return None
    return [s, q] + list(params.unpacked)


def conv_exact(sub, seq, start_index=0, end_index=None):
    s, q = enc(sub), enc(seq)
    if s is None or q is None:
        return None
    return [s, q, start_index, end_index]


def conv_consolidate(matches):
    try:
        ms = list(matches)
    except TypeError:
        return None
    if not all(isinstance(m, Match) for m in ms):
        return None
    return None  # handled by the dedicated wrapper below (needs the materialised input)


def conv_fnm(sub, seq, max_substitutions=None, max_insertions=None, max_deletions=None,
             max_l_dist=None):
    # tests/test_find_near_matches.py:31-51 patches the four search classes with mocks to test
    # dispatch only; results produced under a mock are not reference outputs -> not recorded
    if any(getattr(fuzzysearch, n) is not c for n, c in _REAL_CLASSES.items()):
        return None
    s, q = enc(sub), enc(seq)
    if s is None or q is None:
        return None
    return [s, q, max_substitutions, max_insertions, max_deletions, max_l_dist]


def install_wrappers():
    wrap(levenshtein_ngram, "_py_expand_short", "expand_short", conv_expand)
    wrap(levenshtein_ngram, "_py_expand_long", "expand_long", conv_expand)
    # _expand_short/_expand_long are aliases bound at import: rebind to the wrapped versions
    levenshtein_ngram._expand_short = levenshtein_ngram._py_expand_short
    levenshtein_ngram._expand_long = levenshtein_ngram._py_expand_long
    wrap(levenshtein_ngram, "_expand", "expand", conv_expand)
    wrap(levenshtein_ngram, "find_near_matches_levenshtein_ngrams", "lev_ngrams_raw", conv_psk)
    levenshtein.find_near_matches_levenshtein_ngrams = \
        levenshtein_ngram.find_near_matches_levenshtein_ngrams
    wrap(levenshtein, "find_near_matches_levenshtein_linear_programming", "lev_lp_raw", conv_psk)
    wrap(levenshtein, "find_near_matches_levenshtein", "lev_raw", conv_psk)
    wrap(generic_search, "_find_near_matches_generic_linear_programming", "generic_lp_raw",
         conv_generic)
    generic_search.find_near_matches_generic_linear_programming = \
        generic_search._find_near_matches_generic_linear_programming
    wrap(generic_search, "find_near_matches_generic_ngrams", "generic_ngrams_raw", conv_generic)
    wrap(generic_search, "find_near_matches_generic", "generic_raw", conv_generic)
    wrap(substitutions_only, "find_near_matches_substitutions_lp", "subs_lp", conv_psk)
    wrap(substitutions_only, "find_near_matches_substitutions_ngrams", "subs_ngrams", conv_psk)
    wrap(substitutions_only, "find_near_matches_substitutions", "subs", conv_psk)
    wrap(search_exact_mod, "search_exact", "search_exact", conv_exact)
    for mod in (levenshtein, levenshtein_ngram, generic_search, substitutions_only):
        mod.search_exact = search_exact_mod.search_exact
    wrap(fuzzysearch, "find_near_matches", "find_near_matches", conv_fnm)

    # consolidate: record (input triples -> output triples); ties are hash-order dependent
    orig_cons = common.consolidate_overlapping_matches

    def cons_wrapper(matches):
        ms = list(matches)
        res = orig_cons(ms)
        if all(isinstance(m, Match) for m in ms):
            record("consolidate", [enc_result(ms)], result=enc_result(res))
        return res

    common.consolidate_overlapping_matches = cons_wrapper
    levenshtein.consolidate_overlapping_matches = cons_wrapper
    generic_search.consolidate_overlapping_matches = cons_wrapper


def harvest_reference_suite():
    install_wrappers()
    loader = unittest.TestLoader()
    suite = loader.discover(os.path.join(REF, "tests"), top_level_dir=REF)
    result = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, "w")).run(suite)
    assert result.wasSuccessful(), (result.errors[:2], result.failures[:2])
    print("reference suite: ran %d tests OK (skipped %d); %d distinct calls recorded"
          % (result.testsRun, len(result.skipped), len(RECORDS)))
    out = {"source": "calls made by /root/reference/tests (fuzzysearch 0.8.1 @ 4f6d9d8), "
                     "pure-Python configuration, PYTHONHASHSEED=0",
           "tests_run": result.testsRun,
           "records": RECORDS}
    with open(os.path.join(HERE, "ref_suite_calls.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


# ---------------------------------------------------------------------------------------------
# seeded fuzz through the unwrapped reference functions
# ---------------------------------------------------------------------------------------------
def unwrapped(f):
    while hasattr(f, "__wrapped__"):
        f = f.__wrapped__
    return f


def rand_bytes(rng, alphabet, n):
    return bytes(rng.choice(alphabet) for _ in range(n))


def mutate(rng, pat, alphabet, nedits):
    s = bytearray(pat)
    for _ in range(nedits):
        op = rng.randrange(3)
        pos = rng.randrange(len(s) + 1) if s else 0
        if op == 0 and s:
            s[min(pos, len(s) - 1)] = rng.choice(alphabet)
        elif op == 1:
            s.insert(pos, rng.choice(alphabet))
        elif s:
            del s[min(pos, len(s) - 1)]
    return bytes(s)


def make_case(rng, mmin=1, mmax=24, nmax=160):
    alphabet = rng.choice([b"ab", b"ACGT", b"ACGT", b"abcdefgh", bytes(range(32, 127))])
    m = rng.randint(mmin, mmax)
    pat = rand_bytes(rng, alphabet, m)
    n = rng.randint(0, nmax)
    hay = bytearray(rand_bytes(rng, alphabet, n))
    for _ in range(rng.randint(0, 4)):
        ins = mutate(rng, pat, alphabet, rng.randint(0, 4))
        pos = rng.randint(0, len(hay))
        if rng.random() < 0.5:
            hay[pos:pos + len(ins)] = ins
        else:
            hay[pos:pos] = ins
    return pat, bytes(hay)


def triples(ms):
    return [[m.start, m.end, m.dist] for m in ms]


def fuzz():
    rng = random.Random(20260923)
    recs = []
    f_short = unwrapped(levenshtein_ngram._py_expand_short)
    f_long = unwrapped(levenshtein_ngram._py_expand_long)
    f_ngr = unwrapped(levenshtein_ngram.find_near_matches_levenshtein_ngrams)
    f_lp = unwrapped(levenshtein.find_near_matches_levenshtein_linear_programming)
    f_glp = unwrapped(generic_search._find_near_matches_generic_linear_programming)
    f_gng = unwrapped(generic_search.find_near_matches_generic_ngrams)
    f_slp = unwrapped(substitutions_only.find_near_matches_substitutions_lp)
    f_sng = unwrapped(substitutions_only.find_near_matches_substitutions_ngrams)
    f_fnm = unwrapped(fuzzysearch.find_near_matches)
    # the unwrapped functions still call wrapped module globals internally; harmless.

    # expansions
    for _ in range(1500):
        alphabet = rng.choice([b"ab", b"ACGT", b"abcdefgh"])
        sub = rand_bytes(rng, alphabet, rng.randint(0, 30))
        seq = mutate(rng, sub, alphabet, rng.randint(0, 5)) + rand_bytes(rng, alphabet,
                                                                        rng.randint(0, 6))
        if rng.random() < 0.2:
            seq = rand_bytes(rng, alphabet, rng.randint(0, 34))
        k = rng.randint(0, 6)
        for name, f in (("expand_short", f_short), ("expand_long", f_long)):
            r = f(sub, seq, k)
            recs.append({"fn": name, "args": [sub.hex(), seq.hex(), k], "result": list(r)})

    # Levenshtein n-grams + LP raw streams, final lists
    for _ in range(700):
        pat, hay = make_case(rng)
        m = len(pat)
        k = rng.randint(1, 5)
        if m // (k + 1) >= 1:
            r = triples(f_ngr(pat, hay, k))
            recs.append({"fn": "lev_ngrams_raw", "args": [pat.hex(), hay.hex(), k], "result": r})
        if rng.random() < 0.6:
            pat2, hay2 = make_case(rng, 1, 12, 60)
            k2 = rng.randint(0, 4)
            r = triples(f_lp(pat2, hay2, k2))
            recs.append({"fn": "lev_lp_raw", "args": [pat2.hex(), hay2.hex(), k2], "result": r})
        r = triples(f_fnm(pat, hay, max_l_dist=k))
        recs.append({"fn": "find_near_matches",
                     "args": [pat.hex(), hay.hex(), None, None, None, k], "result": r})

    # generic
    for _ in range(400):
        pat, hay = make_case(rng, 1, 16, 70)
        subs, ins, dels = rng.randint(0, 3), rng.randint(0, 3), rng.randint(0, 3)
        l = rng.choice([None, rng.randint(0, 4)])
        params = LevenshteinSearchParams(subs, ins, dels, l)
        up = list(params.unpacked)
        r = triples(f_glp(pat, hay, params))
        recs.append({"fn": "generic_lp_raw", "args": [pat.hex(), hay.hex()] + up, "result": r})
        if len(pat) // (up[3] + 1) >= 1:
            r = triples(f_gng(pat, hay, params))
            recs.append({"fn": "generic_ngrams_raw", "args": [pat.hex(), hay.hex()] + up,
                         "result": r})
        r = triples(f_fnm(pat, hay, subs, ins, dels, l))
        recs.append({"fn": "find_near_matches",
                     "args": [pat.hex(), hay.hex(), subs, ins, dels, l], "result": r})

    # substitutions only
    for _ in range(400):
        pat, hay = make_case(rng, 1, 24, 120)
        k = rng.randint(0, 5)
        r = triples(f_slp(pat, hay, k))
        recs.append({"fn": "subs_lp", "args": [pat.hex(), hay.hex(), k], "result": r})
        if len(pat) // (k + 1) >= 1:
            r = triples(f_sng(pat, hay, k))
            recs.append({"fn": "subs_ngrams", "args": [pat.hex(), hay.hex(), k], "result": r})
        r = triples(f_fnm(pat, hay, max_substitutions=k, max_insertions=0, max_deletions=0))
        recs.append({"fn": "find_near_matches", "args": [pat.hex(), hay.hex(), k, 0, 0, None],
                     "result": r})

    recs = [r for r in recs if len(json.dumps(r)) <= MAX_RECORD_CHARS]
    print("fuzz: %d records" % len(recs))
    out = {"source": "seeded fuzz (random.Random(20260923)) through the pure-Python reference "
                     "(fuzzysearch 0.8.1 @ 4f6d9d8), PYTHONHASHSEED=0",
           "records": recs}
    with open(os.path.join(HERE, "ref_fuzz.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    harvest_reference_suite()
    fuzz()
