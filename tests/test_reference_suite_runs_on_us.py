"""The REFERENCE'S OWN test-suite, unmodified, against this package -- container only (needs /root/reference/tests).

`fuzzysearch` is aliased to `fuzzysearch_b200` (package and the submodules the tests import: common, levenshtein,
levenshtein_ngram, substitutions_only, generic_search, search_exact) and the reference's unittest modules are loaded
from /root/reference/tests and run.  Every search goes through the C-ABI: of the emulated build of the product
sources here (tests/emu), of libfuzzb200.so on a B200 when run with `-m gpu`-style access to a device
(FZB_REFERENCE_SUITE_ON_GPU=1).  Not included: test_no_deletions (an orphan module the reference's dispatch never
reaches), test_memmem and the *_cython / C-extension classes (they import the reference's own
extension modules, absent here by design -- they skip themselves on ImportError)."""
import gc
import importlib
import os
import sys
import types
import unittest

import pytest

import conftest

REF = "/root/reference"
MODULES = ["tests.test_common", "tests.test_search_exact", "tests.test_levenshtein", "tests.test_substitutions_only",
           "tests.test_generic_search", "tests.test_find_near_matches", "tests.test_find_near_matches_in_file"]


def _alias_package():
    """`fuzzysearch` IS the fuzzysearch_b200 package object (the reference's dispatch tests patch the search classes
    as attributes of the package, tests/test_find_near_matches.py:40-49), and every submodule is registered under
    both names so that nothing is imported twice."""
    import fuzzysearch_b200
    mods = {"fuzzysearch": fuzzysearch_b200}
    for sub in ("common", "levenshtein", "levenshtein_ngram", "substitutions_only", "generic_search", "search_exact",
                "search", "file_search", "sharding", "_native"):
        mods["fuzzysearch." + sub] = importlib.import_module("fuzzysearch_b200." + sub)
    return mods


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="needs /root/reference/tests")
def test_reference_test_suite_passes_against_this_package(monkeypatch):
    from fuzzysearch_b200 import _native, search
    if not os.environ.get("FZB_REFERENCE_SUITE_ON_GPU"):
        monkeypatch.setattr(_native, "_lib", conftest.load_emulated_library())
        monkeypatch.setenv("FZB_EMU_SMS", "2")
    saved_ws = dict(search._WORKSPACE)
    search._WORKSPACE.clear()
    saved = {k: v for k, v in sys.modules.items() if k == "fuzzysearch" or k.startswith("fuzzysearch.")
             or k == "tests" or k.startswith("tests.")}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(_alias_package())
    sys.path.insert(0, REF)
    try:
        suite = unittest.TestSuite()
        loader = unittest.TestLoader()
        for name in MODULES:
            suite.addTests(loader.loadTestsFromModule(importlib.import_module(name)))
        stream = open(os.devnull, "w")
        result = unittest.TextTestRunner(stream=stream, verbosity=0).run(suite)
        stream.close()
        problems = ["%s\n%s" % (t.id(), tb.splitlines()[-1]) for t, tb in result.failures + result.errors]
        assert not problems, "%d of %d reference tests fail:\n%s" % (len(problems), result.testsRun,
                                                                     "\n".join(problems[:40]))
        assert result.testsRun > 300, result.testsRun
        reasons = {}
        for _, why in result.skipped:
            reasons[why[:60]] = reasons.get(why[:60], 0) + 1
        print("reference suite: %d tests run, %d skipped %r" % (result.testsRun, len(result.skipped), reasons))
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "fuzzysearch" or k.startswith("fuzzysearch.") or k == "tests"
                  or k.startswith("tests.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        search.release_workspace()
        gc.collect()
        search._WORKSPACE.update(saved_ws)
