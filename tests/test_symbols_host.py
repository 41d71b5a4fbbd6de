"""CPU: the premise behind the wide-symbol path -- every algorithm on the path only compares a pattern
symbol with a sequence symbol, so renaming the pattern's symbols to 1..d and everything else to 0 changes no
result -- checked against outputs of the REAL reference on general-Unicode str and list / tuple inputs
(tests/golden/ref_symbols.json), through the byte oracle; plus the host half of the product's reduction."""
import numpy as np
import pytest

import oracle
from parity import assert_final_parity, load_golden, tup
from symbols import decode_items, reduce_to_bytes


def test_reduction_is_result_neutral_against_the_reference():
    n = 0
    for rec in load_golden("ref_symbols.json"):
        if rec["fn"] == "find_near_matches_in_file":
            continue
        pat, hay = reduce_to_bytes(rec["pattern"], rec["sequence"])
        ctx = "%s %s %r" % (rec["fn"], rec["kind"], rec["args"])
        if rec["fn"] == "search_exact":
            assert oracle.search_exact(pat, hay, *rec["args"]) == list(rec["result"]), ctx
        else:
            final, raw = oracle.find_near_matches(pat, hay, *rec["args"], return_raw=True)
            subs, ins, dels, l = oracle.normalize_params(*rec["args"])
            if l == 0 or (ins == 0 and dels == 0):
                assert final == tup(rec["result"]), ctx
            else:
                assert_final_parity(final, rec["result"], raw, ctx)
        n += 1
    assert n > 1000


def test_host_side_of_the_reduction():
    from fuzzysearch_b200 import search as S
    from fuzzysearch_b200._native import UnsupportedError
    # str: sorted code points; the renamed pattern is its ranks
    alpha = S._make_alphabet(["βaα\U0001F600a"], "str")
    assert alpha == [ord("a"), 0x3B1, 0x3B2, 0x1F600]
    assert S._rename("βaα\U0001F600a", "str", alpha).tolist() == [3, 1, 2, 4, 1]
    # several patterns share one alphabet (batches)
    alpha = S._make_alphabet(["ab", "bЖ"], "str")
    assert [S._rename(p, "str", alpha).tolist() for p in ("ab", "bЖ")] == [[1, 2], [2, 3]]
    # items: numbered in order of first appearance, any hashable
    ids = S._make_alphabet([[5, "x", (1, 2), 5]], "items")
    assert S._rename([5, "x", (1, 2), 5], "items", ids).tolist() == [1, 2, 3, 1]
    with pytest.raises(UnsupportedError):
        S._make_alphabet(["".join(chr(0x400 + i) for i in range(256))], "str")
    with pytest.raises(UnsupportedError):
        S._make_alphabet([list(range(256))], "items")
    # code units: UCS-2 while every character is in the BMP (lone surrogates included), UTF-32 otherwise
    u = S._code_units("a中\ud800￿")
    assert u.dtype == np.uint16 and u.tolist() == [0x61, 0x4E2D, 0xD800, 0xFFFF]
    u = S._code_units("a\U0001F600\udfff")
    assert u.dtype == np.uint32 and u.tolist() == [0x61, 0x1F600, 0xDFFF]
    assert S._code_units("").size == 0
    # kinds
    assert [S._kind(x) for x in ("s", b"b", bytearray(b"b"), [1], (1,), memoryview(b"x"))] == \
        ["str", "bytes", "bytes", "items", "items", "bytes"]
    assert S._narrow("\xe9t\xe9", "str").tolist() == [0xE9, 0x74, 0xE9] and S._narrow("Ā", "str") is None
