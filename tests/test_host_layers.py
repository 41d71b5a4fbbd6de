"""CPU: the Python layers above the C-ABI (dispatch, limits, types, wide symbols, ``matched`` slices, file chunk
loops, batches, locks) with the ctypes layer replaced by the oracle-backed TEST DOUBLE of tests/fake_backend.py.
The expected values are outputs of the REAL reference (tests/golden/*.json).  The kernels themselves are out
of reach here -- the ``-m gpu`` files replay the same fixtures through the real library."""
import io
import threading

import numpy as np
import pytest

import oracle
from fake_backend import fake_device  # noqa: F401  (fixture)
from fuzzysearch_b200 import (DeviceSequence, Match, find_near_matches, find_near_matches_batch,
                              find_near_matches_in_file, has_near_match, search_exact)
from parity import assert_final_parity, load_golden, tup
from symbols import check_file, check_fnm, decode_items, reduce_to_bytes, triples


def test_public_api_reproduces_the_reference_on_byte_inputs(fake_device):  # noqa: F811
    n = 0
    for name in ("ref_suite_calls.json", "ref_fuzz.json"):
        for rec in load_golden(name):
            a = rec["args"]
            if rec["fn"] == "search_exact" and "exc" not in rec and len(a[0]) > 0:
                assert search_exact(bytes.fromhex(a[0]), bytes.fromhex(a[1]), a[2], a[3]) == list(rec["result"])
                n += 1
            if rec["fn"] != "find_near_matches":
                continue
            pat, hay = bytes.fromhex(a[0]), bytes.fromhex(a[1])
            if "exc" in rec:
                with pytest.raises((ValueError, TypeError)):
                    find_near_matches(pat, hay, *a[2:6])
                continue
            if len(pat) > 255:
                continue
            ms = find_near_matches(pat, hay, *a[2:6])
            subs, ins, dels, l = oracle.normalize_params(*a[2:6])
            if l == 0 or (ins == 0 and dels == 0):
                assert triples(ms) == tup(rec["result"]), a
            else:
                _, raw = oracle.find_near_matches(pat, hay, *a[2:6], return_raw=True)
                assert_final_parity(triples(ms), rec["result"], raw, str(a))
            assert all(m.matched == hay[m.start:m.end] for m in ms)
            n += 1
    assert n > 2000


def test_public_api_reproduces_the_reference_on_wide_symbols(fake_device):  # noqa: F811
    n = 0
    for rec in load_golden("ref_symbols.json"):
        kind = rec["kind"]
        pat, seq = decode_items(rec["pattern"], kind), decode_items(rec["sequence"], kind)
        ctx = "%s %s %r" % (rec["fn"], kind, rec["args"])
        if rec["fn"] == "search_exact":
            assert search_exact(pat, seq, *rec["args"]) == list(rec["result"]), ctx
        elif rec["fn"] == "find_near_matches_in_file":
            ms = find_near_matches_in_file(pat, io.StringIO(seq), *rec["args"][:4], _chunk_size=rec["args"][4])
            check_file(rec, triples(ms), ctx)
        else:
            ms = find_near_matches(pat, seq, *rec["args"])
            check_fnm(rec, triples(ms), ctx)
            assert all(m.matched == seq[m.start:m.end] and type(m.matched) is type(seq[0:0]) for m in ms), ctx
        n += 1
    assert n > 1400


def test_resident_sequences_batches_and_type_rules(fake_device):  # noqa: F811
    text = "αβγδ" * 50 + "needle-in-a-haystack" + "абвг" * 40 + "needle-in-a-haystack"[:9] + "zz" * 30
    ds = DeviceSequence(text)
    assert len(ds) == len(text)
    for pat, k in (("needle-in-a-haystack", 2), ("δαβγ", 0), ("гаЖв", 1), ("needle", 1)):
        got = find_near_matches(pat, ds, max_l_dist=k)
        assert triples(got) == oracle.find_near_matches(*reduce_to_bytes(pat, text), max_l_dist=k)
        assert all(m.matched == text[m.start:m.end] for m in got)
        assert has_near_match(pat, ds, max_l_dist=k) == bool(got)
        assert search_exact(pat[:3], ds, 5, len(text) - 5) == \
            oracle.search_exact(*reduce_to_bytes(pat[:3], text), 5, len(text) - 5)
    pats = ["needle-in-a-haystack", "абвгабвг", "γδαβγδ"]
    assert [triples(b) for b in find_near_matches_batch(pats, ds, [2, 1, 0])] == \
        [triples(find_near_matches(p, text, max_l_dist=k)) for p, k in zip(pats, [2, 1, 0])]
    # k == 0 entries of a batch: the exact-occurrence list, like find_near_matches
    assert triples(find_near_matches_batch([b"aa"], b"aaaa", 0)[0]) == [(0, 2, 0), (1, 3, 0), (2, 4, 0)]
    # latin-1 str is searched as bytes; a latin-1 resident sequence takes a wide pattern too
    lat = DeviceSequence("caf\xe9 au lait, caf\xe9 noir")
    assert triples(find_near_matches("caf\xe9", lat, max_l_dist=0)) == [(0, 4, 0), (14, 18, 0)]
    assert triples(find_near_matches("cafē", lat, max_l_dist=1)) == [(0, 4, 1), (14, 18, 1)]
    assert find_near_matches("caf\xe9", lat, max_l_dist=0)[0].matched == "caf\xe9"
    # bytes-likes of every kind, zero-copy
    for seq in (b"---PATERN---", bytearray(b"---PATERN---"), memoryview(b"---PATERN---"),
                np.frombuffer(b"---PATERN---", dtype=np.uint8)):
        ms = find_near_matches(b"PATTERN", seq, max_l_dist=1)
        assert triples(ms) == [(3, 9, 1)] and bytes(ms[0].matched) == b"PATERN"
    # kinds do not mix; unhashable items are refused before any device work
    for pat, seq in ((b"a", "abc"), ("a", b"abc"), (["a"], "abc"), ("a", ["a"])):
        with pytest.raises(TypeError):
            find_near_matches(pat, seq, max_l_dist=1)
    with pytest.raises(TypeError):
        find_near_matches([[1]], [[1], [2]], max_l_dist=1)
    # consolidate_matches of the search classes returns the device-side list without recomputing it
    from fuzzysearch_b200 import LevenshteinSearch, LevenshteinSearchParams
    raw = LevenshteinSearch.search(b"PATTERN", b"--PATTERN--PATERN--", LevenshteinSearchParams(None, None, None, 1))
    assert LevenshteinSearch.consolidate_matches(raw) == [Match(2, 9, 0, b"PATTERN"), Match(11, 17, 1, b"PATERN")]
    assert len(list(raw)) >= 2 and all(isinstance(m, Match) for m in raw)


def test_binary_and_text_file_loops(fake_device, tmp_path):  # noqa: F811
    needle = b"PATTERNXYZ12"
    for chunk_size in (100, 4096):
        for variant in (needle, b"PATERNXYZ12", b"PATTERNxYZ12"):
            for delta in (-len(needle), -len(needle) + 1, -4, -2, -1, 0, 1):  # test_find_near_matches_in_file.py:84-86
                hay = bytearray(chunk_size + 100)
                hay[chunk_size + delta:chunk_size + delta + len(variant)] = variant
                path = tmp_path / "hay.bin"
                path.write_bytes(bytes(hay))
                with open(path, "rb") as f:
                    got = triples(find_near_matches_in_file(needle, f, max_l_dist=1, _chunk_size=chunk_size))
                assert got == triples(find_near_matches(needle, bytes(hay), max_l_dist=1)) and len(got) == 1
                got2 = find_near_matches_in_file(needle.decode(), io.StringIO(bytes(hay).decode("latin-1")),
                                                 max_l_dist=1, _chunk_size=chunk_size)
                assert triples(got2) == got and isinstance(got2[0].matched, str)
    with pytest.raises(ValueError):
        find_near_matches_in_file(b"", io.BytesIO(b"abc"), max_l_dist=1)


def test_threads_share_workspace_and_resident_sequences(fake_device):  # noqa: F811
    rng = np.random.default_rng(3)
    jobs = []
    for t in range(4):
        hay = bytes(rng.integers(97, 101, size=20000, dtype=np.uint8))
        pat = hay[500 * (t + 1):500 * (t + 1) + 12]
        jobs.append((pat, hay, oracle.find_near_matches(pat, hay, max_l_dist=1)))
    wide = "".join(chr(0x3B1 + int(c)) for c in rng.integers(0, 6, size=5000))
    ds = DeviceSequence(wide)
    wpats = [wide[100 * (t + 1):100 * (t + 1) + 10].replace(chr(0x3B1 + t), chr(0x400 + t)) for t in range(4)]
    wexp = [oracle.find_near_matches(*reduce_to_bytes(w, wide), max_l_dist=2) for w in wpats]
    errors = []

    def run(t):
        try:
            for _ in range(5):
                pat, hay, exp = jobs[t]
                assert triples(find_near_matches(pat, hay, max_l_dist=1)) == exp
                assert triples(find_near_matches(wpats[t], ds, max_l_dist=2)) == wexp[t]
        except BaseException as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


class FakeSeq(object):
    """Stand-in for Bio.Seq.Seq (BioPython is not in this image): text with Seq's slicing / str / find protocol."""

    def __init__(self, data):
        self._data = str(data)

    def __str__(self):
        return self._data

    def __len__(self):
        return len(self._data)

    def __getitem__(self, i):
        return FakeSeq(self._data[i]) if isinstance(i, slice) else self._data[i]

    def __eq__(self, other):
        return str(self) == str(other)

    __hash__ = None

    def find(self, sub, start=0, end=None):
        return self._data.find(str(sub), start, len(self._data) if end is None else end)


def test_bio_seq_objects_are_searched_as_text(fake_device, monkeypatch):  # noqa: F811
    """The reference treats Bio.Seq.Seq as a text class (search_exact.py:14-19): same matches as for the str,
    `matched` a slice of the ORIGINAL Seq.  Sequence, subsequence, resident sequence and search_exact."""
    from fuzzysearch_b200 import DeviceSequence, find_near_matches, has_near_match, search
    monkeypatch.setattr(search, "_BioSeq", FakeSeq)
    text = "GATTACA" * 3 + "TGCACTGTAGGGATAACAAT" + "ACGT" * 5 + "TGCACTGTAGGATAACAAT" + "CC"
    pat = "TGCACTGTAGGGATAACAAT"
    expected = find_near_matches(pat, text, max_l_dist=2)
    assert [(m.start, m.end, m.dist) for m in expected] == oracle.find_near_matches(
        pat.encode(), text.encode(), max_l_dist=2)
    for p, t in ((pat, FakeSeq(text)), (FakeSeq(pat), FakeSeq(text)), (FakeSeq(pat), text)):
        got = find_near_matches(p, t, max_l_dist=2)
        assert got == expected
        for m in got:
            assert type(m.matched) is type(t) and str(m.matched) == text[m.start:m.end]
        assert has_near_match(p, t, max_l_dist=2) is True
        assert find_near_matches(p, t, max_substitutions=1, max_insertions=0, max_deletions=0) == \
            find_near_matches(pat, text, max_substitutions=1, max_insertions=0, max_deletions=0)
    resident = DeviceSequence(FakeSeq(text))
    assert len(resident) == len(text)
    got = find_near_matches(FakeSeq(pat), resident, max_l_dist=2)
    assert got == expected and all(isinstance(m.matched, FakeSeq) for m in got)
    assert search.search_exact("ACGT", FakeSeq(text), 3, 60) == search.search_exact("ACGT", text, 3, 60)
    assert search.search_exact(FakeSeq("ACGT"), resident) == search.search_exact("ACGT", text)
    resident.close()
    with pytest.raises(TypeError):
        find_near_matches(FakeSeq(pat), text.encode(), max_l_dist=2)  # text against bytes, as for a str


def test_module_level_route_functions_on_the_fake_backend(fake_device):  # noqa: F811
    """fuzzysearch_b200.levenshtein / .levenshtein_ngram / .substitutions_only / .generic_search mirror the
    reference's module-level functions (names, arguments, errors); here their Python side, the kernels behind them in
    tests/test_gpu_golden.py (and on the emulator)."""
    from fuzzysearch_b200 import LevenshteinSearchParams
    from fuzzysearch_b200.generic_search import find_near_matches_generic, has_near_match_generic_ngrams
    from fuzzysearch_b200.levenshtein import find_near_matches_levenshtein
    from fuzzysearch_b200.levenshtein_ngram import find_near_matches_levenshtein_ngrams
    from fuzzysearch_b200.search_exact import search_exact as se
    from fuzzysearch_b200.substitutions_only import (find_near_matches_substitutions,
                                                     find_near_matches_substitutions_ngrams,
                                                     has_near_match_substitutions)
    from test_gpu_golden import replay_route_functions
    assert replay_route_functions(load_golden("ref_suite_calls.json")) > 300
    t = lambda ms: [(m.start, m.end, m.dist) for m in ms]  # noqa: E731
    # SURVEY 8c known answers
    assert t(find_near_matches_levenshtein_ngrams(b"PATTERN", b"----------PATT-ERN---------", 2)) == [(10, 18, 1)] * 3
    assert sorted(t(find_near_matches_levenshtein(b"PATTERN", b"----------PATT-ERN---------", 2))) == \
        [(10, 18, 1), (10, 18, 2), (11, 18, 2)]  # 7 // 3 < 3: the router takes the LP route
    assert t(find_near_matches_levenshtein_ngrams(b"def", b"abcddefg", 1)) == [(3, 7, 1), (4, 7, 0), (4, 7, 0), (4, 7, 0)]
    assert t(find_near_matches_levenshtein("def", "abcddefg", 0)) == [(4, 7, 0)]
    assert t(find_near_matches_substitutions(b"abc", b"xbz", 5)) == [(0, 3, 2)]
    assert has_near_match_substitutions(b"abc", b"xbz", 1) is False
    assert se(b"ab", b"abab") == [0, 2]
    params = LevenshteinSearchParams(1, 1, 1, 2)
    assert t(find_near_matches_generic(b"PATTERN", b"---PATERN---", params)) == \
        tup(oracle.generic_raw(b"PATTERN", b"---PATERN---", 1, 1, 1, 2))
    assert has_near_match_generic_ngrams(b"PATTERN", b"---PATERN---", LevenshteinSearchParams(1, 1, 1, 1)) is True
    for bad in (lambda: find_near_matches_levenshtein(b"", b"abc", 1), lambda: find_near_matches_levenshtein(b"a", b"abc", -1),
                lambda: find_near_matches_levenshtein_ngrams(b"ab", b"abc", 2),
                lambda: find_near_matches_substitutions(b"", b"abc", 1),
                lambda: find_near_matches_substitutions(b"ab", b"abc", -1),
                lambda: find_near_matches_substitutions_ngrams(b"ab", b"abc", 2),
                lambda: find_near_matches_generic(b"", b"abc", params)):
        with pytest.raises(ValueError):
            bad()
