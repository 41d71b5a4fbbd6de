#!/usr/bin/env python
"""Golden fixtures for WIDE-symbol sequences (general-Unicode str, list / tuple of hashable items):
outputs of the REAL reference (pure-Python configuration) on seeded inputs.

Run in the build container only (``/root/reference`` is absent on the GPU box):

    PYTHONHASHSEED=0 python tests/golden/gen_golden_symbols.py

-> ``ref_symbols.json``.  Symbols are stored as integers: code points for str, the items themselves for
list / tuple (ints; negative numbers stand for the strings 's<n>' so that non-int hashables are covered too --
see ``decode_items``).  Each record: {"fn", "kind", "pattern", "sequence", "args", "result"}.
"""
import io
import json
import os
import random
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

sys.path.insert(0, os.path.join(REF, "src"))

import fuzzysearch  # noqa: E402
from fuzzysearch import levenshtein_ngram  # noqa: E402
from fuzzysearch.search_exact import search_exact  # noqa: E402

assert levenshtein_ngram._expand_short is levenshtein_ngram._py_expand_short, "natives present?"

ALPHABETS = {
    "greek": [chr(c) for c in range(0x3B1, 0x3B1 + 24)],
    "dna_cjk": list("中文字符"),                      # 4 symbols: repetitive like DNA
    "emoji": [chr(c) for c in range(0x1F600, 0x1F600 + 12)],         # outside the BMP (UTF-32 path)
    "mixed": list("abcdefgh") + [chr(0x3B1), chr(0x416), chr(0x4E2D), chr(0x1F600), chr(0xFF), chr(0x100)],
    "latin_plus_one": list("ACGT") + [chr(0x2028)],
    "surrogate": list("xy") + [chr(0xD800), chr(0xDFFF), chr(0xFFFF)],  # lone surrogates are symbols too
}


def decode_items(codes, kind):
    """ints -> the sequence the reference saw (shared with the tests)."""
    if kind == "str":
        return "".join(map(chr, codes))
    items = [c if c >= 0 else "s%d" % -c for c in codes]
    return items if kind == "list" else tuple(items)


def mutate(rng, pat, alphabet, nedits):
    s = list(pat)
    for _ in range(nedits):
        op = rng.randrange(3)
        pos = rng.randrange(len(s) + 1) if s else 0
        if op == 0 and s:
            s[min(pos, len(s) - 1)] = rng.choice(alphabet)
        elif op == 1:
            s.insert(pos, rng.choice(alphabet))
        elif s:
            del s[min(pos, len(s) - 1)]
    return s


def make_case(rng, alphabet, seq_alphabet, mmin=1, mmax=24, nmax=300):
    m = rng.randint(mmin, mmax)
    pat = [rng.choice(alphabet) for _ in range(m)]
    n = rng.randint(0, nmax)
    hay = [rng.choice(seq_alphabet) for _ in range(n)]
    for _ in range(rng.randint(0, 4)):
        ins = mutate(rng, pat, seq_alphabet, rng.randint(0, 4))
        pos = rng.randint(0, len(hay))
        if rng.random() < 0.5:
            hay[pos:pos + len(ins)] = ins
        else:
            hay[pos:pos] = ins
    return pat, hay


def triples(ms):
    return [[m.start, m.end, m.dist] for m in ms]


def main():
    rng = random.Random(20260924)
    recs = []

    def draw_params():
        r = rng.random()
        if r < 0.45:
            return [None, None, None, rng.randint(0, 4)]
        if r < 0.65:
            return [rng.randint(0, 4), 0, 0, None]
        return [rng.randint(0, 3), rng.randint(0, 3), rng.randint(0, 3), rng.choice([None, rng.randint(0, 4)])]

    # ---- str -------------------------------------------------------------------------------------------
    names = sorted(ALPHABETS)
    for i in range(900):
        name = names[i % len(names)]
        alphabet = ALPHABETS[name]
        seq_alphabet = alphabet
        r = rng.random()
        if r < 0.15:    # the sequence holds symbols the pattern never has (and the other way round)
            seq_alphabet = alphabet[: max(2, len(alphabet) // 2)] + [chr(0x2603), "z"]
        elif r < 0.25:  # latin-1 pattern over a wide sequence
            alphabet = [c for c in alphabet if ord(c) < 256] or ["a", "b"]
            seq_alphabet = alphabet + [chr(0x20AC)]
        elif r < 0.35:  # wide pattern over a latin-1 sequence
            seq_alphabet = [c for c in alphabet if ord(c) < 256] or ["a", "b"]
            alphabet = seq_alphabet + [chr(0x20AC)]
        pat, hay = make_case(rng, alphabet, seq_alphabet)
        pat_s, hay_s = "".join(pat), "".join(hay)
        args = draw_params()
        try:
            res = triples(fuzzysearch.find_near_matches(pat_s, hay_s, *args))
        except ValueError:
            continue
        recs.append({"fn": "find_near_matches", "kind": "str", "pattern": [ord(c) for c in pat_s],
                     "sequence": [ord(c) for c in hay_s], "args": args, "result": res})
        if i % 6 == 0:  # search_exact with and without a window
            sub = pat_s[: rng.randint(1, min(4, len(pat_s)))]
            a, b = rng.randint(-3, len(hay_s) + 3), rng.choice([None, rng.randint(-3, len(hay_s) + 3)])
            recs.append({"fn": "search_exact", "kind": "str", "pattern": [ord(c) for c in sub],
                         "sequence": [ord(c) for c in hay_s], "args": [a, b],
                         "result": list(search_exact(sub, hay_s, a, b))})
        if i % 9 == 0 and len(hay_s) > 40:  # the text-file loop (__init__.py:174-200), small chunks
            chunk = rng.choice([32, 64, 100])
            if chunk > len(pat_s) + 8:
                try:
                    res = triples(fuzzysearch.find_near_matches_in_file(pat_s, io.StringIO(hay_s), *args,
                                                                        _chunk_size=chunk))
                except ValueError:
                    continue
                recs.append({"fn": "find_near_matches_in_file", "kind": "str", "pattern": [ord(c) for c in pat_s],
                             "sequence": [ord(c) for c in hay_s], "args": args + [chunk], "result": res})

    # ---- list / tuple ------------------------------------------------------------------------------------
    for i in range(300):
        base = rng.choice([[1, 2, 3, 4], list(range(10, 40)), [5, 70000, -1, -2, 2 ** 40], [-1, -2, -3]])
        seq_base = base + ([999, -9] if rng.random() < 0.3 else [])
        pat, hay = make_case(rng, base, seq_base, mmax=16, nmax=200)
        kind = "list" if i % 2 == 0 else "tuple"
        args = draw_params()
        try:
            res = triples(fuzzysearch.find_near_matches(decode_items(pat, kind), decode_items(hay, kind), *args))
        except ValueError:
            continue
        recs.append({"fn": "find_near_matches", "kind": kind, "pattern": pat, "sequence": hay, "args": args,
                     "result": res})

    out = {"source": "seeded cases through /root/reference (fuzzysearch 0.8.1, pure-Python configuration, "
                     "PYTHONHASHSEED=0): str with symbols outside latin-1, list / tuple sequences",
           "records": recs}
    with open(os.path.join(HERE, "ref_symbols.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    from collections import Counter
    print(len(recs), Counter((r["fn"], r["kind"]) for r in recs))


if __name__ == "__main__":
    main()
