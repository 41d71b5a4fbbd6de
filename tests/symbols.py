"""Test-side helpers for wide-symbol sequences (str beyond latin-1, list / tuple): decoding of the
ref_symbols.json fixtures and an INDEPENDENT restatement of the reduction to bytes (pattern symbol -> 1 + rank
among the pattern's distinct symbols in order of first appearance, anything else -> 0) used to drive the byte
oracle.  The product numbers the symbols differently (sorted code points); the results must not depend on it."""
import io

import oracle
from parity import assert_final_parity, tup


def decode_items(codes, kind):
    if kind == "str":
        return "".join(map(chr, codes))
    items = [c if c >= 0 else "s%d" % -c for c in codes]
    return items if kind == "list" else tuple(items)


def reduce_to_bytes(pattern, sequence):
    ids = {}
    for x in pattern:
        ids.setdefault(x, len(ids) + 1)
    assert len(ids) <= 255
    return bytes(ids[x] for x in pattern), bytes(ids.get(x, 0) for x in sequence)


def triples(ms):
    return [(m.start, m.end, m.dist) for m in ms]


def check_fnm(rec, ours, ctx):
    pat_b, hay_b = reduce_to_bytes(rec["pattern"], rec["sequence"])
    a = rec["args"][:4]
    subs, ins, dels, l = oracle.normalize_params(*a)
    if l == 0 or (ins == 0 and dels == 0):
        assert ours == tup(rec["result"]), ctx
    else:
        _, raw = oracle.find_near_matches(pat_b, hay_b, *a, return_raw=True)
        assert_final_parity(ours, rec["result"], raw, ctx)


def check_file(rec, ours, ctx):
    """The text-file loop (__init__.py:174-200) re-run with the byte oracle as the per-chunk search: its raw
    stream (chunk-local window clipping included) explains the reference's final list and ours."""
    from fuzzysearch_b200 import LevenshteinSearchParams, choose_search_class
    pat_b, hay_b = reduce_to_bytes(rec["pattern"], rec["sequence"])
    a, chunk_size = rec["args"][:4], rec["args"][4]
    params = LevenshteinSearchParams(*a)
    cls = choose_search_class(params)
    keep = len(pat_b) - 1 + cls.extra_items_for_chunked_search(pat_b, params)
    raw, f = [], io.BytesIO(hay_b)
    chunk, offset = f.read(chunk_size), 0
    while chunk:
        _, r = oracle.find_near_matches(pat_b, chunk, *a, return_raw=True)
        raw += [(s + offset, e + offset, d) for s, e, d in tup(r)]
        n_keep = min(keep, len(chunk))
        offset += len(chunk) - n_keep
        if n_keep:
            chunk = chunk[-n_keep:] + f.read(chunk_size)
            if len(chunk) == n_keep:
                break
        else:
            chunk = f.read(chunk_size)
    subs, ins, dels, l = params.unpacked
    if l == 0 or (ins == 0 and dels == 0):
        assert ours == tup(rec["result"]) == raw, ctx  # unconsolidated classes: the concatenated chunk lists
    else:
        assert_final_parity(ours, rec["result"], raw, ctx)
