"""Test-side helpers for wide-symbol sequences (str beyond latin-1, list / tuple): decoding of the
ref_symbols.json fixtures and an INDEPENDENT restatement of the reduction to bytes (pattern symbol -> 1 + rank
among the pattern's distinct symbols in order of first appearance, anything else -> 0) used to drive the byte
oracle.  The product numbers the symbols differently (sorted code points); the results must not depend on it."""


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
