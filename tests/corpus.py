"""Seeded synthetic corpora with planted near-matches."""
import numpy as np

ASCII = bytes(range(32, 127))
DNA = b"ACGT"


def mutate(rng, pat, alphabet, nedits):
    s = bytearray(pat)
    for _ in range(nedits):
        op = int(rng.integers(3))
        if op == 0 and s:
            s[int(rng.integers(len(s)))] = alphabet[int(rng.integers(len(alphabet)))]
        elif op == 1:
            s.insert(int(rng.integers(len(s) + 1)), alphabet[int(rng.integers(len(alphabet)))])
        elif s:
            del s[int(rng.integers(len(s)))]
    return bytes(s)


def make_corpus(seed, n, alphabet, m, n_plants, max_edits, subs_only=False, clusters=2):
    """-> (pattern bytes, haystack np.uint8[n], plant positions)"""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    hay = alpha[rng.integers(0, len(alpha), size=n)].copy()
    pat = bytes(alpha[rng.integers(0, len(alpha), size=m)])
    positions = []
    if n > 4 * m:
        span = (n - 3 * m) // max(n_plants, 1)
        for i in range(n_plants):
            pos = m + i * span + int(rng.integers(0, max(span - 2 * m, 1)))
            e = int(rng.integers(0, max_edits + 1))
            if subs_only:
                v = bytearray(pat)
                for _ in range(e):
                    v[int(rng.integers(m))] = alphabet[int(rng.integers(len(alphabet)))]
                v = bytes(v)
            else:
                v = mutate(rng, pat, alphabet, e)
            v = v[:max(0, n - pos)]
            hay[pos:pos + len(v)] = np.frombuffer(v, dtype=np.uint8)
            positions.append(pos)
        # adjacent / overlapping clusters to exercise consolidation, and matches at both global ends
        for c in range(clusters):
            pos = m + int(rng.integers(0, n - 4 * m))
            v = pat + pat[m // 2:] + pat
            v = v[:n - pos]
            hay[pos:pos + len(v)] = np.frombuffer(v, dtype=np.uint8)
            positions.append(pos)
        hay[:m - 1] = np.frombuffer(pat[1:], dtype=np.uint8)  # pattern minus first byte at the very start
        hay[n - m + 1:] = np.frombuffer(pat[:m - 1], dtype=np.uint8)  # truncated at the very end
    return pat, hay, positions
