"""Device-level test of the expansion routines (fzb_debug_expand): every recorded call of the reference's
_expand / _py_expand_short / _py_expand_long (its own TestExpandBase vectors, tests/test_levenshtein.py:64-158,
harvested into tests/golden/ref_suite_calls.json, plus the seeded fuzz records) is evaluated by the four
device code paths the verify kernels use -- bit-parallel and cell-by-cell, forwards (right expansion) and
backwards (left expansion) -- and must equal the reference's answer."""
import numpy as np
import pytest

import oracle
from fuzzysearch_b200 import _native as F
from parity import load_golden

pytestmark = pytest.mark.gpu

VARIANT = {"expand": 0, "expand_short": 1, "expand_long": 2}


def _check(cases, expected, ctx):
    out = F.debug_expand(cases)
    for (sub, seq, k, var), exp, row in zip(cases, expected, out.tolist()):
        want = [-1, -1] if exp[0] is None else [int(exp[0]), int(exp[1])]
        paths = {"bp_fwd": row[0:2], "bp_rev": row[2:4], "dp_fwd": row[4:6], "dp_rev": row[6:8]}
        for name, got in paths.items():
            if got == [-2, -2]:
                assert name.startswith("bp") and len(sub) > 64
                continue
            assert got == want, "%s %s: sub=%r seq=%r k=%d variant=%d -> %r, reference %r" % (
                ctx, name, sub, seq, k, var, got, want)


def test_expand_golden_records(cuda_device):
    n = 0
    for name in ("ref_suite_calls.json", "ref_fuzz.json"):
        cases, expected = [], []
        for rec in load_golden(name):
            if rec["fn"] not in VARIANT or "exc" in rec:
                continue
            a = rec["args"]
            cases.append((bytes.fromhex(a[0]), bytes.fromhex(a[1]), int(a[2]), VARIANT[rec["fn"]]))
            expected.append(rec["result"])
        _check(cases, expected, name)
        n += len(cases)
    assert n > 3000


def test_expand_quirk_vectors(cuda_device):
    # _py_expand_short's early break misses a later equal-score expansion
    cases = [(b"TACA", b"TAACAGA", 3, 1), (b"TACA", b"TAACAGA", 3, 2),
             (b"AACAAACAAA", b"AACAAAACCCAAA", 3, 1), (b"AACAAACAAA", b"AACAAAACCCAAA", 3, 2),
             (b"", b"abc", 2, 0), (b"abc", b"", 3, 0), (b"abc", b"", 2, 0)]
    expected = [(1, 3), (1, 5), (3, 9), (3, 13), (0, 0), (3, 0), (None, None)]
    _check(cases, expected, "quirk")


def test_expand_fuzz_vs_oracle(cuda_device):
    rng = np.random.default_rng(11)
    cases, expected = [], []
    for _ in range(20000):
        a = int(rng.integers(2, 6))
        sl = int(rng.integers(0, 70))
        k = int(rng.integers(0, 9))
        sub = bytes(rng.integers(65, 65 + a, size=sl, dtype=np.uint8))
        if rng.random() < 0.6 and sl:
            s = bytearray(sub)
            for _ in range(int(rng.integers(0, k + 2))):
                op = int(rng.integers(3))
                if op == 0 and s:
                    s[int(rng.integers(len(s)))] = 65 + int(rng.integers(a))
                elif op == 1:
                    s.insert(int(rng.integers(len(s) + 1)), 65 + int(rng.integers(a)))
                elif s:
                    del s[int(rng.integers(len(s)))]
            seq = (bytes(s) + bytes(rng.integers(65, 65 + a, size=int(rng.integers(0, 6)), dtype=np.uint8)))[:sl + k]
        else:
            seq = bytes(rng.integers(65, 65 + a, size=int(rng.integers(0, sl + k + 3)), dtype=np.uint8))
        var = int(rng.integers(0, 3))
        cases.append((sub, seq, k, var))
        expected.append(oracle.expand(sub, seq, k, ("auto", "short", "long")[var]))
    _check(cases, expected, "fuzz")
