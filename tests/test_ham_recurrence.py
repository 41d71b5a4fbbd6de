"""CPU: the two counter layouts of the Hamming counting filter (csrc/ham_recur.h) lose no match -- the C++
harness tests/ham_recur_check.cpp replays k_hamming_count's per-row scheme on the host against brute force."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_counting_filter_recurrences_cover_every_match(tmp_path):
    exe = str(tmp_path / "ham_recur_check")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(HERE, "ham_recur_check.cpp")])
    out = subprocess.check_output([exe]).decode()
    assert out.startswith("ok:"), out
