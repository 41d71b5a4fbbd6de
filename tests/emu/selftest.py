"""Self-test of the emulator's ThreadSanitizer mode: a kernel with a missing __syncthreads() MUST be reported, its
correct twins must not; a race between two CTAs is seen in grid mode (FZB_EMU_TSAN_GRID=1) only.  Test infrastructure (run by tests/test_emu_kernels.py when libtsan is present)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402


def run():
    out_dir = os.path.join(build_emu.BUILD, "selftest")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(out_dir, "racy.cu")
    with open(src, "w") as f:
        f.write(build_emu.rewrite(open(os.path.join(HERE, "selftest", "racy.cu")).read()))
    exe = os.path.join(out_dir, "racy")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")
    subprocess.check_call([cxx, "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-DFZB_EMU", "-Wno-unknown-pragmas",
                           "-Wno-attributes", "-I", os.path.join(HERE, "include"), "-x", "c++", src, "-o", exe,
                           "-lpthread"])
    results = {}
    for which, grid in ((0, "0"), (1, "0"), (2, "0"), (3, "0"), (4, "0"), (4, "1"), (0, "1")):
        p = subprocess.run([exe, str(which)], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0", FZB_EMU_TSAN_GRID=grid))
        assert "done %d" % which in p.stdout, (which, p.stdout, p.stderr[-2000:])
        results[(which, grid == "1")] = p.stderr.count("WARNING: ThreadSanitizer: data race")
    return results


if __name__ == "__main__":
    r = run()
    print(r)
    ok = (r[(0, False)] == 0 and r[(1, False)] > 0 and r[(2, False)] == 0 and r[(3, False)] > 0
          and r[(4, False)] == 0 and r[(4, True)] > 0 and r[(0, True)] == 0)
    print("selftest", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
