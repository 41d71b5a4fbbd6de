"""Build tests/emu/_build/libfuzzb200_emu.so: the PRODUCT sources compiled for the CPU emulator.  Test infrastructure.

The sources under fuzzysearch_b200/csrc are copied into tests/emu/_build/src with THREE textual rewrites -- g++ cannot
parse them otherwise -- and compiled with -DFZB_EMU against tests/emu/include (cuda_runtime.h / cuda.h stand-ins):

  kernel<<<grid, block, smem, stream>>>(args);   ->  emu::launch(dim3(grid), dim3(block), smem, [..]{ kernel(args); });
  extern __shared__ __align__(N) T name[];       ->  T *name = emu::dyn_smem<T>();
  __shared__ T a[N], b;                          ->  static T a[N], b; static emu::SharedReg ...(&a, sizeof(a)), ...;

(the argument expressions are evaluated once, into a tuple, before the CTAs run).  Everything else -- the kernels, the
host logic of api.cu, the C-ABI -- is the code that ships; the handful of inline-PTX helpers carry an `#ifdef FZB_EMU`
C++ twin in the product sources.
"""
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "fuzzysearch_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "libfuzzb200_emu.so")


def _match_paren(s, i):
    """s[i] == '(' -> index of the matching ')'"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += c
    parts.append(cur.strip())
    return parts


_KERNEL_RE = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*$")


def rewrite_launches(src):
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return out + src[pos:]
        m = _KERNEL_RE.search(src[:i])
        assert m, "no kernel name before <<< at offset %d" % i
        j = src.index(">>>", i)
        cfg = _split_top(src[i + 3:j])
        assert 2 <= len(cfg) <= 4, cfg
        a = src.index("(", j)
        assert src[j + 3:a].strip() == "", src[j:a + 1]
        b = _match_paren(src, a)
        args = src[a + 1:b]
        smem = cfg[2] if len(cfg) > 2 else "0"
        stream = cfg[3] if len(cfg) > 3 else "nullptr"
        call = ("do { auto emu_args_ = std::make_tuple(%s); emu::launch(dim3(%s), dim3(%s), (size_t)(%s), %s, "
                "[emu_args_]() { std::apply([](const auto &...emu_a_) { %s(emu_a_...); }, emu_args_); }); } while (0)"
                % (args, cfg[0], cfg[1], smem, stream, m.group(1)))
        out += src[pos:m.start(1)] + call
        pos = b + 1


_DYN_RE = re.compile(r"extern\s+__shared__\s+__align__\(\d+\)\s+(\w+)\s+(\w+)\[\];")
_SHARED_RE = re.compile(r"^(\s*)__shared__\s+((?:__align__\(\d+\)\s+)?)([^;]+);", re.M)


def _rewrite_shared(m):
    """`__shared__ T a[N], b;` -> a function-local static plus its registration with the emulator (which saves and
    restores the __shared__ variables of a CTA that has to wait for another one)."""
    indent, align, body = m.group(1), m.group(2), m.group(3)
    decls = _split_top(body)
    first = re.match(r"(.*?)(\w+)\s*((?:\[[^\]]*\])*)$", decls[0].strip())
    assert first, body
    names = [first.group(2)] + [re.match(r"(\w+)", d.strip()).group(1) for d in decls[1:]]
    regs = ", ".join("emu_reg_%s(&%s, sizeof(%s))" % (n, n, n) for n in names)
    return "%sstatic %s%s; static emu::SharedReg %s;" % (indent, align, body, regs)


def rewrite(src):
    src = _DYN_RE.sub(lambda m: "%s *%s = emu::dyn_smem<%s>();" % (m.group(1), m.group(2), m.group(1)), src)
    src = _SHARED_RE.sub(_rewrite_shared, src)
    assert not re.search(r"^\s*(extern\s+)?__shared__", src, re.M), "unhandled __shared__ declaration"
    return rewrite_launches(src)


def _flavour():
    for name in ("COVERAGE", "ASAN", "UBSAN", "TSAN"):
        if os.environ.get("FZB_EMU_" + name):
            return name.lower()
    return ""


def build(force=False, verbose=False):
    flavour = _flavour()  # the instrumented builds live next to the plain one
    out = OUT if not flavour else OUT.replace(".so", "-%s.so" % flavour)
    srcdir = os.path.join(BUILD, "src")
    os.makedirs(srcdir, exist_ok=True)
    digest = hashlib.sha256()
    names = sorted(n for n in os.listdir(CSRC) if n.endswith((".cu", ".cuh", ".h")))
    texts = {}
    for n in names:
        t = rewrite(open(os.path.join(CSRC, n)).read())
        texts[n] = t
        digest.update(n.encode() + b"\0" + t.encode())
    digest.update((os.environ.get("FZB_EMU_COVERAGE", "") + "/" + os.environ.get("FZB_EMU_ASAN", "") + "/" + os.environ.get("FZB_EMU_UBSAN", "") + "/" + os.environ.get("FZB_EMU_TSAN", "")).encode())
    for extra in (os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "include", "cuda.h"),
                  os.path.join(ROOT, "include", "fuzzb200.h"), os.path.abspath(__file__)):
        digest.update(open(extra, "rb").read())
    stamp = os.path.join(BUILD, "stamp" + ("-" + flavour if flavour else ""))
    if (not force and os.path.exists(out) and os.path.exists(stamp)
            and open(stamp).read() == digest.hexdigest()):
        return out
    for n, t in texts.items():
        with open(os.path.join(srcdir, n), "w") as f:
            f.write(t)
    cxx = os.environ.get("CXX", "g++")
    opt = ["-O1"]
    if os.environ.get("FZB_EMU_COVERAGE"):  # gcov line coverage of the product sources under the test-suite
        opt = ["-O0", "--coverage"]
    if os.environ.get("FZB_EMU_ASAN"):  # AddressSanitizer: out-of-bounds accesses of kernels and host code
        opt = ["-O1", "-fsanitize=address"]  # (run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so))
    if os.environ.get("FZB_EMU_TSAN"):  # ThreadSanitizer as a racecheck of the kernels (cuda_runtime.h explains the edges)
        opt = ["-O1", "-fsanitize=thread"]
    if os.environ.get("FZB_EMU_UBSAN"):  # shifts past the width, signed overflow, misaligned accesses, ...
        opt = ["-O1", "-fsanitize=undefined", "-fno-sanitize=vptr,pointer-overflow", "-fno-sanitize-recover=undefined"]  # (W = H - buf_lo: virtual base pointers wrap by design)
    cmd = [cxx] + opt + ["-g", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DFZB_EMU",
           "-fno-omit-frame-pointer", "-Wno-unknown-pragmas", "-Wno-attributes",
           "-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "include"), "-I", srcdir,
           "-x", "c++", os.path.join(srcdir, "api.cu"), "-o", out, "-lpthread", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(digest.hexdigest())
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
