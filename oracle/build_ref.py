#!/usr/bin/env python
"""Compile the REFERENCE itself into oracle/_ref/ (git-ignored, travels with gpurun).

Only runs where /root/reference exists (the build container).  Nothing is copied: every source file
is compiled from where it lies under /root/reference and only binaries land in oracle/_ref/:

* the two C extensions the reference ships (setup.py:110-116): fuzzysearch._common
  (_common.c + memmem.c) and fuzzysearch._substitutions_only (+ memmem.c), with gcc;
* the package's pure-Python modules, compiled to extension modules with Cython so the package is
  importable on the GPU box without sources (same statements, same semantics: configuration "D",
  the build `pip install fuzzysearch` produces, is what bench.py --impl reference times).

Usage: python oracle/build_ref.py
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

REF = "/root/reference"
SRC = os.path.join(REF, "src", "fuzzysearch")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "fuzzysearch")
EXT = sysconfig.get_config_var("EXT_SUFFIX")
INC = sysconfig.get_paths()["include"]

PY_MODULES = ["__init__", "common", "search_exact", "levenshtein", "levenshtein_ngram",
              "substitutions_only", "generic_search", "no_deletions"]
C_EXTS = {"_common": ["_common.c", "memmem.c"],
          "_substitutions_only": ["_substitutions_only.c", "memmem.c"]}


def main():
    if not os.path.isdir(SRC):
        print("reference not present; nothing to do")
        return 0
    stamp = os.path.join(OUT, ".built")
    if os.path.exists(stamp):
        return 0
    shutil.rmtree(os.path.join(HERE, "_ref"), ignore_errors=True)
    os.makedirs(OUT)
    cflags = ["-O2", "-fPIC", "-shared", "-I" + INC, "-I" + REF, "-I" + SRC, "-w"]
    for name, files in C_EXTS.items():
        subprocess.check_call(["gcc"] + cflags + [os.path.join(SRC, f) for f in files] +
                              ["-o", os.path.join(OUT, name + EXT)])
    with tempfile.TemporaryDirectory() as tmp:
        for mod in PY_MODULES:
            c_file = os.path.join(tmp, mod + ".c")
            subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name",
                                   "fuzzysearch" if mod == "__init__" else "fuzzysearch." + mod,
                                   os.path.join(SRC, mod + ".py"), "-o", c_file])
            subprocess.check_call(["gcc"] + cflags + [c_file, "-o", os.path.join(OUT, mod + EXT)])
    open(stamp, "w").write("built from %s\n" % REF)
    print("reference compiled into", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
