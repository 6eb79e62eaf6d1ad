#!/usr/bin/env python
"""Builds the reference program with its search path routed through libswipe_b200.so -- the drop-in
claim of INTEGRATION.md section 2, compiled.  TEST INFRASTRUCTURE: needs /root/reference.

The reference sources are copied to a scratch directory (never into this repository) and swipe.cc
receives exactly these edits, each anchored on one line of the original:

  1. after  `db_mapsequences(sdp->dbt, s1, s2);`  in search_chunk (swipe.cc:1401):
         swb_gpu_search_chunk(sdp); return;          -- the cascade below it becomes dead code
  2. the `search16s(` call of align_chunk (swipe.cc:381) is renamed to a macro that expands to
         swb_gpu_search16s(sdp, qstrand, qframe)
  3. `#include "swb_search_chunk.inc"` (integration/, ours) is appended, with forward declarations
     inserted before search_data's first use.

The binary is written to oracle/_ref/swipe_b200_patched (git-ignored like the other reference builds;
it travels to the GPU box with the snapshot).  tests/test_gpu_integration.py diffs its output with
the stock reference program's on the CLI fixtures."""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SWB_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref", "swipe_b200_patched")
CSRC = os.path.join(ROOT, "swipe_b200", "csrc")


def patch(text):
    anchor = "db_mapsequences(sdp->dbt, s1, s2);"
    assert text.count(anchor) == 1
    text = text.replace(anchor, anchor + "\n  swb_gpu_search_chunk(sdp); return;   /* swipe_b200 */\n")
    call = "search16s((WORD**)qtable,"
    assert text.count(call) == 1
    text = text.replace(call, "SWB_GPU_SEARCH16S((WORD**)qtable,")
    decl = "void fatal(const char * message)\n"
    assert text.count(decl) == 1
    text = text.replace(decl, "void swb_gpu_search_chunk(struct search_data * sdp);            /* swipe_b200 */\n"
                              "void swb_gpu_search16s(struct search_data * sdp, long, long);   /* swipe_b200 */\n"
                              "#define SWB_GPU_SEARCH16S(...) swb_gpu_search16s(sdp, qstrand, qframe)\n\n" + decl, 1)
    return text + '\n#include "swb_search_chunk.inc"   /* swipe_b200 */\n'


def build():
    if not os.path.isdir(REF):
        print("integration: %s absent - keeping the prebuilt %s" % (REF, OUT))
        return OUT if os.path.exists(OUT) else None
    from swipe_b200 import build as b
    lib = b.build_lib()
    tmp = tempfile.mkdtemp(prefix="swb_integration_")
    try:
        for f in glob.glob(os.path.join(REF, "*.cc")) + glob.glob(os.path.join(REF, "*.h")) + \
                glob.glob(os.path.join(REF, "*.c")):
            shutil.copy(f, tmp)
        shutil.copy(os.path.join(ROOT, "integration", "swb_search_chunk.inc"), tmp)
        src = os.path.join(tmp, "swipe.cc")
        patched = patch(open(src).read())
        os.chmod(src, 0o644)
        open(src, "w").write(patched)
        objs = ["database", "asnparse", "align", "matrices", "stats", "hits", "query", "search63", "search16",
                "search16s", "search7", "swipe"]
        flags = ["-O3", "-g", "-w", "-I" + os.path.join(ROOT, "include")]
        for o in objs:
            subprocess.run(["g++"] + flags + ["-c", "-o", o + ".o", o + ".cc"], cwd=tmp, check=True)
        subprocess.run(["g++"] + flags + ["-mssse3", "-DSWIPE_SSSE3", "-c", "-o", "search7_ssse3.o", "search7.cc"],
                       cwd=tmp, check=True)
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-o", OUT] + [o + ".o" for o in objs] + ["search7_ssse3.o", "-L" + CSRC,
                       "-lswipe_b200", "-Wl,-rpath," + CSRC, "-Wl,-rpath,$ORIGIN/../../swipe_b200/csrc", "-lpthread"],
                       cwd=tmp, check=True)
        assert lib
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print(build())
