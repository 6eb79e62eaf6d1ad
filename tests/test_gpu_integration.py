"""GPU: the drop-in claim at the reference's own seam.  oracle/_ref/swipe_b200_patched is the
reference program built by integration/build_patched.py with search_chunk's cascade and
align_chunk's search16s call routed through libswipe_b200.so; everything
else in it -- options, database reader, hits_enter, statistics, aligner, report writers -- is the
reference's own code.  Its output must equal what the stock reference printed for the same
database files and command lines (tests/golden/cli_out/)."""
import os
import subprocess

import pytest

import cli_cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "cli_out")
EXE = os.path.join(ROOT, "oracle", "_ref", "swipe_b200_patched")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not os.path.exists(EXE):
        if os.path.isdir("/root/reference"):
            subprocess.run(["python", os.path.join(ROOT, "integration", "build_patched.py")], check=True)
        else:
            pytest.skip("patched reference not built (needs /root/reference; run integration/build_patched.py)")
    d = tmp_path_factory.mktemp("integration")
    cli_cases.build(str(d))
    return str(d)


@pytest.mark.parametrize("name", sorted(n for n in cli_cases.CASES if not n.startswith("dump_")))
def test_patched_reference_prints_the_reference_output(workdir, name):
    r = subprocess.run([EXE] + cli_cases.CASES[name].split(), cwd=workdir, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = cli_cases.normalise(r.stdout)
    exp = open(os.path.join(GOLD, name + ".txt")).read()
    if got != exp:
        g, e = got.splitlines(), exp.splitlines()
        for k in range(max(len(g), len(e))):
            a = g[k] if k < len(g) else "<missing>"
            b = e[k] if k < len(e) else "<missing>"
            assert a == b, "%s line %d:\n  got: %r\n  exp: %r" % (name, k + 1, a, b)


def test_patched_reference_many_threads(workdir):
    """-a 8: eight reference worker threads feed chunks to the one GPU handle (serialised by the
    shim's mutex); the report is the single-thread one."""
    outs = []
    for a in ("1", "8"):
        r = subprocess.run([EXE] + "-d p -i q.fa -m 8 -b 40".split() + ["-a", a], cwd=workdir,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(cli_cases.normalise(r.stdout))
    assert outs[0] == outs[1] == open(os.path.join(GOLD, "protein_tsv.txt")).read()
