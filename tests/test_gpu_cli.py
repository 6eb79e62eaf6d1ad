"""GPU: the command-line front end (swipe_b200/csrc/swipe-b200: the reference's options, hit list,
statistics and report formats over the CUDA scan) against what the unmodified reference program
printed for the same database files and command lines (tests/golden/cli_out/, made by
tests/golden/make_golden_cli_out.py).  Bar: identical text, apart from the banner / wall-clock
lines listed in cli_cases.DROP."""
import os
import subprocess

import pytest

import cli_cases
from swipe_b200 import build

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "cli_out")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    cli_cases.build(str(d))
    return str(d)


@pytest.mark.parametrize("name", sorted(cli_cases.CASES))
def test_cli_matches_reference_output(workdir, name):
    exe = build.build_cli()
    r = subprocess.run([exe] + cli_cases.CASES[name].split(), cwd=workdir, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    got = cli_cases.normalise(r.stdout)
    exp = open(os.path.join(GOLD, name + ".txt")).read()
    if got != exp:
        g, e = got.splitlines(), exp.splitlines()
        for k in range(max(len(g), len(e))):
            a = g[k] if k < len(g) else "<missing>"
            b = e[k] if k < len(e) else "<missing>"
            assert a == b, "%s line %d:\n  got: %r\n  exp: %r" % (name, k + 1, a, b)


def test_cli_two_gpus_same_output_as_one(workdir):
    """-a N shards the database over N GPUs (as many as the box has); the merged hit list and the
    report must not change."""
    exe = build.build_cli()
    outs = []
    for n in ("1", "2", "8"):
        r = subprocess.run([exe] + "-d p -i q.fa -m 8 -b 40".split() + ["-a", n], cwd=workdir,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] == outs[1] == outs[2]
    assert cli_cases.normalise(outs[0]) == open(os.path.join(GOLD, "protein_tsv.txt")).read()


def test_cli_stdin_query_and_output_file(workdir):
    """-i - reads the query from stdin (the default, swipe.cc:34) and -o writes the report to a file."""
    exe = build.build_cli()
    q = open(os.path.join(workdir, "q.fa")).read()
    r = subprocess.run([exe] + "-d p -m 8 -b 40 -o out.tsv".split(), cwd=workdir, input=q, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and r.stdout == "", r.stderr
    got = cli_cases.normalise(open(os.path.join(workdir, "out.tsv")).read())
    assert got == open(os.path.join(GOLD, "protein_tsv.txt")).read()
