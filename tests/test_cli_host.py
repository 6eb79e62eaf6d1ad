"""Host-side checks of the command-line front end that need no GPU: option errors end with the
reference's messages and exit status 1 (fatal(), swipe.cc:158-170), and without a device the program
fails loudly instead of computing anything on the CPU."""
import subprocess

import pytest

import blastdb
import fixtures
from swipe_b200 import build, synth


def run(args, cwd=None):
    return subprocess.run([build.build_cli()] + args, capture_output=True, text=True, cwd=cwd, timeout=120)


@pytest.mark.parametrize("args,msg", [
    ([], "No database specified."),
    (["-d", "x", "-m", "5"], "Illegal view type."),
    (["-d", "x", "-p", "9", "-G", "5", "-E", "1"], "Illegal symbol type."),
    (["-d", "x", "-S", "2"], "Illegal strand specified for protein query."),
    (["-d", "x", "-M", "nosuchmatrix"], "Unknown score matrix. Gap penalties must be specified (-G and -E)."),
    (["-d", "x", "-Q", "7"], "Illegal query genetic code specified."),
    (["-d", "x", "-a", "0"], "Illegal number of threads specified"),
    (["-d", "x", "-C", "T"], "Composition-based score adjustments not supported."),
    (["-d", "/nonexistent/db"], "Unable to open file /nonexistent/db.pin."),
])
def test_option_errors(args, msg):
    r = run(args)
    assert r.returncode == 1
    assert msg in r.stderr


def test_help_lists_the_reference_options():
    r = run(["-h"])
    assert r.returncode == 1
    for opt in ("--db=FILE", "--matrix=NAME/FILE", "--gapopen=NUM", "--num_alignments=NUM", "--evalue=REAL",
                "--outfmt=NUM", "--symtype=NAME/NUM", "--strand=NAME/NUM", "--db_gencode=NUM", "--dbsize=NUM"):
        assert opt in r.stdout


def test_no_cpu_path(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    q = synth.protein_query(50)
    blastdb.write_protein(str(tmp_path / "p"), fixtures.blast_protein_subjects(q)[:5])
    blastdb.write_fasta(str(tmp_path / "q.fa"), q)
    r = run(["-d", "p", "-i", "q.fa"], cwd=str(tmp_path))
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr and r.stdout == ""


def test_database_dump_matches_reference(tmp_path):
    """-N 1 / -N 2 (db_show_fasta, database.cc:1483-1537) need no GPU: reader, defline parser and
    nucleotide unpacking against the reference program's output for the same files."""
    import os
    import cli_cases
    cli_cases.build(str(tmp_path))
    gold = os.path.join(os.path.dirname(__file__), "golden", "cli_out")
    for name in ("dump_protein", "dump_protein_split", "dump_nt"):
        r = run(cli_cases.CASES[name].split(), cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr
        assert cli_cases.normalise(r.stdout) == open(os.path.join(gold, name + ".txt")).read(), name
