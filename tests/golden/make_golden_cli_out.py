"""Generates tests/golden/cli_out/*.txt: the output of the UNMODIFIED reference program
(oracle/_ref/swipe) for every command line in tests/cli_cases.py, run on the seeded BLAST-file
fixtures that module writes.  Run in the build container:  python tests/golden/make_golden_cli_out.py"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cli_cases  # noqa: E402

SWIPE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "swipe")


def main():
    tmp = tempfile.mkdtemp()
    cli_cases.build(tmp)
    outdir = os.path.join(HERE, "cli_out")
    os.makedirs(outdir, exist_ok=True)
    for name, args in cli_cases.CASES.items():
        r = subprocess.run([SWIPE] + args.split() + ["-a", "1"], cwd=tmp, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("%s failed: %s" % (name, r.stderr))
        open(os.path.join(outdir, name + ".txt"), "w").write(cli_cases.normalise(r.stdout))
        print("%-28s %6d bytes" % (name, len(r.stdout)))


if __name__ == "__main__":
    main()
