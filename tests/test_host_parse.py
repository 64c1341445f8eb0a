"""Host text pipeline of the CLI (gemma_b200/host/line_pipeline.h): the fast decimal reader must return exactly what atof
returns (the reference parses genotypes with atof, src/gemma_io.cpp:741), and the threaded line pipeline must preserve
file order.  CPU only."""
import gzip
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_decimal_reader_equals_atof_and_pipeline_keeps_order(tmp_path):
    exe = str(tmp_path / "parse_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "host", "parse_check.cpp"), "-lz"])
    lines = ["%d %s" % (i, "x" * ((i * 7919) % 5000)) for i in range(20000)]
    plain = tmp_path / "lines.txt"
    plain.write_text("\n".join(lines))                      # no trailing newline on the last line
    gz = tmp_path / "lines.txt.gz"
    with gzip.open(gz, "wt") as f:
        f.write("\r\n".join(lines) + "\r\n")                # CRLF endings
    for path in (plain, gz):
        r = subprocess.run([exe, str(path)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "mismatches 0" in r.stdout and "lines 20000 out_of_order 0" in r.stdout, r.stdout
