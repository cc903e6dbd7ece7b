"""The CLI's -f formatter (igd_b200/csrc/igd_fmt.h) against printf with the reference's format strings
(src/igd_search.c:545,565,592): the same bytes for edge values, random values, long names and a buffer small enough to
flush every few lines.  Host C only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fmt_is_printf(tmp_path):
    exe = str(tmp_path / "fmt_check")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-fsanitize=address,undefined", "-I",
                    os.path.join(ROOT, "igd_b200", "csrc"), os.path.join(ROOT, "tests", "csrc", "fmt_check.c"), "-o", exe],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
