#!/usr/bin/env python
"""`igd create` of this repo against the unmodified reference's, same inputs, same host:

    python profiles/bench_create.py [n_files] [n_per_file]

Writes BED files of BASELINE config 2's shape (hg38, lognormal(6.0, 1.0) lengths), runs both binaries (wall clock,
cold output directory, warm page cache for the inputs), checks that the two .igd files are the same bytes (contig
name padding aside, see tests/test_create.py) and prints one JSON line."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from igd_b200 import api, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.test_create import canonical_igd  # noqa: E402


def main():
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    n_per = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    with tempfile.TemporaryDirectory(prefix="cb", dir=os.environ.get("TMPDIR", "/tmp")) as work:
        bed = os.path.join(work, "bed")
        synth.write_bed_files(bed, 2, n_files, n_per, ("lognormal", 6.0, 1.0))
        res = {"n_files": n_files, "intervals": n_files * n_per, "cores": os.cpu_count()}
        igd = {}
        for tag, binary in (("reference", O.REF_BIN), ("igd_b200", api.CLI_PATH)):
            if not os.path.exists(binary):
                continue
            out = os.path.join(work, tag) + "/"
            t0 = time.perf_counter()
            r = subprocess.run([binary, "create", bed, out, "db"], capture_output=True)
            res[tag + "_s"] = round(time.perf_counter() - t0, 3)
            assert r.returncode == 0, r.stderr.decode()
            igd[tag] = out + "db.igd"
        res["bytes"] = os.path.getsize(igd["igd_b200"])
        if len(igd) == 2:
            res["identical"] = canonical_igd(igd["reference"]) == canonical_igd(igd["igd_b200"])
            res["speedup"] = round(res["reference_s"] / res["igd_b200_s"], 2)
        print(json.dumps(res))


if __name__ == "__main__":
    main()
