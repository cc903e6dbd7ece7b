#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference; `make -C oracle ref` first):

    python tests/golden/make_golden.py

Every expected value below is produced by oracle/_ref/igd (the reference CLI) or oracle/_ref/libigdref.so
(the reference objects behind oracle/ref_driver.c) -- never by the oracle restatement or the CUDA path.
The databases are built by the reference's own `igd create`.
"""
from __future__ import annotations

import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from igd_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

V_LIST = [1, 500, 1000]


def run_case(name: str, bed_dir: str, qfile: str, extra_regions=(), full_head: int | None = None):
    """full_head: if set, the -f fixtures use a reduced query file qf.bed (the first full_head lines plus the
    last 13) to keep the committed stdout small."""
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    with tempfile.TemporaryDirectory() as tmp:
        igd = O.ref_create(bed_dir, os.path.join(tmp, "db"), "db")
        stem = igd[:-4]
        exp = {}
        ref = O.RefLib(igd)
        h, r = ref.count_file(qfile)
        exp["hits"] = h.tolist()
        exp["ret"] = int(r)
        for v in V_LIST:
            h, r = ref.count_file(qfile, v)
            exp[f"hits_v{v}"] = h.tolist()
            exp[f"ret_v{v}"] = int(r)
        exp["seqpare_hex"] = [float(x).hex() for x in ref.seqpare_file(qfile)]
        exp["hitsmap"] = ref.hitsmap(0).tolist()
        exp["hitsmap_v500"] = ref.hitsmap(500).tolist()
        exp["regions"] = []
        for chrm, qs, qe, v in extra_regions:
            h, r = ref.count_one(chrm, qs, qe, v)
            exp["regions"].append({"chrm": chrm, "qs": qs, "qe": qe, "v": v, "hits": h.tolist(), "ret": int(r)})
        ref.close()
        qf = qfile
        if full_head is not None:
            lines = open(qfile).read().split("\n")
            qf = os.path.join(tmp, "qf.bed")
            with open(qf, "w") as f:
                f.write("\n".join(lines[:full_head] + lines[-13:]))
        # CLI stdout, byte for byte
        cli = {
            "q": ["-q", qfile],
            "q_v500": ["-q", qfile, "-v", "500"],
            "q_s": ["-q", qfile, "-s"],
            "s_q": ["-s", "-q", qfile],
            "f": ["-q", qf, "-f"],
        }
        # -f ignores -v (src/igd_search.c:975-995): recorded as a fact, not as a second copy
        exp["f_ignores_v"] = O.ref_cli(["search", igd, "-q", qf, "-f"]) == O.ref_cli(
            ["search", igd, "-q", qf, "-f", "-v", "500"])
        for chrm, qs, qe, v in extra_regions[:3]:
            cli[f"r_{chrm}_{qs}_{qe}_{v}"] = ["-r", chrm, str(qs), str(qe)] + (["-v", str(v)] if v else [])
            cli[f"rf_{chrm}_{qs}_{qe}"] = ["-r", chrm, str(qs), str(qe), "-f"]
        exp["cli_args"] = {}
        for key, a in cli.items():
            text = O.ref_cli(["search", igd] + a)
            text = text.replace(qfile, "q.bed").replace(qf, "qf.bed")
            with gzip.open(os.path.join(out, f"stdout_{key}.txt.gz"), "wt") as f:
                f.write(text)
            exp["cli_args"][key] = [{qf: "qf.bed", qfile: "q.bed"}.get(x, x) for x in a]
        with open(igd, "rb") as f, gzip.open(os.path.join(out, "db.igd.gz"), "wb", compresslevel=9) as g:
            g.write(f.read())
        shutil.copy(stem + "_index.tsv", os.path.join(out, "db_index.tsv"))
        shutil.copy(qfile, os.path.join(out, "q.bed"))
        if qf != qfile:
            shutil.copy(qf, os.path.join(out, "qf.bed"))
        with open(os.path.join(out, "expected.json"), "w") as f:
            json.dump(exp, f)
    print(name, "hits", sum(exp["hits"]), "v500", sum(exp["hits_v500"]),
          "size", sum(os.path.getsize(os.path.join(out, x)) for x in os.listdir(out)) // 1024, "KiB")


def run_case_g0(name: str, bed_dir: str, qfile: str, regions, full_head: int):
    """A gType-0 database (`igd create -s 0`: gdata0_t records, src/igd_create.c:245-332).  Only what the reference
    defines on such a database is recorded: count (-q / -r, where -v is not looked at, src/igd_search.c:1024,1042)
    and -f (get_overlaps_f0); its -m and -s read 16-byte records out of the 12-byte file."""
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    with tempfile.TemporaryDirectory() as tmp:
        igd = O.ref_create(bed_dir, os.path.join(tmp, "db"), "db", gtype=0)
        stem = igd[:-4]
        exp = {"gtype": 0}
        ref = O.RefLib(igd)
        assert ref.gtype == 0
        h, r = ref.count_file(qfile)
        exp["hits"], exp["ret"] = h.tolist(), int(r)
        exp["regions"] = []
        for chrm, qs, qe in regions:
            h, r = ref.count_one(chrm, qs, qe)
            exp["regions"].append({"chrm": chrm, "qs": qs, "qe": qe, "v": 0, "hits": h.tolist(), "ret": int(r)})
        ref.close()
        lines = open(qfile).read().split("\n")
        qf = os.path.join(tmp, "qf.bed")
        with open(qf, "w") as f:
            f.write("\n".join(lines[:full_head] + lines[-13:]))
        cli = {"q": ["-q", qfile], "q_v500": ["-q", qfile, "-v", "500"], "f": ["-q", qf, "-f"]}
        for chrm, qs, qe in regions[:2]:
            cli[f"r_{chrm}_{qs}_{qe}_0"] = ["-r", chrm, str(qs), str(qe)]
            cli[f"r_{chrm}_{qs}_{qe}_500"] = ["-r", chrm, str(qs), str(qe), "-v", "500"]
            cli[f"rf_{chrm}_{qs}_{qe}"] = ["-r", chrm, str(qs), str(qe), "-f"]
        exp["cli_args"] = {}
        for key, a in cli.items():
            text = O.ref_cli(["search", igd] + a).replace(qfile, "q.bed").replace(qf, "qf.bed")
            with gzip.open(os.path.join(out, f"stdout_{key}.txt.gz"), "wt") as f:
                f.write(text)
            exp["cli_args"][key] = [{qf: "qf.bed", qfile: "q.bed"}.get(x, x) for x in a]
        with open(igd, "rb") as f, gzip.open(os.path.join(out, "db.igd.gz"), "wb", compresslevel=9) as g:
            g.write(f.read())
        shutil.copy(stem + "_index.tsv", os.path.join(out, "db_index.tsv"))
        shutil.copy(qfile, os.path.join(out, "q.bed"))
        shutil.copy(qf, os.path.join(out, "qf.bed"))
        with open(os.path.join(out, "expected.json"), "w") as f:
            json.dump(exp, f)
    print(name, "hits", sum(exp["hits"]),
          "size", sum(os.path.getsize(os.path.join(out, x)) for x in os.listdir(out)) // 1024, "KiB")


def case_quirk(tmp):
    """the first-bin-empty rule (plain 0, -v 1 = 10, -f 0)."""
    bed = os.path.join(tmp, "quirk_bed")
    os.makedirs(bed)
    for i in range(10):
        with open(os.path.join(bed, f"f{i:05d}.bed"), "w") as f:
            f.write("chr1\t100\t200\t.\t5\nchr1\t40000\t41000\t.\t5\n")
            if i % 2:
                f.write("chr2\t16000\t50000\t.\t700\n")      # spans bins 0..3
    q = os.path.join(tmp, "quirk_q.bed")
    with open(q, "w") as f:
        f.write("chr1\t20000\t45000\n")        # first bin (1) empty -> plain finds nothing
        f.write("chr1\t150\t45000\n")          # first bin 0 non-empty -> bins 1,2 searched
        f.write("chr2\t16383\t16385\n")        # straddles a bin boundary
        f.write("chr2\t16384\t32768\n")        # qe exactly on a boundary: (qe-1)/nbp
        f.write("chr2\t40000\t70000\n")        # runs past the last tile
        f.write("chr2\t70000\t80000\n")        # beyond the last tile
        f.write("chr3\t1\t100\n")              # unknown chrom
    run_case("quirk", bed, q, extra_regions=[("chr1", 20000, 45000, 0), ("chr1", 20000, 45000, 1),
                                             ("chr2", 0, 100000, 0), ("chr9", 5, 50, 0)])


def case_dense(tmp):
    """10 files x 2,000 intervals on 3 short contigs (~64 records per tile), 2,000 queries plus the degenerate
    lines of SURVEY.md Appendix D.13."""
    sizes = np.array([400_000, 250_000, 120_000], dtype=np.int64)
    names = ["chr1", "chr2", "chr3"]
    bed = os.path.join(tmp, "dense_bed")
    os.makedirs(bed)
    for i in range(10):
        iv = synth.gen_intervals(4200 + i, 2000, ("lognormal", 6.5, 1.3), sizes)
        synth.write_bed(os.path.join(bed, f"f{i:05d}.bed"), iv, names)
    q = synth.gen_intervals(4299, 2000, ("lognormal", 7.5, 1.6), sizes)
    qf = os.path.join(tmp, "dense_q.bed")
    synth.write_bed(qf, q, names)
    with open(qf, "a") as f:
        f.write("chrom\tstart\tend\n")                  # header-like line: en = atol("end") = 0 -> rejected
        f.write("chr1\t100\n")                          # two columns
        f.write("chr1\t100\t0\n")                       # end = 0
        f.write("chr1 100 200\n")                       # space separated
        f.write("chr1\t-5\t300\n")                      # negative start (n1 = 0 by truncation)
        f.write("chr1\t16384\t16384\n")                 # qs == qe on a boundary
        f.write("chr1\t20000\t19000\n")                 # qs > qe
        f.write("chr2\t32767\t32769\r\n")               # CRLF
        f.write("chrUn\t5\t500\n")                      # unknown chrom
        f.write("chr3\t119000\t500000\n")               # straddles the last tile
        f.write("chr3\t900000\t900100\n")               # beyond it
        f.write("1\t5\t500\n")                          # no chr prefix
        f.write("chr2\t1000\t2000")                     # last line without newline
    # (-r on tile 0 of contig 0 dereferences the reference's unset tile cache and crashes: preIdx/preChr
    #  start at 0 and only the -q paths reset them, src/igd_search.c:707 -- so chr2 regions start past bin 0)
    run_case("dense", bed, qf, extra_regions=[("chr1", 1000, 60000, 0), ("chr2", 20000, 240000, 500),
                                              ("chr3", 16384, 32768, 0)], full_head=100)


def case_dense0(tmp):
    """The `dense` inputs (other seeds) as a gType-0 database, plus queries that hit the first-bin-empty rule."""
    sizes = np.array([400_000, 250_000, 120_000], dtype=np.int64)
    names = ["chr1", "chr2", "chr3"]
    bed = os.path.join(tmp, "dense0_bed")
    os.makedirs(bed)
    for i in range(10):
        iv = synth.gen_intervals(5200 + i, 1500, ("lognormal", 6.5, 1.3), sizes)
        synth.write_bed(os.path.join(bed, f"f{i:05d}.bed"), iv, names)
    q = synth.gen_intervals(5299, 1500, ("lognormal", 7.5, 1.6), sizes)
    qf = os.path.join(tmp, "dense0_q.bed")
    synth.write_bed(qf, q, names)
    with open(qf, "a") as f:
        f.write("chr1\t100\t0\n")                       # end = 0
        f.write("chr1\t-5\t300\n")                      # negative start
        f.write("chr1\t16384\t16384\n")                 # qs == qe on a boundary
        f.write("chr1\t20000\t19000\n")                 # qs > qe
        f.write("chr2\t32767\t32769\r\n")               # CRLF
        f.write("chrUn\t5\t500\n")                      # unknown chrom
        f.write("chr3\t119000\t500000\n")               # straddles the last tile
        f.write("chr3\t900000\t900100\n")               # beyond it
        f.write("chr2\t1000\t2000")                     # last line without newline
    # the first two regions also go through the CLI; chr1:1000 would crash it (-r on tile 0 of contig 0 reads
    # the unset tile cache, see case_dense) and is taken at library level only
    run_case_g0("dense0", bed, qf, regions=[("chr2", 20000, 240000), ("chr3", 16384, 32768),
                                            ("chr1", 1000, 60000)], full_head=60)


def case_cfg1(tmp):
    """BASELINE config 1: 10 files x 10k lognormal intervals on hg38, 10k queries."""
    law = ("lognormal", 6.5, 1.2)
    bed = os.path.join(tmp, "cfg1_bed")
    synth.write_bed_files(bed, 1, 10, 10000, law)
    q = synth.gen_intervals(synth.seed_for(1, synth.QUERY_FILE_INDEX), 10000, law)
    qf = os.path.join(tmp, "cfg1_q.bed")
    synth.write_bed(qf, q)
    run_case("cfg1", bed, qf, extra_regions=[("chr1", 1_000_000, 3_000_000, 0)])


if __name__ == "__main__":
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    only = sys.argv[1:]                     # e.g. `make_golden.py dense0` regenerates one case
    with tempfile.TemporaryDirectory(prefix="g") as tmp:
        for case in (case_quirk, case_dense, case_dense0, case_cfg1):
            if not only or case.__name__[5:] in only:
                case(tmp)
