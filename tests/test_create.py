"""`igd create` of this repo (host C, igd_b200/csrc/igd_create.c) against the UNMODIFIED reference's `igd create`
(oracle/_ref/igd, compiled in place from /root/reference): same .igd bytes -- header, tile tables and every
record in the same order, ties of the reference's unstable radix sort included; the 40-byte contig names are
compared as strings because the reference writes heap bytes behind the terminator (src/igd_base.c:422) -- same
_index.tsv, same stdout.  The committed fixture tests/golden/create holds the reference's outputs for a box
without /root/reference."""
import gzip
import hashlib
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from igd_b200 import api, synth
from oracle import oracle as O
from tests.conftest import SMALL_SIZES

NAMES = ["chr1", "chr2", "chr3"]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "create")


def canonical_igd(path: str) -> bytes:
    """the .igd file with the bytes behind each contig name's terminator zeroed"""
    raw = bytearray(open(path, "rb").read())
    nbp, gtype, m = np.frombuffer(raw, np.int32, 3)
    ntile = np.frombuffer(raw, np.int32, m, 12)
    off = 12 + 4 * m + 4 * int(ntile.sum())
    for i in range(m):
        nm = raw[off + 40 * i: off + 40 * (i + 1)]
        k = nm.find(b"\0")
        if k >= 0:
            raw[off + 40 * i + k: off + 40 * (i + 1)] = b"\0" * (40 - k)
    return bytes(raw)


def run_create(binary, src, out, name, *opts):
    r = subprocess.run([binary, "create", src, out, name, *opts], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    return r.stdout.decode(errors="replace")


def write_inputs(bed, seed=0, n_files=12, n=3000, gz_every=4):
    """BED files with many equal starts (ties), long records spanning tiles, a dense tile of several hundred
    records (the radix path, > 64 per tile), 3- to 6-column lines and some awkward ones."""
    os.makedirs(bed)
    rng = np.random.default_rng(77 + seed)
    for i in range(n_files):
        iv = synth.gen_intervals(9100 + 31 * seed + i, n, ("lognormal", 6.5, 1.4), SMALL_SIZES)
        lines = []
        for c, s, e, v in zip(iv.ctg, iv.start, iv.end, iv.value):
            s = int(s) // 64 * 64                      # many ties on start
            e = max(int(e), s + 1)
            lines.append(f"{NAMES[c]}\t{s}\t{e}\t.\t{v}\n")
        for k in range(400):                            # one hot tile: ties inside a > 64-record run
            s = 70000 + int(rng.integers(0, 40)) * 16
            lines.append(f"chr2\t{s}\t{s + int(rng.integers(1, 60000))}\tx\t{int(rng.integers(0, 1000))}\t+\n")
        lines.insert(5, "chr1\t500\t400\t.\t3\n")      # start >= end: counted in the tsv, not stored
        lines.insert(9, "chr1\t100\t200\n")            # three columns: no value
        lines.insert(10, "chr3\t10\t20\t.\t7\textra\tcols\there\n")
        lines.append("chrNew%d\t5\t50000\t.\t1\n" % (i % 3))       # contigs appearing late
        if i % 5 == 2:      # the splitter's field limit crosses files: this one ends on 3-column lines ...
            lines += ["chr1\t%d\t%d\n" % (900 + k, 1900 + k) for k in range(3)]
        if i % 5 == 3:      # ... so the next one's first 6-column line is cut at 4 fields and loses its value
            lines.insert(0, "chr2\t4000\t4800\tname\t777\t-\n")
            lines.insert(1, "chr2\t4100\t4900\tname\t778\t-\n")
        p = os.path.join(bed, f"f{i:05d}.bed")
        if gz_every and i % gz_every == 1:
            with gzip.open(p + ".gz", "wt") as f:
                f.writelines(lines)
        else:
            with open(p, "w") as f:
                f.writelines(lines)


CASES = {
    "default": [],
    "b12": ["-b", "12"],
    "gtype0": ["-s", "0"],
    "b17": ["-b", "17"],
}


def outputs(binary, work, bed, tag, opts):
    out = os.path.join(work, tag) + "/"
    stdout = run_create(binary, bed, out, "db", *opts)
    return {"igd": canonical_igd(out + "db.igd"), "tsv": open(out + "db_index.tsv").read(),
            "stdout": stdout.replace(out, "<OUT>/").replace(work, "<W>"), "listing": sorted(os.listdir(out))}


@pytest.fixture(scope="module")
def inputs(workdir):
    bed = os.path.join(workdir, "create_bed")
    if not os.path.isdir(bed):
        write_inputs(bed)
    return bed


@pytest.mark.parametrize("case", list(CASES))
def test_create_matches_committed_reference_output(case, inputs, workdir):
    """works without /root/reference: digests of the reference's outputs are committed (make_fixture below)"""
    exp = json.load(open(os.path.join(GOLD, "expected.json")))[case]
    got = outputs(api.CLI_PATH, workdir, inputs, "mine_gold_" + case, CASES[case])
    assert hashlib.sha256(got["igd"]).hexdigest() == exp["igd_sha256"]
    assert got["tsv"] == exp["tsv"]
    assert got["stdout"] == exp["stdout"]
    assert got["listing"] == exp["listing"]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", list(CASES))
def test_create_matches_reference_live(case, inputs, workdir):
    ref = outputs(O.REF_BIN, workdir, inputs, "ref_" + case, CASES[case])
    got = outputs(api.CLI_PATH, workdir, inputs, "mine_" + case, CASES[case])
    assert got["igd"] == ref["igd"]
    assert got["tsv"] == ref["tsv"]
    assert got["stdout"] == ref["stdout"]
    assert got["listing"] == ref["listing"]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_create_list_mode_and_bed4(inputs, workdir):
    """-f (a file listing the sources; parse_bed's rule per line) and -s 2 (one file, 4th column = dataset)"""
    lst = os.path.join(workdir, "sources.txt")
    files = sorted(os.path.join(inputs, f) for f in os.listdir(inputs))
    with open(lst, "w") as f:
        f.write("\n".join(files[:11] + ["/nonexistent/file.bed"]) + "\n")
    outs = []
    for binary, tag in ((O.REF_BIN, "ref_list"), (api.CLI_PATH, "mine_list")):
        out = os.path.join(workdir, tag) + "/"
        so = run_create(binary, lst, out, "db", "-f")
        outs.append((canonical_igd(out + "db.igd"), open(out + "db_index.tsv").read(),
                     so.replace(out, "<OUT>/").replace(workdir, "<W>")))
    a, b = outs
    # -f passes an UNINITIALISED value to igd_add in the reference (src/igd_create.c:139,189): compare all but it
    ra, rb = (np.frombuffer(x[0][-(len(x[0]) - hdr_len(x[0])):], np.int32).reshape(-1, 4) for x in (a, b))
    assert a[0][:hdr_len(a[0])] == b[0][:hdr_len(b[0])]
    assert np.array_equal(ra[:, :3], rb[:, :3]) and not rb[:, 3].any()
    assert a[1] == b[1] and a[2] == b[2]

    bed4 = os.path.join(workdir, "jaspar.bed")
    rng = np.random.default_rng(5)
    with open(bed4, "w") as f:
        for k in range(20000):
            s = int(rng.integers(0, 900000))
            f.write(f"chr{1 + k % 3}\t{s}\t{s + int(rng.integers(1, 3000))}\tTF{int(rng.integers(0, 37))}\t{k % 1000}\n")
    outs = []
    for binary, tag in ((O.REF_BIN, "ref_bed4"), (api.CLI_PATH, "mine_bed4")):
        out = os.path.join(workdir, tag) + "/"
        so = run_create(binary, bed4, out, "db", "-s", "2")
        outs.append((canonical_igd(out + "db.igd"), open(out + "db_index.tsv").read(),
                     so.replace(out, "<OUT>/").replace(workdir, "<W>")))
    assert outs[0] == outs[1]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_create_dense_tiles_radix_levels(workdir):
    """thousands of records per tile with heavy ties: the in-place byte-radix levels of the reference's sort, whose
    order of equal starts is an artefact of its displacement cycles, reproduced record for record"""
    bed = os.path.join(workdir, "create_bed_dense")
    write_inputs(bed, seed=1, n_files=10, n=40000, gz_every=0)
    ref = outputs(O.REF_BIN, workdir, bed, "ref_dense", [])
    got = outputs(api.CLI_PATH, workdir, bed, "mine_dense", [])
    assert got == ref
    nrec = (len(got["igd"]) - hdr_len(got["igd"])) // 16
    m = int(np.frombuffer(got["igd"], np.int32, 1, 8)[0])
    ntiles = int(np.frombuffer(got["igd"], np.int32, m, 12).sum())
    assert nrec / ntiles > 1000


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_create_hg38_threads(workdir):
    """hg38-sized tile table (188 k tiles): the multi-threaded fill and sort, against the reference and against a
    single-threaded run of itself"""
    bed = os.path.join(workdir, "create_bed_hg38")
    synth.write_bed_files(bed, 2, 20, 20000, ("lognormal", 6.0, 1.0))
    ref = outputs(O.REF_BIN, workdir, bed, "ref_hg38", [])
    got = outputs(api.CLI_PATH, workdir, bed, "mine_hg38", [])
    assert got == ref
    out1 = os.path.join(workdir, "mine_hg38_1t") + "/"
    r = subprocess.run([api.CLI_PATH, "create", bed, out1, "db"], capture_output=True,
                       env=dict(os.environ, IGD_B200_CREATE_THREADS="1"))
    assert r.returncode == 0
    assert canonical_igd(out1 + "db.igd") == got["igd"]


def test_create_under_asan_ubsan(inputs, workdir):
    """the host C of `igd create` built with -fsanitize=address,undefined (sanitizer builds of the
    host C) over every flavour: no report, and the same bytes as the product binary"""
    import shutil
    src = os.path.join(os.path.dirname(api.LIB_PATH), "..", "csrc", "igd_create.c")
    inc = os.path.dirname(api.HEADER_PATH)
    main_c = os.path.join(workdir, "asan_main.c")
    with open(main_c, "w") as f:
        f.write("int igd_create(int argc, char **argv);\nint main(int c, char **v) { return igd_create(c, v); }\n")
    exe = os.path.join(workdir, "igd_create_asan")
    cc = shutil.which("gcc") or "gcc"
    r = subprocess.run([cc, "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", inc, "-o", exe,
                        main_c, src, "-lz", "-lpthread", "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer runtime here: " + r.stderr[-200:])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:halt_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    for case, opts in CASES.items():
        out = os.path.join(workdir, "asan_" + case) + "/"
        r = subprocess.run([exe, "create", inputs, out, "db", *opts], capture_output=True, env=env)
        assert r.returncode == 0 and b"ERROR" not in r.stderr and b"runtime error" not in r.stderr, r.stderr.decode()[-2000:]
        ref = os.path.join(workdir, "mine_gold_" + case, "db.igd")
        if os.path.exists(ref):
            assert canonical_igd(out + "db.igd") == canonical_igd(ref)
    lst = os.path.join(workdir, "asan_sources.txt")
    with open(lst, "w") as f:
        f.write("\n".join(sorted(os.path.join(inputs, x) for x in os.listdir(inputs))) + "\n/nonexistent\n")
    r = subprocess.run([exe, "create", lst, os.path.join(workdir, "asan_list"), "db", "-f"], capture_output=True, env=env)
    assert r.returncode == 0 and b"ERROR" not in r.stderr and b"runtime error" not in r.stderr, r.stderr.decode()[-2000:]
    bed4 = os.path.join(workdir, "asan_bed4.bed")
    with open(bed4, "w") as f:
        for k in range(5000):
            f.write(f"chr{1 + k % 3}\t{k * 37}\t{k * 37 + 500 + k % 900}\tTF{k % 23}\t{k % 1000}\n")
        f.write("chr1\t5\n\nchr2\n")
    r = subprocess.run([exe, "create", bed4, os.path.join(workdir, "asan_b4"), "db", "-s", "2"], capture_output=True, env=env)
    assert r.returncode == 0 and b"ERROR" not in r.stderr and b"runtime error" not in r.stderr, r.stderr.decode()[-2000:]


def _awkward_files(bed, with_overflow=False):
    os.makedirs(bed)
    rng = np.random.default_rng(3)

    def base(n=300):
        L = []
        for k in range(n):
            a = int(rng.integers(0, 900000))
            L.append(f"chr{1 + k % 3}\t{a}\t{a + int(rng.integers(1, 5000))}\tn{k}\t{int(rng.integers(0, 1000))}\n")
        return L

    cases = {
        1: lambda L: L[:5] + ["chr1\t100\t200\t" + "x" * 3000 + "\t5\n"] + L[5:],      # longer than the 1024-byte line buffer
        2: lambda L: L[:7] + ["chr2\t500\t400\t.\t9\n", "chr2\t500\t500\t.\t9\n"] + L[7:],   # start >= end
        3: lambda L: [x.replace("\n", "\r\n") for x in L],                              # CRLF
        4: lambda L: L[:3] + ["chr1\t-50\t300\t.\t4\n"] + L[3:],                        # negative start
        5: lambda L: L[:4] + ["chr1 10 20 . 5\n"] + L[4:],                                # blanks instead of tabs
        6: lambda L: L[:2] + ["chr3\t1000\t2000\n"] * 3 + L[2:],                          # three columns in the middle
        7: lambda L: L[:9] + ["chr1\t 77\t 99 \t.\t 12\n"] + L[9:],                      # blanks around the numbers
        9: lambda L: L + ["chr2\t10\t20\t.\t3"],                                        # no newline at the end
        10: lambda L: L[:8] + ["chrVeryLongContigName_0123456789_0123456789_0123456789\t5\t50\t.\t1\n"] + L[8:],
        11: lambda L: L[:1] + ["chr1\t100\t200\t.\t1e3\n", "chr1\t0x10\t0x20\t.\t7\n", "chr1\t+5\t+50\t.\t-3\n"] + L[1:],
    }
    if with_overflow:       # beyond int32: the reference indexes a negative tile and dies
        cases[8] = lambda L: L[:6] + ["chr1\t3000000000\t3000000100\t.\t1\n"] + L[6:]
    for i in range(12):
        L = base()
        if i in cases:
            L = cases[i](L)
        with open(os.path.join(bed, f"f{i:05d}.bed"), "w", newline="") as f:
            f.writelines(L)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_create_awkward_lines_like_the_reference(workdir):
    """over-long lines (split by the 1024-byte gzgets buffer), CRLF, start >= end, a negative start, blanks for
    tabs, 3-column lines in the middle of 5-column ones, a 54-character contig name, odd numerals, a last line
    without newline: same bytes, tsv and stdout as the unmodified reference"""
    bed = os.path.join(workdir, "create_bed_awkward")
    _awkward_files(bed)
    ref = outputs(O.REF_BIN, workdir, bed, "ref_awkward", [])
    got = outputs(api.CLI_PATH, workdir, bed, "mine_awkward", [])
    assert got == ref
    # (the reference's -s 0 flavour segfaults on these files; this one must at least get through them)
    got0 = outputs(api.CLI_PATH, workdir, bed, "mine_awkward0", ["-s", "0"])
    assert len(got0["igd"]) > 12


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_create_odd_input_directories_like_the_reference(workdir):
    """an empty file (its tsv line reads `-nan`), a sub-directory among the inputs, a multi-member gzip file and
    a header line: same bytes, tsv and stdout as the unmodified reference."""
    bed = os.path.join(workdir, "create_bed_odd")
    _awkward_files(bed)
    for i in (1, 2, 3, 4, 5, 6, 7, 9, 10, 11):          # plain files again, the oddities are about the directory
        rng = np.random.default_rng(40 + i)
        with open(os.path.join(bed, f"f{i:05d}.bed"), "w") as f:
            for k in range(200):
                a = int(rng.integers(0, 900000))
                f.write(f"chr{1 + k % 3}\t{a}\t{a + int(rng.integers(1, 5000))}\tn{k}\t{int(rng.integers(0, 1000))}\n")
    open(os.path.join(bed, "f00003.bed"), "w").close()                       # empty
    os.makedirs(os.path.join(bed, "f00005.dir"))                              # a directory in the glob
    lines = open(os.path.join(bed, "f00006.bed")).readlines()
    os.remove(os.path.join(bed, "f00006.bed"))
    with open(os.path.join(bed, "f00006.bed.gz"), "wb") as f:                 # two gzip members in one file
        f.write(gzip.compress("".join(lines[:100]).encode()))
        f.write(gzip.compress("".join(lines[100:]).encode()))
    first = open(os.path.join(bed, "f00000.bed")).readlines()
    with open(os.path.join(bed, "f00000.bed"), "w") as f:
        f.writelines(["chrom\tstart\tend\tname\tscore\n"] + first)           # header line: atol("start") = 0 = atol("end")
    ref = outputs(O.REF_BIN, workdir, bed, "ref_odd", [])
    got = outputs(api.CLI_PATH, workdir, bed, "mine_odd", [])
    assert got == ref
    assert "nan" in got["tsv"]


def test_create_gets_through_blank_and_comment_lines(workdir):
    """two lines in a row with fewer fields than the code reads make the reference dereference the NULL that
    str_splits just planted and die; this build reads an absent field as 0 and carries on.  (ONE such line does
    not crash the reference -- it parses whatever the previous line left behind those stale pointers -- and there
    the two agree, blanks-for-tabs case of the awkward-lines test.)"""
    bed = os.path.join(workdir, "create_bed_blank")
    _awkward_files(bed)
    with open(os.path.join(bed, "f00000.bed"), "a") as f:
        f.write("\n\n# a comment\ntrack name=x\n")
    got = outputs(api.CLI_PATH, workdir, bed, "mine_blank", [])
    assert "Save igd database to" in got["stdout"]
    db = O.OracleDB(os.path.join(workdir, "mine_blank", "db.igd"))
    assert db.nfiles == 12 and db.file_nr[0] == 304          # the four lines are counted in the tsv like any other


def test_create_survives_coordinates_beyond_int32(workdir):
    """`chr1 3000000000 3000000100` narrows to a negative interval: the reference indexes tile -79040 and
    segfaults; here the line is counted in the tsv and stored nowhere"""
    bed = os.path.join(workdir, "create_bed_overflow")
    _awkward_files(bed, with_overflow=True)
    got = outputs(api.CLI_PATH, workdir, bed, "mine_overflow", [])
    clean = os.path.join(workdir, "create_bed_awkward")
    if os.path.isdir(clean):
        base = outputs(api.CLI_PATH, workdir, clean, "mine_awkward_again", [])
        assert got["igd"] == base["igd"]                 # the extra line added no record
    assert "\t301\t" in got["tsv"].split("\n")[9]       # file 8 has one line more


def hdr_len(b: bytes) -> int:
    m = int(np.frombuffer(b, np.int32, 1, 8)[0])
    return 12 + 44 * m + 4 * int(np.frombuffer(b, np.int32, m, 12).sum())


def test_create_refuses_existing_and_prints_help(inputs, workdir):
    out = os.path.join(workdir, "mine_twice") + "/"
    run_create(api.CLI_PATH, inputs, out, "db")
    again = run_create(api.CLI_PATH, inputs, out, "db")
    assert again == f"The igd database file {out}db.igd exists!\n"
    r = subprocess.run([api.CLI_PATH, "create", inputs], capture_output=True, text=True)
    assert r.returncode == 0 and "usage:   igd create <input dir>" in r.stderr


def test_created_database_searches_like_the_reference_one(inputs, workdir):
    """end to end on the CPU side: the oracle gives the same hits[] on our database as the committed count"""
    out = os.path.join(workdir, "mine_search") + "/"
    run_create(api.CLI_PATH, inputs, out, "db")
    db = O.OracleDB(out + "db.igd")
    q = synth.gen_intervals(4242, 3000, ("lognormal", 7.0, 1.5), SMALL_SIZES)
    hits, _ = db.count(q.ctg, q.start, q.end)
    exp = json.load(open(os.path.join(GOLD, "expected.json")))
    assert hits.tolist() == exp["search_hits"]


def make_fixture():
    """python -m tests.test_create  (needs oracle/_ref): writes tests/golden/create/expected.json"""
    import tempfile
    assert O.have_ref()
    with tempfile.TemporaryDirectory(prefix="c") as work:
        bed = os.path.join(work, "create_bed")
        write_inputs(bed)
        exp = {}
        for case, opts in CASES.items():
            o = outputs(O.REF_BIN, work, bed, "ref_" + case, opts)
            exp[case] = {"igd_sha256": hashlib.sha256(o["igd"]).hexdigest(), "tsv": o["tsv"], "stdout": o["stdout"],
                         "listing": o["listing"]}
        db = O.OracleDB(os.path.join(work, "ref_default", "db.igd"))
        q = synth.gen_intervals(4242, 3000, ("lognormal", 7.0, 1.5), SMALL_SIZES)
        exp["search_hits"] = db.count(q.ctg, q.start, q.end)[0].tolist()
        os.makedirs(GOLD, exist_ok=True)
        with open(os.path.join(GOLD, "expected.json"), "w") as f:
            json.dump(exp, f, indent=1)
    print("wrote", os.path.join(GOLD, "expected.json"))


if __name__ == "__main__":
    make_fixture()


def test_create_abandons_when_an_input_does_not_open(workdir):
    """a dangling symlink among the inputs: the reference returns from create_igd at the first file that does not
    open (src/igd_create.c:57-58) and writes nothing; so does this build -- no crash, no _index.tsv, no empty
    db.igd that would make every re-run say 'exists!'"""
    bed = os.path.join(workdir, "create_bed_dangling")
    _awkward_files(bed)
    os.symlink(os.path.join(bed, "no_such_file.bed"), os.path.join(bed, "f00004x.bed"))
    out = os.path.join(workdir, "mine_dangling") + "/"
    for opts in ([], ["-s", "0"]):
        r = subprocess.run([api.CLI_PATH, "create", bed, out, "db", *opts], capture_output=True)
        assert r.returncode == 0, (r.returncode, r.stderr.decode())
        assert not os.path.exists(out + "db.igd") and not os.path.exists(out + "db_index.tsv")
    if O.have_ref():
        refout = os.path.join(workdir, "ref_dangling") + "/"
        r = subprocess.run([O.REF_BIN, "create", bed, refout, "db"], capture_output=True)
        assert not os.path.exists(refout + "db.igd")


def test_create_lines_wider_than_the_field_table(workdir):
    """50-column lines: the splitter's table holds 40 pointers and must not be overrun (the reference's 17-entry
    table is, and it aborts); fields 0-4 are all that create reads"""
    bed = os.path.join(workdir, "create_bed_wide")
    os.makedirs(bed)
    rng = np.random.default_rng(5)
    want = 0
    for i in range(12):
        with open(os.path.join(bed, f"f{i:05d}.bed"), "w") as f:
            for k in range(60):
                a = int(rng.integers(0, 900000))
                f.write("\t".join([f"chr{1 + k % 3}", str(a), str(a + 100), "n", str(k)] + ["x"] * 45) + "\n")
                want += 1 + ((a + 99) // 16384 - a // 16384)         # one record per bin touched
    out = os.path.join(workdir, "mine_wide") + "/"
    stdout = run_create(api.CLI_PATH, bed, out, "db")
    assert "Save igd database to" in stdout
    db = O.OracleDB(out + "db.igd")
    assert db.nctg == 3 and sorted(db.ctg_names) == ["chr1", "chr2", "chr3"]
    assert db.nrecords == want and sum(db.file_nr) == 12 * 60
