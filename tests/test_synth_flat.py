"""write_igd_flat (the contig-by-contig writer behind the 1e8 / 1e9-interval benchmark databases) writes the same
bytes as write_igd_direct, which tests/test_synth.py holds to the reference's `igd create`."""
import numpy as np

from igd_b200 import synth


def test_flat_writer_equals_direct_writer(tmp_path):
    sizes = np.array([900_000, 40_000, 500_000, 120_000], dtype=np.int64)
    names = ["chr1", "chrQ", "chr2", "chr3"]
    per_file, n_files = 3000, 14
    files = [synth.gen_intervals(4100 + i, per_file, ("lognormal", 7.0, 1.6), sizes) for i in range(n_files)]
    files[3].end[5:40] = files[3].start[5:40]                    # dropped by igd_add (s >= e), still in the tsv
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    synth.write_igd_direct(a, files, names=names)
    cat = lambda k, dt: np.concatenate([getattr(f, k) for f in files]).astype(dt)
    st = synth.write_igd_flat(b, cat("ctg", np.int8), cat("start", np.int32), cat("end", np.int32),
                              cat("value", np.int16), per_file, names=names, device="cpu")
    assert open(a + ".igd", "rb").read() == open(b + ".igd", "rb").read()
    assert open(a + "_index.tsv").read() == open(b + "_index.tsv").read()
    assert st["n_intervals"] == per_file * n_files
