import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


from igd_b200 import synth  # noqa: E402

SMALL_NAMES = ["chrA", "chrB", "chrC"]
SMALL_SIZES = np.array([1_000_000, 500_000, 200_000], dtype=np.int64)


def make_dense_db(stem, n_files=12, n_per_file=3000, seed=7, law=("lognormal", 6.0, 1.2)):
    """A small, dense database (hundreds of records per tile) on a 3-contig toy genome."""
    files = [synth.gen_intervals(seed * 1000 + i, n_per_file, law, SMALL_SIZES) for i in range(n_files)]
    synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    return files


def make_queries(seed, n, law, sizes=SMALL_SIZES):
    return synth.gen_intervals(seed, n, law, sizes)


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("igd"))


@pytest.fixture(scope="session")
def dense_db(workdir):
    stem = os.path.join(workdir, "dense")
    make_dense_db(stem)
    return stem + ".igd"


@pytest.fixture(scope="session")
def sparse_db(workdir):
    """BASELINE config 1 shape: 10 files x 10k lognormal intervals on hg38 (0.5 records per tile)."""
    stem = os.path.join(workdir, "cfg1")
    law = ("lognormal", 6.5, 1.2)
    files = [synth.gen_intervals(synth.seed_for(1, i), 10000, law) for i in range(10)]
    synth.write_igd_direct(stem, files)
    return stem + ".igd"
