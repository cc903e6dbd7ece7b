"""examples/count_file.c -- the caller-side change INTEGRATION.md §2.2 describes -- compiles against the public
header, links the library, fails with the library's message on a box without a GPU and prints the reference's
hits[] on one with."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from igd_b200 import api
from tests.golden_util import GoldenCase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(workdir):
    out = os.path.join(workdir, "count_file")
    lib = os.path.dirname(api.LIB_PATH)
    r = subprocess.run([shutil.which("gcc") or "gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "count_file.c"), "-L", lib, "-ligd_b200",
                        f"-Wl,-rpath,{lib}", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_example_builds_and_fails_loudly_without_gpu(exe, workdir):
    case = GoldenCase("quirk", workdir)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage:" in r.stderr
    if api.load_library().igd_b200_device_count() <= 0:
        r = subprocess.run([exe, case.igd, case.qfile], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr and r.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("v", [0, 500])
def test_example_prints_the_reference_hits(exe, workdir, v):
    case = GoldenCase("dense", workdir)
    r = subprocess.run([exe, case.igd, case.qfile] + ([str(v)] if v else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = case.hits(v)
    lines = r.stdout.strip().split("\n")
    got = np.zeros_like(want)
    for ln in lines[:-1]:
        i, c = ln.split("\t")
        got[int(i)] = int(c)
    assert np.array_equal(got, want) and lines[-1] == f"total\t{int(want.sum())}"


@pytest.fixture(scope="module")
def sharded_exe(workdir):
    """examples/sharded_count.c: a multi-rank count search from plain C (fork, igd_b200_sharded_*, the CUDA runtime)"""
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("no CUDA toolkit headers")
    out = os.path.join(workdir, "sharded_count")
    lib = os.path.dirname(api.LIB_PATH)
    r = subprocess.run([shutil.which("gcc") or "gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "examples", "sharded_count.c"),
                        "-L", lib, "-ligd_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                        f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_sharded_example_builds_and_fails_loudly_without_gpu(sharded_exe, workdir):
    case = GoldenCase("quirk", workdir)
    r = subprocess.run([sharded_exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage:" in r.stderr
    if api.load_library().igd_b200_device_count() <= 0:
        r = subprocess.run([sharded_exe, case.igd, case.qfile, "2"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "no CUDA device" in r.stderr and r.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("world,v", [(2, 0), (3, 500), (1, 0)])
def test_sharded_example_prints_the_reference_hits(sharded_exe, workdir, world, v):
    """ranks as plain processes (sharing the GPU when the box has fewer): the golden hits[] of the reference"""
    case = GoldenCase("dense", workdir)
    r = subprocess.run([sharded_exe, case.igd, case.qfile, str(world)] + ([str(v)] if v else []),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    want = case.hits(v)
    lines = r.stdout.strip().split("\n")
    got = np.zeros_like(want)
    for ln in lines[:-1]:
        i, c = ln.split("\t")
        got[int(i)] = int(c)
    assert np.array_equal(got, want) and lines[-1] == f"total\t{int(want.sum())}"
