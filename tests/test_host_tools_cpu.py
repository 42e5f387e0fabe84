"""The command-line tools (nmf, hierclust, flatclust) and the smallk:: example program on the CPU: the tests of
tests/test_gpu_host_api.py run a second time on copies of the tools linked against the CPU mock of the C ABI
(tests/cpp/mock_capi.cpp, the oracle's solvers behind the ABI's entry points) — what is exercised is everything above the ABI: option
handling, CSV / MatrixMarket / dictionary readers, the Nmf / Clust / FlatClust / smallk:: call chains, the result writers.
Test infrastructure only: nothing in smallk_b200/ links the mock."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import test_gpu_host_api as T                                       # noqa: E402
from test_gpu_host_api import c1, hier_case                           # noqa: E402,F401  (fixtures)
from test_host_driver_cpu import mock_host, BUILD, HOST               # noqa: E402,F401


@pytest.fixture(scope="module", autouse=True)
def tools_on_the_mock(mock_host):
    bindir = os.path.join(BUILD, "bin")
    os.makedirs(bindir, exist_ok=True)
    for exe, src in (("nmf", "nmf_cli.cpp"), ("hierclust", "hierclust_cli.cpp"), ("flatclust", "flatclust_cli.cpp"),
                     ("smallk_example", "smallk_example.cpp")):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-pthread", "-o", os.path.join(bindir, exe), os.path.join(HOST, src),
                               "-L" + BUILD, "-lsmallk_host_mock", "-lsmallk_mock", "-Wl,-rpath," + BUILD])
    saved = T.BIN
    T.BIN = bindir
    yield bindir
    T.BIN = saved


def test_nmf_cli_bpp_c1_on_the_mock(c1, oracle):
    T.test_nmf_cli_bpp_c1(c1, oracle)


def test_nmf_cli_rejects_bad_options_on_the_mock(c1):
    T.test_nmf_cli_rejects_bad_options(c1)


def test_smallk_api_hals_sparse_mtx_on_the_mock(c1, oracle, tmp_path):
    T.test_smallk_api_hals_sparse_mtx(c1, oracle, tmp_path)


def test_hierclust_cli_matches_reference_fixture_on_the_mock(hier_case):
    T.test_hierclust_cli_matches_reference_fixture(hier_case)


def test_smallk_api_hiernmf2_writes_tree_and_assignments_on_the_mock(hier_case):
    T.test_smallk_api_hiernmf2_writes_tree_and_assignments(hier_case)


def test_flatclust_cli_matches_oracle_on_the_mock(oracle, tmp_path):
    T.test_flatclust_cli_matches_oracle(oracle, tmp_path)
