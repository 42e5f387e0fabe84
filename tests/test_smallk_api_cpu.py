"""The part of the smallk:: C++ API that needs no device (settings and their clamps, LoadMatrix overloads, LoadDictionary,
SetOutputDir, the argument checks of Nmf / HierNmf2), probed by tests/cpp/smallk_api_probe.cpp built against the host layer,
and compared line by line — return values, exception types, exception texts — with
  * the transcript of the same probe built against the reference's own smallk.hpp / smallk.cpp (tests/golden/
    smallk_api_probe_ref.txt, generated here by this test's `regenerate` path), and
  * where /root/reference and oracle/_ref are present, a live build of that reference probe."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PROBE = os.path.join(HERE, "cpp", "smallk_api_probe.cpp")
GOLDEN = os.path.join(HERE, "golden", "smallk_api_probe_ref.txt")
REF_INC = "/root/reference/smallk/include"
REF_LIB = os.path.join(ROOT, "oracle", "_ref")
HOST_LIB = os.path.join(ROOT, "smallk_b200", "lib")


def _inputs(d):
    with open(os.path.join(d, "a.csv"), "w") as f:
        f.write("1,2,3,4\n5,6,7,8\n9,10,11,12\n")
    with open(os.path.join(d, "a.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n3 3 3\n1 1 1.0\n2 2 2.0\n3 3 3.0\n")
    with open(os.path.join(d, "bad.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n3 3 2\n1 1 1.0\n")
    with open(os.path.join(d, "dict.txt"), "w") as f:
        f.write("a\nb\nc\n")


def _run(exe, d):
    out = subprocess.run([exe, d], capture_output=True, text=True, cwd=d, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("PROBE|")]
    return [ln.replace(d, "<DIR>") for ln in lines]


def _build(exe, includes, libdir, libs):
    cmd = ["g++", "-std=c++14", "-O1", "-o", exe, PROBE] + ["-I" + i for i in includes] + ["-L" + libdir] + ["-l" + x for x in libs] + ["-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.fixture(scope="module")
def host_transcript(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(HOST_LIB, "libsmallk_host.so")):
        pytest.skip("g++ or the host library is missing")
    d = str(tmp_path_factory.mktemp("probe"))
    _inputs(d)
    exe = os.path.join(d, "probe_host")
    _build(exe, [os.path.join(ROOT, "smallk_b200", "host")], HOST_LIB, ["smallk_host", "smallk_b200"])
    return _run(exe, d), d


def test_host_api_matches_reference_transcript(host_transcript):
    got, _ = host_transcript
    want = open(GOLDEN).read().splitlines()
    assert len(got) == len(want) and len(got) >= 50
    for g, w in zip(got, want):
        assert g == w


def test_host_api_matches_live_reference_probe(host_transcript):
    got, d = host_transcript
    if not (os.path.isdir(REF_INC) and os.path.exists(os.path.join(REF_LIB, "libsmallk_ref.so"))):
        pytest.skip("/root/reference or oracle/_ref not present on this machine")
    exe = os.path.join(d, "probe_ref")
    _build(exe, [REF_INC, "/root/reference/common/include", "/root/reference/flatclust/include"], REF_LIB, ["smallk_ref"])
    want = _run(exe, d)
    if os.environ.get("SMK_REGENERATE_GOLDEN"):
        with open(GOLDEN, "w") as f:
            f.write("\n".join(want) + "\n")
    assert got == want
