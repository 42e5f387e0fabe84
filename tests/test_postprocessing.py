"""Clustering post-processing, no GPU needed: ComputeAssignments, ComputeFuzzyAssignments
(common/include/assignments.hpp:32-113) and TopTerms (common/include/terms.hpp:62-108) of the host layer against the
reference's own functions compiled into oracle/_ref — identical integers, identical float bits, ties included (TopTerms
uses an unstable sort: the same libstdc++ call on the same data must be made for ties to fall the same way)."""
import ctypes
import os

import numpy as np
import pytest

import smallk_b200 as sk

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libsmallk_ref.so")
dp = ctypes.POINTER(ctypes.c_double)


def _libs():
    if not os.path.exists(sk.HOST_LIB_PATH) or not os.path.exists(REF_SO):
        pytest.skip("host library or oracle/_ref not built on this machine")
    ref = ctypes.CDLL(REF_SO)
    if not hasattr(ref, "ref_compute_assignments"):
        pytest.skip("oracle/_ref predates the post-processing entry points")
    return ctypes.CDLL(sk.HOST_LIB_PATH), ref


def _factors():
    rng = np.random.default_rng(12)
    for m, n, k, ld_extra in ((40, 60, 4, 0), (300, 200, 16, 3), (1000, 50, 2, 0), (64, 64, 64, 1)):
        W = np.zeros((m + ld_extra, k), order="F"); W[:m] = rng.random((m, k))
        H = np.zeros((k + ld_extra, n), order="F"); H[:k] = rng.random((k, n))
        W[:m][rng.random((m, k)) < 0.3] = 0.0                 # exact zeros, as NNLS leaves them
        H[:k][rng.random((k, n)) < 0.3] = 0.0
        W[: m // 2, 0] = np.round(W[: m // 2, 0], 1)          # ties among the term weights
        H[:k, ::5] = 0.25                                     # ties among the cluster weights of a document
        H[:k, 7] = 0.0                                        # a document with no weight at all
        yield m, n, k, W, H


def test_assignments_match_reference():
    host, ref = _libs()
    for m, n, k, W, H in _factors():
        a, b = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        up = ctypes.POINTER(ctypes.c_uint)
        ref.ref_compute_assignments(H.ctypes.data_as(dp), H.shape[0], k, n, a.ctypes.data_as(up))
        host.smkh_compute_assignments(H.ctypes.data_as(dp), H.shape[0], k, n, b.ctypes.data_as(up))
        assert np.array_equal(a, b), (m, n, k)


def test_fuzzy_assignments_match_reference_bit_for_bit():
    host, ref = _libs()
    fp = ctypes.POINTER(ctypes.c_float)
    for m, n, k, W, H in _factors():
        Hc = np.asfortranarray(H[:k])                         # the reference indexes its output with the same ldim: tight H
        a, b = np.zeros(k * n, dtype=np.float32), np.zeros(k * n, dtype=np.float32)
        ref.ref_compute_fuzzy_assignments(Hc.ctypes.data_as(dp), k, k, n, a.ctypes.data_as(fp))
        host.smkh_compute_fuzzy_assignments(Hc.ctypes.data_as(dp), k, k, n, b.ctypes.data_as(fp))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (m, n, k)


def test_top_terms_match_reference_including_ties():
    host, ref = _libs()
    ip = ctypes.POINTER(ctypes.c_int)
    for m, n, k, W, H in _factors():
        # tight W (ldim == height), the only way the reference's callers pass it: its TopTerms steps from column to column
        # by `height`, not `ldim` (terms.hpp:93), so a padded buffer would make it rank shifted data. The host version
        # honours ldim; with a tight buffer the two must agree, ties included.
        Wt = np.asfortranarray(W[:m])
        for maxterms in (1, 5, min(m, 12)):
            a, b = np.zeros(maxterms * k, dtype=np.int32), np.zeros(maxterms * k, dtype=np.int32)
            ref.ref_top_terms_matrix(maxterms, Wt.ctypes.data_as(dp), m, m, k, a.ctypes.data_as(ip))
            host.smkh_top_terms_matrix(maxterms, Wt.ctypes.data_as(dp), m, m, k, b.ctypes.data_as(ip))
            assert np.array_equal(a, b), (m, k, maxterms)
            if W.shape[0] != m:                               # padded buffer: the host version reads the same columns
                c = np.zeros(maxterms * k, dtype=np.int32)
                host.smkh_top_terms_matrix(maxterms, W.ctypes.data_as(dp), W.shape[0], m, k, c.ctypes.data_as(ip))
                assert np.array_equal(b, c), (m, k, maxterms)


def test_random_initialisers_are_the_reference_sequential_stream():
    """Random + RandomMatrix (common/include/random.hpp, matrix_generator.hpp:61-82,229-248): the W then H draws of the
    clustering drivers from one seeded generator. With one thread the reference uses the sequential generator, whose stream
    the host layer reproduces at every size; with several threads and >= 16k elements the reference switches to per-thread
    seeds (a different stream: not reproduced)."""
    host, ref = _libs()
    if not hasattr(ref, "ref_random_matrices"):
        pytest.skip("oracle/_ref predates the random-initialiser entry point")
    for seed, (h1, w1), (h2, w2) in ((1, (50, 2), (2, 70)), (32, (9000, 2), (2, 9000)), (7, (300, 16), (16, 200)), (5, (20000, 2), (2, 3))):
        a1, a2 = np.zeros(h1 * w1), np.zeros(h2 * w2)
        b1, b2 = np.zeros(h1 * w1), np.zeros(h2 * w2)
        ia, ib = ctypes.c_int(0), ctypes.c_int(0)
        ref.ref_random_matrices(seed, 1, h1, w1, a1.ctypes.data_as(dp), h2, w2, a2.ctypes.data_as(dp), ctypes.byref(ia))
        host.smkh_random_matrices(seed, h1, w1, b1.ctypes.data_as(dp), h2, w2, b2.ctypes.data_as(dp), ctypes.byref(ib))
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2) and ia.value == ib.value, (seed, h1, w1)
    # the parallel generator is a different stream (this is why C4-scale trees are compared with the one-thread reference)
    a1, a2 = np.zeros(20000 * 2), np.zeros(2 * 3)
    ref.ref_random_matrices(5, 4, 20000, 2, a1.ctypes.data_as(dp), 2, 3, a2.ctypes.data_as(dp), None)
    assert not np.array_equal(a1, b1)
    ref.ref_random_matrices(5, 1, 4, 2, a1.ctypes.data_as(dp), 2, 3, a2.ctypes.data_as(dp), None)     # leave the thread count at 1


def test_option_validation_agrees_with_reference(capfd):
    """IsValid(NmfOptions / ClustOptions / FlatClustOptions): common/src/nmf_options.cpp:23-110, hierclust/src/clust_options.cpp:24-110,
    flatclust/src/flat_clust_options.cpp — one field at a time pushed out of range, with and without matrix validation."""
    host, ref = _libs()
    if not hasattr(ref, "ref_is_valid"):
        pytest.skip("oracle/_ref predates the option-validation entry point")
    #        tol   alg prog  h    w   k  min max tolc thr maxterms unbal trial clusters
    good = [1e-4, 1,  0,  100, 80,  4,  5, 500,  1,  1,   5,     0.1,   3,    4]
    edits = [(0, 0.0), (0, 1.0), (0, -1.0), (0, 0.5), (1, 7), (1, 3), (1, 0), (1, 2), (2, 5), (2, 1), (3, 0), (3, -4), (4, 0), (4, 3), (5, 0),
             (5, 81), (5, 80), (5, 2), (6, 0), (7, 0), (7, -1), (8, 0), (9, 0), (9, -1), (10, 0), (10, -3), (11, -0.1), (11, 1.0), (11, 0.0),
             (12, -1), (12, 0), (13, 1), (13, 0), (13, 2), (13, 500)]
    checked = 0
    for which in (0, 1, 2):
        for vm in (1, 0):
            for idx, val in [(None, None)] + edits:
                for alg in (None, 3):                         # 3 = RANK2: k must be 2
                    v = list(good)
                    if alg is not None:
                        v[1] = alg
                    if idx is not None:
                        v[idx] = val
                    arr = np.array(v, dtype=np.float64)
                    a = ref.ref_is_valid(which, arr.ctypes.data_as(dp), vm)
                    b = host.smkh_is_valid(which, arr.ctypes.data_as(dp), vm)
                    assert a == b, (which, vm, idx, val, alg, a, b)
                    checked += 1
    capfd.readouterr()                                        # both sides explain every rejection on stderr
    assert checked > 400


@pytest.mark.parametrize("fmt", [1, 2])           # FileFormat::XML, FileFormat::JSON
def test_flatclust_result_files_are_byte_identical(tmp_path, fmt):
    """FlatClustWriteResults (common/src/flat_clust_output.cpp:52-170): assignments_flat_<k>.csv, assignments_fuzzy_<k>.csv and
    clusters_<k>.{xml,json} written by the host layer and by the reference from the same inputs — an empty cluster included."""
    host, ref = _libs()
    if not hasattr(ref, "ref_flatclust_write_results"):
        pytest.skip("oracle/_ref predates the result-writer entry point")
    rng = np.random.default_rng(3)
    up, fp, ip = ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    for k, n, m, maxterms in ((4, 30, 50, 5), (6, 12, 20, 3), (2, 9, 10, 1)):
        assign = rng.integers(0, k, size=n).astype(np.uint32)
        assign[assign == k - 1] = 0                            # the last cluster receives no document
        prob = rng.random(k * n).astype(np.float32)
        terms = rng.integers(0, m, size=k * maxterms).astype(np.int32)
        words = [f"term{i}_{'x' * (i % 4)}" for i in range(m)]
        blob = b"".join(w.encode() + b"\0" for w in words)
        dirs = []
        for name, fn in (("ref", ref.ref_flatclust_write_results), ("host", host.smkh_flatclust_write_results)):
            d = tmp_path / f"{name}_{k}_{fmt}"
            d.mkdir()
            rc = fn(str(d).encode(), assign.ctypes.data_as(up), prob.ctypes.data_as(fp), blob, m, terms.ctypes.data_as(ip), fmt, maxterms, n, k)
            assert rc == 0, name
            dirs.append(d)
        names = sorted(p.name for p in dirs[0].iterdir())
        assert names == sorted(p.name for p in dirs[1].iterdir()) and len(names) == 3, names
        for nm in names:
            assert (dirs[0] / nm).read_bytes() == (dirs[1] / nm).read_bytes(), (k, fmt, nm)


@pytest.mark.parametrize("fmt", [1, 2])           # FileFormat::XML, FileFormat::JSON
def test_tree_and_its_writers_are_byte_identical(tmp_path, fmt):
    """Tree<T> (hierclust/include/tree.hpp, src/tree.cpp: SplitRoot / Split / PartitionDocs / MinMaxLeafPriorities /
    ComputeTopTerms / ComputeAssignments / WriteAssignments / WriteTree) and the XML / JSON tree writers, driven on both sides
    by the same script of factors and priorities (no factorization): the assignment file and the tree file must be identical
    byte for byte — zero memberships, ties (H(0,c) == H(1,c) == 0), unsplittable small leaves and outliers included."""
    host, ref = _libs()
    if not hasattr(ref, "ref_tree_script"):
        pytest.skip("oracle/_ref predates the tree-script entry point")
    # (4, 50, 9, 5, 2) cannot grow its five leaves from nine documents: the reference then prints never-created node slots
    # with uninitialised parent / child fields (garbage that changes run to run); the host tree prints -1 / false there.
    # For that case only the assignment file is compared.
    for seed, m, n, clusters, maxterms in ((1, 30, 40, 4, 5), (2, 80, 200, 12, 5), (3, 25, 60, 16, 3), (4, 50, 9, 5, 2), (5, 200, 1000, 30, 8)):
        complete = (seed != 4)
        files = []
        for name, fn in (("ref", ref.ref_tree_script), ("host", host.smkh_tree_script)):
            a, t = tmp_path / f"{name}_{seed}_{fmt}_assign.csv", tmp_path / f"{name}_{seed}_{fmt}_tree.out"
            rc = fn(seed, m, n, clusters, maxterms, fmt, str(a).encode(), str(t).encode())
            assert rc == 0, (name, seed, rc)
            files.append((a.read_bytes(), t.read_bytes()))
        assert files[0][0] == files[1][0], ("assignments", seed, fmt)
        if complete:
            assert files[0][1] == files[1][1], ("tree", seed, fmt)
        assert len(files[0][1]) > 100
        # the same script through the hierclust driver's own tree operations (SplitCompact: W on its non-zero rows; a decoy split
        # taken back by UndoSplit before every real one): not a byte may change
        a, t = tmp_path / f"compact_{seed}_{fmt}_assign.csv", tmp_path / f"compact_{seed}_{fmt}_tree.out"
        rc = host.smkh_tree_script_compact(seed, m, n, clusters, maxterms, fmt, str(a).encode(), str(t).encode())
        assert rc == 0, ("compact", seed, rc)
        assert a.read_bytes() == files[1][0] and t.read_bytes() == files[1][1], ("compact", seed, fmt)
