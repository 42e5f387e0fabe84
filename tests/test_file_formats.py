"""File formats either side of the NMF path, no GPU needed: the host readers / writers
(smallk_b200/host/matrix_io.hpp) against the reference's own (common/include/sparse_matrix_io.hpp:117-260,
delimited_file.hpp:49-135) compiled into oracle/_ref. The CSC arrays a MatrixMarket file turns into (expansion of symmetric
and skew-symmetric files, pattern values, stable order inside a column, duplicates kept), the buffer a CSV file turns into,
and the bytes a CSV writer produces must be identical."""
import ctypes
import os

import numpy as np
import pytest

import smallk_b200 as sk

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libsmallk_ref.so")
up = ctypes.POINTER(ctypes.c_uint)
dp = ctypes.POINTER(ctypes.c_double)


def _libs():
    if not os.path.exists(sk.HOST_LIB_PATH) or not os.path.exists(REF_SO):
        pytest.skip("host library or oracle/_ref not built on this machine")
    ref = ctypes.CDLL(REF_SO)
    if not hasattr(ref, "ref_load_matrix_market"):
        pytest.skip("oracle/_ref predates the file-format entry points")
    return ctypes.CDLL(sk.HOST_LIB_PATH), ref


def _load_mtx(fn, path, cap_cols=4096, cap_nz=1 << 16):
    h, w, nz = ctypes.c_uint(0), ctypes.c_uint(0), ctypes.c_uint(0)
    colp = np.zeros(cap_cols, dtype=np.uint32); rowi = np.zeros(cap_nz, dtype=np.uint32); val = np.zeros(cap_nz)
    rc = fn(path.encode(), ctypes.byref(h), ctypes.byref(w), ctypes.byref(nz), colp.ctypes.data_as(up), cap_cols,
            rowi.ctypes.data_as(up), val.ctypes.data_as(dp), cap_nz)
    if rc != 0:
        return rc, None
    return 0, (h.value, w.value, nz.value, colp[: w.value + 1].copy(), rowi[: nz.value].copy(), val[: nz.value].copy())


def _write_mtx(path, header, m, n, entries, comments=("% a comment",), blank_tail=False):
    with open(path, "w") as f:
        f.write(header + "\n")
        for c in comments:
            f.write(c + "\n")
        f.write(f"{m} {n} {len(entries)}\n")
        for e in entries:
            f.write(" ".join(str(x) for x in e) + "\n")
        if blank_tail:
            f.write("\n")


def _cases(tmp):
    rng = np.random.default_rng(9)
    out = []
    # general real, unsorted, with duplicates
    m, n = 40, 30
    ent = [(int(rng.integers(1, m + 1)), int(rng.integers(1, n + 1)), repr(float(rng.random()))) for _ in range(300)]
    ent += [ent[3], ent[10]]
    p = os.path.join(tmp, "general.mtx"); _write_mtx(p, "%%MatrixMarket matrix coordinate real general", m, n, ent); out.append(p)
    # symmetric pattern (what a graph file looks like), lower triangle + diagonal entries
    m = 50
    ent = sorted({(int(max(a, b)), int(min(a, b))) for a, b in rng.integers(1, m + 1, size=(200, 2))})
    p = os.path.join(tmp, "sym_pattern.mtx"); _write_mtx(p, "%%MatrixMarket matrix coordinate pattern symmetric", m, m, ent); out.append(p)
    # skew-symmetric real
    ent = [(int(a), int(b), repr(float(v))) for (a, b), v in zip(sorted({(int(max(a, b)), int(min(a, b))) for a, b in rng.integers(1, m + 1, size=(120, 2)) if a != b}), rng.random(200))]
    p = os.path.join(tmp, "skew.mtx"); _write_mtx(p, "%%MatrixMarket matrix coordinate real skew-symmetric", m, m, ent); out.append(p)
    # integer general, mixed-case banner, several comment lines
    ent = [(int(rng.integers(1, 21)), int(rng.integers(1, 26)), int(rng.integers(1, 9))) for _ in range(90)]
    p = os.path.join(tmp, "integer.mtx")
    _write_mtx(p, "%%MatrixMarket MATRIX Coordinate Integer General", 20, 25, ent, comments=("%", "% two", "%three")); out.append(p)
    # symmetric real with exponents and extra spaces
    ent = [(7, 2, "1.5e-3"), (9, 9, " 2.25E+1"), (10, 1, "-4.0"), (3, 3, "0.125")]
    p = os.path.join(tmp, "sym_real.mtx"); _write_mtx(p, "%%MatrixMarket matrix coordinate real symmetric", 10, 10, ent); out.append(p)
    return out


def test_matrix_market_reader_builds_the_reference_csc(tmp_path):
    host, ref = _libs()
    for path in _cases(str(tmp_path)):
        rc_r, want = _load_mtx(ref.ref_load_matrix_market, path)
        rc_h, got = _load_mtx(host.smkh_load_matrix_market, path)
        assert rc_r == 0 and rc_h == 0, (path, rc_r, rc_h)
        assert got[:3] == want[:3], (path, got[:3], want[:3])
        for a, b in zip(got[3:], want[3:]):
            assert np.array_equal(a, b), path


def test_matrix_market_reader_rejects_what_the_reference_rejects(tmp_path):
    host, ref = _libs()
    bad = {
        "array.mtx": "%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
        "complex.mtx": "%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1.0 0.0\n",
        "hermitian.mtx": "%%MatrixMarket matrix coordinate real hermitian\n2 2 1\n1 1 1.0\n",
        "nobanner.mtx": "2 2 1\n1 1 1.0\n",
        "zeroindex.mtx": "%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 1.0\n",
    }
    for name, text in bad.items():
        p = os.path.join(str(tmp_path), name)
        open(p, "w").write(text)
        rc_r, _ = _load_mtx(ref.ref_load_matrix_market, p)
        rc_h, _ = _load_mtx(host.smkh_load_matrix_market, p)
        assert rc_r != 0, name
        assert rc_h != 0, name
    assert _load_mtx(host.smkh_load_matrix_market, os.path.join(str(tmp_path), "missing.mtx"))[0] != 0


def test_csv_writer_and_reader_match_the_reference(tmp_path):
    host, ref = _libs()
    rng = np.random.default_rng(4)
    for (h, w), prec in (((7, 5), 6), ((1, 9), 4), ((12, 1), 12), ((33, 17), 6)):
        A = np.asfortranarray(rng.standard_normal((h, w)) * 10.0 ** rng.integers(-8, 8, size=(h, w)))
        A[0, 0] = 0.0
        pr, ph = os.path.join(str(tmp_path), f"ref_{h}_{w}.csv"), os.path.join(str(tmp_path), f"host_{h}_{w}.csv")
        assert ref.ref_write_delimited(A.ctypes.data_as(dp), h, h, w, pr.encode(), prec) == 0
        assert host.smkh_write_delimited(A.ctypes.data_as(dp), h, h, w, ph.encode(), prec) == 0
        assert open(pr, "rb").read() == open(ph, "rb").read(), (h, w, prec)
        for fn in (ref.ref_load_delimited, host.smkh_load_delimited):
            hh, ww = ctypes.c_uint(0), ctypes.c_uint(0)
            buf = np.zeros(h * w)
            assert fn(pr.encode(), ctypes.byref(hh), ctypes.byref(ww), buf.ctypes.data_as(dp), h * w) == 0
            assert (hh.value, ww.value) == (h, w)
            back = buf.reshape((w, h)).T
            assert np.allclose(back, A, rtol=10.0 ** (-prec), atol=0), (h, w)
        # both readers produce the same bits from the same file
        b1, b2 = np.zeros(h * w), np.zeros(h * w)
        hh, ww = ctypes.c_uint(0), ctypes.c_uint(0)
        ref.ref_load_delimited(pr.encode(), ctypes.byref(hh), ctypes.byref(ww), b1.ctypes.data_as(dp), h * w)
        host.smkh_load_delimited(pr.encode(), ctypes.byref(hh), ctypes.byref(ww), b2.ctypes.data_as(dp), h * w)
        assert np.array_equal(b1, b2)


MM = "%%MatrixMarket matrix coordinate real general\n"
MALFORMED_MTX = {
    "trailing_blank": MM + "3 3 2\n1 1 1.5\n3 2 2.5\n\n",          # a blank line counts as an entry line: count mismatch
    "middle_blank": MM + "3 3 2\n1 1 1.5\n\n3 2 2.5\n",
    "too_few": MM + "3 3 3\n1 1 1.5\n3 2 2.5\n",
    "too_many": MM + "3 3 1\n1 1 1.5\n3 2 2.5\n",
    "no_newline_end": MM + "3 3 2\n1 1 1.5\n3 2 2.5",               # accepted
    "comment_after_size": MM + "3 3 2\n% c\n1 1 1.5\n3 2 2.5\n",
    "blank_before_size": MM + "\n3 3 2\n1 1 1.5\n3 2 2.5\n",        # accepted
    "row_out_of_range": MM + "3 3 1\n4 1 1.0\n",                    # throws
    "col_out_of_range": MM + "3 3 1\n1 5 1.0\n",                    # throws
    "negative_index": MM + "3 3 1\n-1 1 1.0\n",                     # wraps to a huge index: throws
    "no_entries": MM + "3 3 0\n",                                   # throws
    "crlf": MM.replace("\n", "\r\n") + "3 3 2\r\n1 1 1.5\r\n3 2 2.5\r\n",
    "pattern_with_values": "%%MatrixMarket matrix coordinate pattern general\n3 3 2\n1 1 9\n2 2\n",
    "symmetric_upper": "%%MatrixMarket matrix coordinate real symmetric\n3 3 2\n1 2 1.0\n3 3 2.0\n",
    "symmetric_out_of_range": "%%MatrixMarket matrix coordinate real symmetric\n2 3 1\n1 3 1.0\n",
    "lowercase_banner": "%%matrixmarket matrix coordinate real general\n3 3 1\n1 1 1.0\n",
    "tabs": MM + "3\t3\t1\n1\t1\t1.0\n",
}


def test_matrix_market_reader_on_malformed_files_behaves_like_the_reference(tmp_path):
    """Accept / reject / throw, and what is built when accepted, file by file. (A data line without a value is left out: the
    reference reads an uninitialised variable there.)"""
    host, ref = _libs()
    if not hasattr(ref, "ref_io_last_exception"):
        pytest.skip("oracle/_ref predates the exception-reporting entry point")
    ref.ref_io_last_exception.restype = ctypes.c_char_p
    host.smkh_io_last_exception.restype = ctypes.c_char_p
    for name, text in MALFORMED_MTX.items():
        p = os.path.join(str(tmp_path), name + ".mtx")
        with open(p, "w", newline="") as f:
            f.write(text)
        rc_r, want = _load_mtx(ref.ref_load_matrix_market, p)
        rc_h, got = _load_mtx(host.smkh_load_matrix_market, p)
        assert rc_r == rc_h, (name, rc_r, rc_h)
        if rc_r == 0:
            assert got[:3] == want[:3], name
            for a, b in zip(got[3:], want[3:]):
                assert np.array_equal(a, b), name
        elif rc_r == -3:                                       # both threw: the same exception text
            assert ref.ref_io_last_exception() == host.smkh_io_last_exception(), name


MALFORMED_CSV = {
    "crlf": "1,2,3\r\n4,5,6\r\n", "blank_lead": "\n\n1,2\n3,4\n", "comment_hash": "# hello\n1,2\n3,4\n", "comment_pct": "% hello\n1,2\n3,4\n",
    "trailing_blank": "1,2\n3,4\n\n", "middle_blank": "1,2\n\n3,4\n", "spaces": "1, 2 ,3\n 4,5, 6\n", "sci": "1e-3,-2.5E+2,+3\n.5,5.,-0.0\n",
    "ragged": "1,2,3\n4,5\n", "trailing_comma": "1,2,\n3,4,\n", "single": "7\n", "one_row": "1,2,3,4\n", "one_col": "1\n2\n3\n",
    "empty": "", "only_comments": "# a\n% b\n", "text": "a,b\n1,2\n", "no_newline_end": "1,2\n3,4", "only_line_no_newline": "1,2",
    "tabs": "1\t2\n3\t4\n", "nan_inf": "nan,inf\n1,2\n", "comment_in_middle": "1,2\n# x\n3,4\n",
}


def test_csv_reader_on_malformed_files_behaves_like_the_reference(tmp_path):
    """The reference's reader defines the format by what it does (blank rows are rows, a last line without a newline is not,
    short rows are padded with zeros, ...): same accept / reject decision, same shape, same buffer."""
    host, ref = _libs()
    for name, text in MALFORMED_CSV.items():
        p = os.path.join(str(tmp_path), name + ".csv")
        with open(p, "w", newline="") as f:
            f.write(text)
        res = []
        for fn in (ref.ref_load_delimited, host.smkh_load_delimited):
            h, w = ctypes.c_uint(0), ctypes.c_uint(0)
            buf = np.full(64, -99.0)
            rc = fn(p.encode(), ctypes.byref(h), ctypes.byref(w), buf.ctypes.data_as(dp), 64)
            res.append((rc, h.value, w.value, buf[: h.value * w.value].tolist() if rc == 0 else None))
        assert res[0] == res[1], (name, res)


def test_dictionary_reader_matches_reference(tmp_path):
    """LoadStringsFromFile (common/src/utils.cpp:220-239): one string per line, appended to what the vector holds."""
    host, ref = _libs()
    if not hasattr(ref, "ref_load_strings"):
        pytest.skip("oracle/_ref predates the dictionary entry point")
    cases = {"plain": "alpha\nbeta\ngamma\n", "no_newline_end": "alpha\nbeta\ngamma", "crlf": "alpha\r\nbeta\r\n", "blank_lines": "alpha\n\nbeta\n\n",
             "spaces": "two words\n  padded  \n", "empty": "", "single_no_newline": "alpha"}
    for name, text in cases.items():
        p = os.path.join(str(tmp_path), name + ".txt")
        with open(p, "w", newline="") as f:
            f.write(text)
        res = []
        for fn in (ref.ref_load_strings, host.smkh_load_strings):
            buf = ctypes.create_string_buffer(4096)
            cnt = ctypes.c_int(0)
            rc = fn(p.encode(), buf, 4096, ctypes.byref(cnt))
            res.append((rc, cnt.value, buf.value))
        assert res[0] == res[1], (name, res)
    missing = os.path.join(str(tmp_path), "missing.txt").encode()
    buf = ctypes.create_string_buffer(64); cnt = ctypes.c_int(0)
    assert ref.ref_load_strings(missing, buf, 64, ctypes.byref(cnt)) == host.smkh_load_strings(missing, buf, 64, ctypes.byref(cnt)) == -1
