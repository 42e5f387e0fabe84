"""TEST INFRASTRUCTURE — not part of the product.

NumPy restatement of the reference's tf-idf preprocessing, the upstream producer of the sparse
matrices the NMF path factors: preprocess_tf, preprocessor/src/preprocess.cpp:81-250, with

    PruneRows      :278-369   a term stays iff its total count >= docs_per_term AND it does not occur in every document
    PrunableCols   :372-401   a document stays iff it has >= terms_per_doc distinct terms
    UniqueCols     :665-760   of a group of identical columns (same rows, same counts, rows sorted) the one with the LARGEST
                              column index stays (:561-565, :641-655); the reference finds the groups by SpookyHash + exact
                              comparison, which is an implementation detail of "identical"
    PruneCols      :404-445   compaction, order kept
    scores         :193-230   (1 + ln count) * ln(width / document frequency), every column scaled to unit 2-norm

The iteration structure (prune rows; prune columns or else remove duplicates; stop when nothing was removed or after max_iter
rounds) follows :126-180 step by step. Checked against the reference's own function in tests/test_oracle_preprocess.py.

One corner is NOT reproduced: when column pruning leaves exactly one document, the reference's UniqueCols never visits it
(its loop runs while c1 < width - 1, :658) and reads a stale mask entry, usually dropping the last document and then indexing
past its hash array with width = 0 on the next round (heap overflow under ASan). Here a single remaining document is unique.
"""
import numpy as np


def _drop_columns(colptr, rows, counts, keep):
    lens = np.diff(colptr)
    entry_keep = np.repeat(keep, lens)
    new_colptr = np.concatenate([[0], np.cumsum(lens[keep])]).astype(np.int64)
    return new_colptr, rows[entry_keep], counts[entry_keep]


def _unique_mask(colptr, rows, counts):
    """True for the surviving column of every group of identical columns: the one with the largest index."""
    width = len(colptr) - 1
    groups = {}
    for c in range(width):
        s, e = colptr[c], colptr[c + 1]
        key = (rows[s:e].tobytes(), counts[s:e].tobytes())
        groups[key] = c                                      # later (larger) index wins
    keep = np.zeros(width, dtype=bool)
    keep[list(groups.values())] = True
    return keep


def preprocess_tf(m, n, colptr, rows, counts, max_iter=1000, docs_per_term=3, terms_per_doc=5):
    """Returns None if every column is pruned, else a dict with the pruned CSC (colptr, rows, counts), scores, the original
    index of every surviving term / document, and the new shape."""
    colptr = np.asarray(colptr, dtype=np.int64).copy()
    rows = np.asarray(rows, dtype=np.int64).copy()
    counts = np.asarray(counts, dtype=np.int64).copy()
    # SortRows (:120): rows ascending inside every column
    for c in range(n):
        s, e = colptr[c], colptr[c + 1]
        order = np.argsort(rows[s:e], kind="stable")
        rows[s:e] = rows[s:e][order]; counts[s:e] = counts[s:e][order]
    term_idx = np.arange(m, dtype=np.int64)
    doc_idx = np.arange(n, dtype=np.int64)
    height = m
    it = 0
    while it < max_iter:
        width = len(colptr) - 1
        # PruneRows
        hist = np.bincount(rows, weights=counts, minlength=height).astype(np.int64)
        hist_nz = np.bincount(rows, minlength=height)
        keep_r = (hist >= docs_per_term) & (hist_nz < width)
        if not keep_r.all():
            renum = np.cumsum(keep_r) - 1
            entry_keep = keep_r[rows]
            lens = np.diff(colptr)
            col_of = np.repeat(np.arange(width), lens)
            new_lens = np.bincount(col_of[entry_keep], minlength=width)
            colptr = np.concatenate([[0], np.cumsum(new_lens)]).astype(np.int64)
            rows = renum[rows[entry_keep]]; counts = counts[entry_keep]
            term_idx = term_idx[keep_r]
            height = int(keep_r.sum())
        # PrunableCols
        keep_c = np.diff(colptr) >= terms_per_doc
        new_width = int(keep_c.sum())
        if new_width == width:
            mask = _unique_mask(colptr, rows, counts)
            new_width = int(mask.sum())
            if new_width == width:
                break
        else:
            if new_width == 0:
                return None
            colptr, rows, counts = _drop_columns(colptr, rows, counts, keep_c)
            doc_idx = doc_idx[keep_c]
            width = new_width
            mask = _unique_mask(colptr, rows, counts)
            new_width = int(mask.sum())
        if width != new_width:
            colptr, rows, counts = _drop_columns(colptr, rows, counts, mask)
            doc_idx = doc_idx[mask]
        it += 1
    width = len(colptr) - 1
    hist_nz = np.bincount(rows, minlength=height)
    idf = np.log(float(width) / hist_nz.astype(np.float64))
    scores = (1.0 + np.log(counts.astype(np.float64))) * idf[rows]
    for c in range(width):                                   # the reference accumulates each column's sum of squares in order
        s, e = colptr[c], colptr[c + 1]
        ss = 0.0
        for v in scores[s:e]:
            ss += v * v
        scores[s:e] *= 1.0 / np.sqrt(ss)
    return {"m": height, "n": width, "colptr": colptr, "rows": rows, "counts": counts, "scores": scores,
            "term_indices": term_idx, "doc_indices": doc_idx}
