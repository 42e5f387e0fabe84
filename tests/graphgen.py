"""Seeded synthetic inputs shared by the golden generators and the tests (hierclust / flatclust workloads)."""
import numpy as np
def powerlaw_graph(n, avg_deg, seed, exponent=2.5):
    """Chung-Lu style undirected power-law graph, no self loops, no isolated nodes; returns symmetric CSC pattern (values 1.0)."""
    rng = np.random.default_rng(seed)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** (-1.0 / (exponent - 1.0))
    p = w / w.sum()
    ne = int(n * avg_deg / 2)
    u = rng.choice(n, size=ne, p=p); v = rng.choice(n, size=ne, p=p)
    keep = u != v
    u, v = u[keep], v[keep]
    # attach isolated nodes to a random neighbour
    deg = np.bincount(np.concatenate([u, v]), minlength=n)
    iso = np.nonzero(deg == 0)[0]
    if len(iso):
        t = rng.choice(n, size=len(iso), p=p)
        t = np.where(t == iso, (t + 1) % n, t)
        u = np.concatenate([u, iso]); v = np.concatenate([v, t])
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    e = np.unique(lo.astype(np.int64) * n + hi)
    lo, hi = e // n, e % n
    rows = np.concatenate([lo, hi]); cols = np.concatenate([hi, lo])
    order = np.lexsort((rows, cols))
    rows, cols = rows[order], cols[order]
    colp = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n))]).astype(np.uint32)
    return colp, rows.astype(np.uint32), np.ones(len(rows))


def topic_matrix(m, n, topics, seed, terms_per_doc=25):
    """Term-document matrix with planted topics: each document draws its terms mostly from one topic's vocabulary
    block; values are tf-idf-like positives. Returns scipy CSC with sorted indices and no empty rows or columns."""
    import scipy.sparse as sps
    rng = np.random.default_rng(seed)
    block = m // topics
    rows, cols, vals = [], [], []
    for j in range(n):
        t = j % topics
        own = rng.integers(t * block, (t + 1) * block, size=int(terms_per_doc * 0.8))
        other = rng.integers(0, m, size=terms_per_doc - len(own))
        r = np.unique(np.concatenate([own, other]))
        rows.append(r); cols.append(np.full(len(r), j)); vals.append(0.2 + rng.random(len(r)))
    rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
    missing = np.setdiff1d(np.arange(m), rows)            # no zero rows (clust_hier_generic.hpp:87-88)
    if len(missing):
        rows = np.concatenate([rows, missing]); cols = np.concatenate([cols, rng.integers(0, n, size=len(missing))])
        vals = np.concatenate([vals, 0.2 + rng.random(len(missing))])
    S = sps.csc_matrix((vals, (rows, cols)), shape=(m, n))
    S.sum_duplicates(); S.sort_indices()
    return S


def community_graph(n, edges, communities, seed, p_in=0.85, exponent=2.5):
    """Degree-corrected planted-partition graph: power-law expected degrees, `communities` groups of power-law sizes, a
    fraction p_in of the edges inside a group. Undirected, no self loops, no isolated nodes; returns the full symmetric
    CSC pattern (values 1.0, rows ascending) — what LoadMatrixMarketFile makes of a `pattern symmetric` file."""
    rng = np.random.default_rng(seed)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** (-1.0 / (exponent - 1.0))
    w = w[rng.permutation(n)]
    sizes = (np.arange(1, communities + 1, dtype=np.float64)) ** (-0.7)
    comm = rng.choice(communities, size=n, p=sizes / sizes.sum())
    order = np.argsort(comm, kind="stable")
    start = np.searchsorted(comm[order], np.arange(communities + 1))
    cw = np.concatenate([[0.0], np.cumsum(w[order])])          # cumulative weights in community order
    tot = cw[-1]
    u = order[np.minimum(np.searchsorted(cw, rng.random(edges) * tot, side="right") - 1, n - 1)]
    inside = rng.random(edges) < p_in
    cu = comm[u]
    lo, hi = cw[start[cu]], cw[start[cu + 1]]
    r = rng.random(edges)
    target = np.where(inside, lo + r * (hi - lo), r * tot)
    v = order[np.minimum(np.searchsorted(cw, target, side="right") - 1, n - 1)]
    keep = u != v
    u, v = u[keep], v[keep]
    deg = np.bincount(np.concatenate([u, v]), minlength=n)
    iso = np.nonzero(deg == 0)[0]
    if len(iso):
        # attach isolated nodes to a member of their own community
        ci = comm[iso]
        t = order[np.minimum(start[ci] + (rng.random(len(iso)) * (start[ci + 1] - start[ci])).astype(np.int64), n - 1)]
        t = np.where(t == iso, (t + 1) % n, t)
        u = np.concatenate([u, iso]); v = np.concatenate([v, t])
    a, b = np.minimum(u, v).astype(np.int64), np.maximum(u, v).astype(np.int64)
    e = np.unique(a * n + b)
    a, b = e // n, e % n
    rows = np.concatenate([a, b]); cols = np.concatenate([b, a])
    o = np.lexsort((rows, cols))
    rows, cols = rows[o], cols[o]
    colp = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n))]).astype(np.uint32)
    return colp, rows.astype(np.uint32), np.ones(len(rows))
