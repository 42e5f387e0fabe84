"""Host-side partition arithmetic of the multi-GPU NMF, shared by bench.py, the multi-GPU tests and
mirrored by the C++ library (csrc/context.h: m_loc / w_row0 / w_rows; csrc/solver.cu: prod_HAt / gather_Wt).

* A and H are split by contiguous COLUMN blocks, one per rank; W is replicated.
* Per outer iteration the ranks exchange: H*H' (k x k, all-reduce), H*A' (k x m: reduce-scatter by row slice of W for
  BPP and MU, whose W update is row-separable; all-reduce for HALS and RANK2), the updated W slices (all-gather), the
  "some column was non-optimal" flag of each BPP NNLS call (1 int, max) and the two partial sums of the progress metric.
"""


def column_block(n, rank, world):
    """[c0, c1): this rank's columns of A and H."""
    return (n * rank) // world, (n * (rank + 1)) // world


def row_slice(m, rank, world):
    """(r0, rows, m_loc): this rank's rows of W in the row-sharded W update. Buffers are padded to m_loc * world rows so
    that reduce-scatter / all-gather move equal pieces; the last ranks may own fewer (or no) real rows."""
    m_loc = (m + world - 1) // world
    r0 = rank * m_loc
    rows = max(0, min(m_loc, m - r0))
    return r0, rows, m_loc
