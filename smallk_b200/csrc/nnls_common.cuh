// smallk_b200 — pieces shared by the two NNLS-BPP kernels (register fast path, shared-memory slow path).
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace smk {

constexpr double kZeroThresh = 1.0e-12;    // ZeroizeSmallValues threshold, nnls.hpp:215,226-227
constexpr int kPbar = 3;                   // nnls.hpp:153

// Pivoting state of one right-hand-side column. 24 bytes.
struct BppColState
{
    unsigned long long pm;   // passive set, bit r = row r
    int P, Ninf, round, col;
};

// BitMatrix::MaxRowIndex as the reference computes it, defect included
// (common/src/bit_matrix.cpp:432-472): a highest set bit found in a full 32-row word other than
// word 0 is reported 32 rows too low; an empty column reports 0.
__device__ __forceinline__ int max_row_index_ref(unsigned long long mask, int k)
{
    if (mask == 0ull) return 0;
    const int h = 63 - __clzll(static_cast<long long>(mask));
    const int full = k >> 5, extra = k & 31;
    const int w = h >> 5;
    if (extra > 0 && w == full) return h;
    return (w > 0) ? h - 32 : h;
}

// UpdatePassiveSet for one column (common/src/nnls.cpp:18-74): full exchange, P-counted full exchange,
// or the backup rule (toggle the largest "wrong" row). Every lane of the owning warp runs it on identical state; the
// `leader` lane counts the backup-rule firings in status[ST_BACKUP_COUNT] (a diagnostic the parity tests read).
__device__ __forceinline__ void update_passive_set(unsigned long long& pm, int& P, int& Ninf, int not_good,
                                                   unsigned long long nonopt, unsigned long long infeas, int k,
                                                   int* status, bool leader)
{
    if (not_good < Ninf)
    {
        P = kPbar; Ninf = not_good;
        pm = (pm | nonopt) & ~infeas;
    }
    else if (P >= 1)
    {
        P -= 1;
        pm = (pm | nonopt) & ~infeas;
    }
    else
    {
        const int ra = max_row_index_ref(nonopt, k), rb = max_row_index_ref(infeas, k);
        pm ^= (1ull << (ra > rb ? ra : rb));
        if (leader) atomicAdd(&status[ST_BACKUP_COUNT], 1);
    }
}

} // namespace smk
