// smallk_b200 — the 2 x 2 non-negative least-squares step of Rank2 NMF, shared by the generic and the fused kernels.
//
// SystemSolveH / SystemSolveW (common/include/nmf_solver_rank2.hpp:81-131, :157-211): one Givens-like elimination of
// the 2 x 2 Gram matrix, applied to every right-hand side; OptimalActiveSetH / W (:218-318): where the unconstrained
// solution has a non-positive entry, the better of the two single-variable solutions.
#pragma once

#include <cfloat>

namespace smk {

struct Rank2Solve
{
    double t, b2, inv_a2, inv_d2, inv00, inv11, sq00, sq11;
    bool cosine, w_side, fail;

    // G is 2 x 2 column-major: G[0]=(0,0) G[1]=(1,0) G[2]=(0,1) G[3]=(1,1)
    __device__ __forceinline__ void init(const double* __restrict__ G, bool w)
    {
        const double a00 = G[0], a10 = G[1], a01 = G[2], a11 = G[3];
        const double eps = DBL_EPSILON;
        w_side = w;
        fail = (fabs(a00) < eps) && (fabs(a01) < eps);
        double a2 = 1.0, d2 = 1.0;
        t = 0.0; b2 = 0.0;
        cosine = fabs(a00) >= fabs(a01);
        if (!fail)
        {
            if (!w_side)
            {
                if (cosine) { t = -a10 / a00; a2 = a00 - t * a10; b2 = a01 - t * a11; d2 = a11 + t * a01; }
                else        { t = -a00 / a10; a2 = -a10 + t * a00; b2 = -a11 + t * a01; d2 = a01 + t * a11; }
            }
            else
            {
                if (cosine) { t = a01 / a00; a2 = a00 + t * a01; b2 = a10 + t * a11; d2 = a11 - t * a10; }
                else        { t = a00 / a01; a2 = -a01 - t * a00; b2 = -a11 - t * a10; d2 = a10 - t * a11; }
            }
            if (fabs(d2 / a2) < eps) fail = true;
        }
        inv_a2 = 1.0 / a2; inv_d2 = 1.0 / d2;
        inv00 = 1.0 / a00; inv11 = 1.0 / a11; sq00 = sqrt(a00); sq11 = sqrt(a11);
    }

    __device__ __forceinline__ double2 apply(const double b0, const double b1) const
    {
        double e2, f2;
        if (!w_side)
        {
            if (cosine) { e2 = b0 - t * b1;  f2 = b1 + t * b0; }
            else        { e2 = -b1 + t * b0; f2 = b0 + t * b1; }
        }
        else
        {
            if (cosine) { e2 = b0 + t * b1;  f2 = b1 - t * b0; }
            else        { e2 = -b1 - t * b0; f2 = b0 - t * b1; }
        }
        double x1 = f2 * inv_d2;
        double x0 = (e2 - b2 * x1) * inv_a2;
        if (x0 <= 0.0 || x1 <= 0.0)
        {
            double v1 = b0 * inv00, v2 = b1 * inv11;
            const double vv1 = v1 * sq00, vv2 = v2 * sq11;
            if (vv1 >= vv2) v2 = 0.0; else v1 = 0.0;
            x0 = v1; x1 = v2;
        }
        return make_double2(x0, x1);
    }
};

} // namespace smk
