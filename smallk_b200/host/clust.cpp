// smallk_b200 host — HierNMF2 driver: Clust / ClustSparse of the reference (hierclust/src/clust.cpp:108-203) and
// the generic tree-growing code behind them (hierclust/include/clust_hier_generic.hpp:77-517), re-organised
// around a matrix that lives on the GPU:
//   * the input matrix is uploaded once (smk_load_csc / smk_load_dense);
//   * each node's column subset is extracted and row-compacted ON THE DEVICE (smk_select_columns), only the
//     new->old row map comes back;
//   * every rank-2 factorization is one smk_nmf call on the active subset;
//   * labels, the (sparse) scatter of W back to m rows, priorities and tree updates are host work on k = 2 data.
#include "clust.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <future>
#include <iostream>
#include <limits>
#include <mutex>
#include <sstream>
#include <stdexcept>

#include "host_internal.hpp"
#include "matrix_io.hpp"

using std::cerr;
using std::cout;
using std::endl;

// --------------------------------------------------------------------------------------------------------------
// IsValid(ClustOptions): hierclust/src/clust_options.cpp:24-110
// --------------------------------------------------------------------------------------------------------------
bool IsValid(const ClustOptions& opts, bool validate_matrix)
{
    if (validate_matrix)
    {
        if (opts.nmf_opts.height <= 0) { cerr << "error: matrix height must be a positive integer" << endl; return false; }
        if (opts.nmf_opts.width <= 0) { cerr << "error: matrix width must be a positive integer" << endl; return false; }
        if (opts.nmf_opts.k <= 0) { cerr << "error: cluster count must be a positive integer" << endl; return false; }
        if (opts.nmf_opts.k > opts.nmf_opts.width) { cerr << "error: k value cannot exceed the matrix width" << endl; return false; }
    }
    if (opts.num_clusters <= 1) { cerr << "error: number of clusters must be >= 2" << endl; return false; }
    if (opts.nmf_opts.tol <= 0.0 || opts.nmf_opts.tol >= 1.0) { cerr << "error: tolerance must be in the interval (0.0, 1.0)" << endl; return false; }
    if (opts.nmf_opts.min_iter <= 0) { cerr << "error: miniter must be a positive integer" << endl; return false; }
    if (opts.nmf_opts.max_iter <= 0) { cerr << "error: maxiter must be a positive integer" << endl; return false; }
    if (opts.maxterms <= 0) { cerr << "error: maxterms must be a positive integer" << endl; return false; }
    if (opts.trial_allowance < 0) { cerr << "error: trial_allowance for hierarchical clustering is negative" << endl; return false; }
    if (opts.unbalanced < 0.0 || opts.unbalanced >= 1.0) { cerr << "error: the unbalanced value should be in the interval [0, 1)" << endl; return false; }
    if (NmfProgressAlgorithm::PG_RATIO != opts.nmf_opts.prog_est_algorithm && NmfProgressAlgorithm::DELTA_FNORM != opts.nmf_opts.prog_est_algorithm)
    { cerr << "error: unknown stopping criterion " << endl; return false; }
    return true;
}

// --------------------------------------------------------------------------------------------------------------
// priority score
// --------------------------------------------------------------------------------------------------------------
R compute_priority_on(smk_ctx* ctx, const R* W_parent, const R* W_child, int n);
R compute_priority_plain(smk_ctx* ctx, const R* W_parent, const R* W_child, int n);
R compute_priority_rows(smk_ctx* ctx, const R* W_parent, const R* W_child, int n, const unsigned int* child_rows, int n_child_rows,
                        const unsigned int* parent_rows = nullptr, int n_parent_rows = 0);

namespace {

// Row indices in decreasing order of v, ties by ascending row (desc_ordered, clust_hier_util.hpp:46-57). The order is
// total, so any correct sort reproduces it; factor columns are mostly exact zeros below the root, so the rows are
// split into positive / zero / negative groups and only the two outer groups are sorted.
// Sorts of more than this many keys go to the GPU (smk_argsort_desc / smk_sort_desc: one stable radix sort instead of
// an O(n log n) comparison sort on one host core); shorter ones are not worth the round trip.
const int kDeviceSortMin = 8192;

void desc_order(const R* v, const int n, std::vector<int>& order, smk_ctx* ctx)
{
    order.resize(n);
    int npos = 0, nzero = 0;
    for (int i = 0; i < n; ++i) { if (v[i] > 0) ++npos; else if (v[i] == 0) ++nzero; }
    int ip = 0, iz = npos, in = npos + nzero;
    for (int i = 0; i < n; ++i)
    {
        if (v[i] > 0) order[ip++] = i;
        else if (v[i] == 0) order[iz++] = i;
        else order[in++] = i;                       // negative or NaN (NaN never occurs in a successful factorization)
    }
    auto cmp = [v](int a, int b) { return v[a] > v[b] || (v[a] == v[b] && a < b); };
    if (ctx && npos >= kDeviceSortMin)
    {
        // positives in index order -> stable descending sort keeps the index tie-break
        std::vector<R> keys(npos);
        std::vector<int> perm(npos);
        for (int i = 0; i < npos; ++i) keys[i] = v[order[i]];
        if (smk_argsort_desc(ctx, keys.data(), npos, perm.data()) != SMK_OK) throw std::runtime_error(smk_last_error(ctx));
        std::vector<int> sorted(npos);
        for (int i = 0; i < npos; ++i) sorted[i] = order[perm[i]];
        std::copy(sorted.begin(), sorted.end(), order.begin());
    }
    else std::sort(order.begin(), order.begin() + npos, cmp);
    std::sort(order.begin() + npos + nzero, order.end(), cmp);
}

// log(i) and log2(i) of small integers are needed ~4m times per split; tabulated once per matrix height
struct LogTables
{
    std::vector<double> ln, lg2;
    void ensure(const int n)
    {
        if (static_cast<int>(ln.size()) > n + 1) return;
        const int old = static_cast<int>(ln.size());
        ln.resize(n + 2); lg2.resize(n + 2);
        for (int i = old; i < n + 2; ++i) { ln[i] = std::log(static_cast<double>(i)); lg2[i] = std::log2(static_cast<double>(i)); }
    }
};
LogTables g_logs;

// NDCG_part (clust_hier_util.hpp:61-100) numerator: positions i of `test` in order, gain = weight_part at the
// parent rank of row test[i], discounted by log2(i+1) for i > 0, accumulated left to right.
R dcg_of(const std::vector<int>& test, const std::vector<int>& rank_parent, const std::vector<R>& weight_part)
{
    const int n = static_cast<int>(test.size());
    R cum = weight_part[rank_parent[test[0]]];
    for (int i = 1; i < n; ++i)
    {
        const R g = weight_part[rank_parent[test[i]]];
        cum = cum + g / g_logs.lg2[i + 1];
    }
    return cum;
}

} // namespace

R compute_priority(const R* W_parent, const R* W_child, const int n) { return compute_priority_on(nullptr, W_parent, W_child, n); }

// The straightforward evaluation: every step of clust_hier_util.hpp:105-173 on full-length arrays. Used when a factor
// has a negative or NaN entry (never after a successful factorization) and as the yardstick of the streaming version.
R compute_priority_plain(smk_ctx* ctx, const R* W_parent, const R* W_child, const int n)
{
    std::vector<int> ord_p, ord_1, ord_2;
    int n_part = 0;
    for (int i = 0; i < n; ++i) if (W_parent[i] != 0) ++n_part;
    if (n_part <= 1) return R(-3);
    desc_order(W_parent, n, ord_p, ctx);
    desc_order(W_child, n, ord_1, ctx);
    desc_order(W_child + n, n, ord_2, ctx);
    g_logs.ensure(n);
    const std::vector<double>& ln = g_logs.ln;

    std::vector<R> weight(n), weight_part(n);
    for (int i = 0; i < n; ++i) weight[i] = ln[n - i];
    int first_zero = -1;
    for (int i = 0; i < n; ++i) if (W_parent[ord_p[i]] == 0) { first_zero = i; break; }
    if (first_zero > -1) for (int i = first_zero; i < n; ++i) weight[i] = 1;
    for (int i = 0; i < n_part; ++i) weight_part[i] = ln[n_part - i];
    for (int i = n_part; i < n; ++i) weight_part[i] = 0;

    // rank of every row in the two child orderings; the worse of the two discounts the row
    std::vector<int> rank_1(n), rank_2(n), rank_p(n);
    for (int p = 0; p < n; ++p) { rank_1[ord_1[p]] = p; rank_2[ord_2[p]] = p; rank_p[ord_p[p]] = p; }
    for (int i = 0; i < n; ++i)
    {
        const int row = ord_p[i];
        const int worst = rank_1[row] < rank_2[row] ? rank_2[row] : rank_1[row];
        R discount = ln[n - worst];
        if (discount == 0) discount = ln[2];
        weight[i] = weight[i] / discount;
        weight_part[i] = weight_part[i] / discount;
    }

    const R dcg1 = dcg_of(ord_1, rank_p, weight_part);
    const R dcg2 = dcg_of(ord_2, rank_p, weight_part);
    // ideal score: the weights in decreasing order with the same discounting (identical for both children)
    if (ctx && n >= kDeviceSortMin)
    {
        if (smk_sort_desc(ctx, weight.data(), n) != SMK_OK) throw std::runtime_error(smk_last_error(ctx));
    }
    else std::sort(weight.begin(), weight.end(), std::greater<R>());
    R ideal = weight[0];
    for (int i = 1; i < n; ++i) ideal = ideal + weight[i] / g_logs.lg2[i + 1];
    return (dcg1 / ideal) * (dcg2 / ideal);
}

// --------------------------------------------------------------------------------------------------------------
// The same score without the full-length random-access passes. Below the root a node's factors are zero in all
// but the node's own rows, while every vector is m long, so the plain evaluation spends its time ranking, gathering
// and sorting zeros. What the score needs, and how each piece is obtained with the reference's exact arithmetic:
//   * desc_ordered of a non-negative vector = its positive entries sorted (ties by row), then its zero rows in row
//     order: a zero row's rank is (number of positives) + (number of zero rows before it), a running count;
//   * the DCG sums add weight_part[rank_parent[row]] / log2(position + 1) over the positions of a child ordering, and
//     weight_part is exactly 0 beyond the parent's positive rows: cum + 0 == cum, so only parent-positive rows are
//     visited, in increasing position — child-positive rows in sorted order, then child-zero rows in row order;
//   * the ideal sum needs the weights sorted in decreasing order. A row that is zero in the parent and in both
//     children has weight 1 / log(n - max(rank_1, rank_2)), both ranks running counts: these weights are already
//     non-decreasing in the row index, so they are merged, from the last row down, with the sorted weights of the
//     other rows. Equal weights are interchangeable in the sum, so the merge equals std::sort + loop bit for bit.
// --------------------------------------------------------------------------------------------------------------
namespace {

// the three ranks of a row side by side (one cache line per row where three arrays were three): rank among the parent's positive
// entries (-1: the parent entry is not positive), rank in the ordering of child 1 / child 2
struct RowRanks { int p, c1, c2, pad; };

struct PriorityScratch
{
    std::vector<RowRanks> rr;
    std::vector<int> pos_p, pos_1, pos_2, pz_1, pz_2, other, rank_p, rank_1, rank_2, perm, tmp_i;
    std::vector<R> keys, wfull, wpart, wsort;
    std::vector<std::pair<R, int>> pairs;
    std::vector<unsigned long long> zero_1, zero_2, zero_all;   // one bit per row: zero in child 1 / child 2 / everywhere
};
PriorityScratch g_ps;
// g_ps / g_rs belong to whoever holds this: every evaluation on the calling thread, and the tree driver's worker thread
// if it ever has to leave its own scratch for the general evaluation (compute_priority_on)
std::mutex g_priority_mutex;

// rows[0..cnt) (ascending row order) sorted by decreasing v[row], ties by ascending row
void sort_rows_desc(const R* v, std::vector<int>& rows, smk_ctx* ctx, PriorityScratch& S = g_ps)
{
    const int cnt = static_cast<int>(rows.size());
    if (ctx && cnt >= kDeviceSortMin)
    {
        S.keys.resize(cnt); S.perm.resize(cnt); S.tmp_i.resize(cnt);
        for (int i = 0; i < cnt; ++i) S.keys[i] = v[rows[i]];
        if (smk_argsort_desc(ctx, S.keys.data(), cnt, S.perm.data()) != SMK_OK) throw std::runtime_error(smk_last_error(ctx));
        for (int i = 0; i < cnt; ++i) S.tmp_i[i] = rows[S.perm[i]];
        rows.swap(S.tmp_i);
    }
    else
    {
        // (value, row) pairs sort on contiguous keys; the order is the same total order
        std::vector<std::pair<R, int>>& pr = S.pairs;
        pr.resize(cnt);
        for (int i = 0; i < cnt; ++i) pr[i] = std::make_pair(v[rows[i]], rows[i]);
        std::sort(pr.begin(), pr.end(), [](const std::pair<R, int>& a, const std::pair<R, int>& b) {
            return a.first > b.first || (a.first == b.first && a.second < b.second);
        });
        for (int i = 0; i < cnt; ++i) rows[i] = pr[i].second;
    }
}

} // namespace

R compute_priority_on(smk_ctx* ctx, const R* W_parent, const R* W_child, const int n)
{
    const R* P = W_parent; const R* C1 = W_child; const R* C2 = W_child + n;
    PriorityScratch& S = g_ps;
    S.pos_p.clear(); S.pos_1.clear(); S.pos_2.clear(); S.pz_1.clear(); S.pz_2.clear(); S.other.clear();
    if (static_cast<int>(S.rank_p.size()) < n) { S.rank_p.resize(n); S.rank_1.resize(n); S.rank_2.resize(n); }
    const int nwords = (n + 63) / 64;
    S.zero_1.assign(nwords, 0ull); S.zero_2.assign(nwords, 0ull); S.zero_all.assign(nwords, 0ull);
    // one pass over the rows: positive rows of each vector, the zero-row counts that will be looked up (the number of
    // positives is added once it is known), and three bit masks for the reverse pass of the ideal sum. Anything but
    // non-negative finite entries goes the plain way.
    int z1 = 0, z2 = 0;
    bool regular = true;
    for (int i = 0; i < n; ++i)
    {
        const R pv = P[i], av = C1[i], bv = C2[i];
        if (!(pv >= 0) || !(av >= 0) || !(bv >= 0)) { regular = false; break; }
        const bool p = pv > 0, a = av > 0, b = bv > 0;
        const unsigned long long bit = 1ull << (i & 63);
        if (p) S.pos_p.push_back(i);
        if (a) S.pos_1.push_back(i); else { if (p || b) S.rank_1[i] = z1; if (p) S.pz_1.push_back(i); ++z1; S.zero_1[i >> 6] |= bit; }
        if (b) S.pos_2.push_back(i); else { if (p || a) S.rank_2[i] = z2; if (p) S.pz_2.push_back(i); ++z2; S.zero_2[i >> 6] |= bit; }
        if (!p) { if (a || b) S.other.push_back(i); else S.zero_all[i >> 6] |= bit; }
    }
    if (!regular) return compute_priority_plain(ctx, W_parent, W_child, n);
    const int np = static_cast<int>(S.pos_p.size()), n1 = static_cast<int>(S.pos_1.size()), n2 = static_cast<int>(S.pos_2.size());
    const int n_part = np;
    if (n_part <= 1) return R(-3);
    g_logs.ensure(n);
    const std::vector<double>& ln = g_logs.ln;
    const std::vector<double>& lg2 = g_logs.lg2;
    auto discount_of = [&](const int worst) { const R d = ln[n - worst]; return d == 0 ? ln[2] : d; };
    // zero rows rank after the positives
    for (const int row : S.pz_1) S.rank_1[row] += n1;
    for (const int row : S.pz_2) S.rank_2[row] += n2;
    for (const int row : S.other)
    {
        if (!(C1[row] > 0)) S.rank_1[row] += n1;
        if (!(C2[row] > 0)) S.rank_2[row] += n2;
    }
    sort_rows_desc(P, S.pos_p, ctx);
    sort_rows_desc(C1, S.pos_1, ctx);
    sort_rows_desc(C2, S.pos_2, ctx);
    for (int q = 0; q < np; ++q) S.rank_p[S.pos_p[q]] = q;
    for (int q = 0; q < n1; ++q) S.rank_1[S.pos_1[q]] = q;
    for (int q = 0; q < n2; ++q) S.rank_2[S.pos_2[q]] = q;

    // weights of the parent-positive rows (positions 0..np-1 of the parent ordering), then of the rows that are zero in
    // the parent but positive in a child (weight 1 / discount, like every position from first_zero on)
    S.wfull.resize(static_cast<size_t>(np) + S.other.size()); S.wpart.resize(np);
    for (int i = 0; i < np; ++i)
    {
        const int row = S.pos_p[i];
        const R d = discount_of(std::max(S.rank_1[row], S.rank_2[row]));
        S.wfull[i] = ln[n - i] / d;
        S.wpart[i] = ln[n_part - i] / d;
    }
    for (size_t u = 0; u < S.other.size(); ++u)
    {
        const int row = S.other[u];
        S.wfull[np + u] = R(1) / discount_of(std::max(S.rank_1[row], S.rank_2[row]));
    }

    // DCG of a child ordering: its positive rows in sorted order, then the parent-positive rows among its zero rows
    auto dcg = [&](const std::vector<int>& pos, const std::vector<int>& pz, const std::vector<int>& rank_c, const R* Pv) {
        R cum = 0;
        const int cnt = static_cast<int>(pos.size());
        for (int q = 0; q < cnt; ++q)
        {
            const int row = pos[q];
            if (!(Pv[row] > 0)) continue;                    // weight_part is 0 there
            const R g = S.wpart[S.rank_p[row]];
            cum = q == 0 ? g : cum + g / lg2[q + 1];
        }
        for (const int row : pz)
        {
            const int q = rank_c[row];
            const R g = S.wpart[S.rank_p[row]];
            cum = q == 0 ? g : cum + g / lg2[q + 1];
        }
        return cum;
    };
    const R dcg1 = dcg(S.pos_1, S.pz_1, S.rank_1, P);
    const R dcg2 = dcg(S.pos_2, S.pz_2, S.rank_2, P);

    // ideal score: sorted irregular weights merged with the all-zero rows' weights, last row first
    const int nw = static_cast<int>(S.wfull.size());
    if (ctx && nw >= kDeviceSortMin)
    {
        if (smk_sort_desc(ctx, S.wfull.data(), nw) != SMK_OK) throw std::runtime_error(smk_last_error(ctx));
    }
    else std::sort(S.wfull.begin(), S.wfull.end(), std::greater<R>());
    R ideal = 0;
    int pos = 0, head = 0;
    auto take = [&](const R w) { ideal = pos == 0 ? w : ideal + w / lg2[pos + 1]; ++pos; };
    int zb1 = n - n1, zb2 = n - n2;                              // zero rows of each child not yet passed, counting from the end
    for (int wd = nwords - 1; wd >= 0; --wd)
    {
        const unsigned long long m1 = S.zero_1[wd], m2 = S.zero_2[wd], ma = S.zero_all[wd];
        if (ma == 0) { zb1 -= __builtin_popcountll(m1); zb2 -= __builtin_popcountll(m2); continue; }
        const int top = (wd == nwords - 1) ? ((n - 1) & 63) : 63;
        for (int bpos = top; bpos >= 0; --bpos)
        {
            const unsigned long long bit = 1ull << bpos;
            if (m1 & bit) --zb1;
            if (m2 & bit) --zb2;
            if (!(ma & bit)) continue;
            const R w = R(1) / discount_of(std::max(n1 + zb1, n2 + zb2));
            while (head < nw && S.wfull[head] >= w) take(S.wfull[head++]);
            take(w);
        }
    }
    while (head < nw) take(S.wfull[head++]);
    return (dcg1 / ideal) * (dcg2 / ideal);
}

// --------------------------------------------------------------------------------------------------------------
// The same score again, for the tree driver, which KNOWS where the child factors can be non-zero: child_rows (ascending) is
// the row map of the node's compacted matrix. compute_priority_on still reads all 3 m entries and walks m bits per call
// (5-6 ms at m = 320 000, whatever the node's size: 1.1 s of a 2.7 s 64-leaf run); here
//   * the rows that can matter are U = (non-zero rows of the parent vector) u child_rows: one lean scan of the parent vector,
//     everything else in O(|U|): a zero row's rank in a child ordering is (positives of that child) + row - (positives of
//     that child before the row), a running count over U;
//   * the ideal sum still has one term per row (the reference adds m terms one after the other, and so must this), but
//     between two rows of U the all-zero rows form a run whose weights are 1 / discount(row + constant): a table of
//     1 / discount values (one division per table entry, once per matrix height) turns the run into lookup, one division by
//     log2(position + 1), one addition per row, in the reference's order.
// Same operations on the same operands in the same order as compute_priority_on: bit-identical (tests/test_priority.py).
// --------------------------------------------------------------------------------------------------------------
namespace {

struct InvDiscountTable
{
    std::vector<double> inv;        // inv[x] = 1 / (ln(x) == 0 ? ln(2) : ln(x))
    void ensure(const int n)
    {
        const int old = static_cast<int>(inv.size());
        if (old > n + 1) return;
        g_logs.ensure(n);
        inv.resize(n + 2);
        for (int x = old; x < n + 2; ++x) { const double d = g_logs.ln[x]; inv[x] = 1.0 / (d == 0 ? g_logs.ln[2] : d); }
    }
};
InvDiscountTable g_invd;

struct RowsScratch
{
    std::vector<int> pnz, u, a1, a2;
    std::vector<unsigned char> allzero;
};
RowsScratch g_rs;

} // namespace

namespace {
// SMK_PRIORITY_PROF=1: seconds per section of compute_priority_rows, summed over the calls, on stderr at process exit
struct PriorityProf
{
    bool on = false;
    double t[6] = {0, 0, 0, 0, 0, 0};
    long long calls = 0, rows = 0;
    PriorityProf() { const char* e = getenv("SMK_PRIORITY_PROF"); on = e && atoi(e) != 0; }
    ~PriorityProf()
    {
        if (on) fprintf(stderr, "compute_priority_rows: %lld calls, %lld rows of U; scan+merge %.3f s, classify %.3f s, three sorts %.3f s, weights+dcg %.3f s, "
                                "weight sort %.3f s, ideal sum %.3f s\n", calls, rows, t[0], t[1], t[2], t[3], t[4], t[5]);
    }
};
PriorityProf g_pprof;
struct Lap
{
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit Lap(bool timed) : on(timed && g_pprof.on), t0(std::chrono::steady_clock::now()) {}
    void mark(int i) { if (!on) return; const auto t1 = std::chrono::steady_clock::now(); g_pprof.t[i] += std::chrono::duration<double>(t1 - t0).count(); t0 = t1; }
};
} // namespace

namespace {
// acc = (...((acc + a[0] / b[0]) + a[1] / b[1]) + ...) + a[c-1] / b[c-1]: the quotients of a block are formed first (independent
// IEEE divisions, two or four per instruction: the same values a scalar division gives), then added one after the other in
// the order the reference adds them. The ideal-score sum of compute_priority has one such term per row of the matrix; done one
// term at a time it is bound by the throughput of the scalar divider.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx"))) void quotients_avx(const double* a, const double* b, const int cnt, double* q)
{
    int j = 0;
    for (; j + 4 <= cnt; j += 4) _mm256_storeu_pd(q + j, _mm256_div_pd(_mm256_loadu_pd(a + j), _mm256_loadu_pd(b + j)));
    for (; j < cnt; ++j) q[j] = a[j] / b[j];
}
void quotients_sse2(const double* a, const double* b, const int cnt, double* q)
{
    int j = 0;
    for (; j + 2 <= cnt; j += 2) _mm_storeu_pd(q + j, _mm_div_pd(_mm_loadu_pd(a + j), _mm_loadu_pd(b + j)));
    for (; j < cnt; ++j) q[j] = a[j] / b[j];
}
typedef void (*QuotientFn)(const double*, const double*, int, double*);
const QuotientFn g_quotients = __builtin_cpu_supports("avx") ? quotients_avx : quotients_sse2;
#else
void quotients_plain(const double* a, const double* b, const int cnt, double* q) { for (int j = 0; j < cnt; ++j) q[j] = a[j] / b[j]; }
typedef void (*QuotientFn)(const double*, const double*, int, double*);
const QuotientFn g_quotients = quotients_plain;
#endif
inline R add_quotients(R acc, const double* a, const double* b, const int c)
{
    double q[32];
    for (int j0 = 0; j0 < c; j0 += 32)
    {
        const int cnt = std::min(32, c - j0);
        g_quotients(a + j0, b + j0, cnt, q);
        for (int j = 0; j < cnt; ++j) acc = acc + q[j];
    }
    return acc;
}

// The evaluation proper, on the scratch it is given: g_ps / g_rs (the caller holds g_priority_mutex) or the tree driver's
// worker-thread set. `timed` = this call feeds the SMK_PRIORITY_PROF section timers (the calling thread's evaluations only).
// parent_rows (optional, ascending): the rows outside of which the parent vector is zero (the row map of the node the vector
// was factored on) — the scan for its non-zero rows then visits those instead of all n.
R priority_rows_impl(smk_ctx* ctx, PriorityScratch& S, RowsScratch& Q, const bool timed, const R* W_parent, const R* W_child, const int n,
                     const unsigned int* child_rows, const int n_child_rows, const unsigned int* parent_rows = nullptr, const int n_parent_rows = 0)
{
    Lap lap(timed);
    const R* P = W_parent; const R* C1 = W_child; const R* C2 = W_child + n;
    S.pos_p.clear(); S.pos_1.clear(); S.pos_2.clear(); S.pz_1.clear(); S.pz_2.clear(); S.other.clear();
    if (static_cast<int>(S.rr.size()) < n) S.rr.resize(n);
    RowRanks* const rr = S.rr.data();
    // U = non-zero rows of the parent vector, merged with the child's rows
    Q.pnz.clear();
    if (parent_rows) { for (int t = 0; t < n_parent_rows; ++t) { const int i = static_cast<int>(parent_rows[t]); if (P[i] != 0) Q.pnz.push_back(i); } }
    else for (int i = 0; i < n; ++i) if (P[i] != 0) Q.pnz.push_back(i);
    Q.u.clear();
    {
        size_t a = 0; int b = 0;
        const size_t na = Q.pnz.size();
        while (a < na || b < n_child_rows)
        {
            const int ra = a < na ? Q.pnz[a] : n, rb = b < n_child_rows ? static_cast<int>(child_rows[b]) : n;
            if (ra < rb) { Q.u.push_back(ra); ++a; }
            else if (rb < ra) { Q.u.push_back(rb); ++b; }
            else { Q.u.push_back(ra); ++a; ++b; }
        }
    }
    const int nu = static_cast<int>(Q.u.size());
    lap.mark(0);
    if (timed && g_pprof.on) { g_pprof.calls += 1; g_pprof.rows += nu; }
    Q.a1.resize(nu); Q.a2.resize(nu); Q.allzero.resize(nu);
    // one pass over U: the lists compute_priority_on builds from all m rows, with the zero-row counts in closed form
    int c1 = 0, c2 = 0;                     // positive rows of child 1 / child 2 seen so far
    bool regular = true;
    for (int t = 0; t < nu; ++t)
    {
        const int i = Q.u[t];
        const R pv = P[i], av = C1[i], bv = C2[i];
        if (!(pv >= 0) || !(av >= 0) || !(bv >= 0)) { regular = false; break; }
        const bool p = pv > 0, a = av > 0, b = bv > 0;
        if (p) S.pos_p.push_back(i);
        if (a) S.pos_1.push_back(i); else { if (p || b) rr[i].c1 = i - c1; if (p) S.pz_1.push_back(i); }
        if (b) S.pos_2.push_back(i); else { if (p || a) rr[i].c2 = i - c2; if (p) S.pz_2.push_back(i); }
        if (!p && (a || b)) { S.other.push_back(i); rr[i].p = -1; }
        Q.allzero[t] = (!p && !a && !b) ? 1 : 0;
        if (a) ++c1;
        if (b) ++c2;
        Q.a1[t] = c1; Q.a2[t] = c2;
    }
    // rows outside child_rows must be zero in the child factors (the driver scatters into a zeroed buffer); anything irregular
    // goes the general way
    if (!regular)
    {
        if (&S == &g_ps) return compute_priority_on(ctx, W_parent, W_child, n);        // the caller holds the mutex
        std::lock_guard<std::mutex> lock(g_priority_mutex);
        return compute_priority_on(ctx, W_parent, W_child, n);
    }
    const int np = static_cast<int>(S.pos_p.size()), n1 = static_cast<int>(S.pos_1.size()), n2 = static_cast<int>(S.pos_2.size());
    const int n_part = np;
    if (n_part <= 1) return R(-3);
    g_logs.ensure(n);
    g_invd.ensure(n);
    const std::vector<double>& ln = g_logs.ln;
    const std::vector<double>& lg2 = g_logs.lg2;
    const double* invd = g_invd.inv.data();
    auto discount_of = [&](const int worst) { const R d = ln[n - worst]; return d == 0 ? ln[2] : d; };
    for (const int row : S.pz_1) rr[row].c1 += n1;
    for (const int row : S.pz_2) rr[row].c2 += n2;
    for (const int row : S.other)
    {
        if (!(C1[row] > 0)) rr[row].c1 += n1;
        if (!(C2[row] > 0)) rr[row].c2 += n2;
    }
    lap.mark(1);
    sort_rows_desc(P, S.pos_p, ctx, S);
    sort_rows_desc(C1, S.pos_1, ctx, S);
    sort_rows_desc(C2, S.pos_2, ctx, S);
    lap.mark(2);
    for (int q = 0; q < np; ++q) rr[S.pos_p[q]].p = q;
    for (int q = 0; q < n1; ++q) rr[S.pos_1[q]].c1 = q;
    for (int q = 0; q < n2; ++q) rr[S.pos_2[q]].c2 = q;

    S.wfull.resize(static_cast<size_t>(np) + S.other.size()); S.wpart.resize(np);
    for (int i = 0; i < np; ++i)
    {
        const int row = S.pos_p[i];
        const R d = discount_of(std::max(rr[row].c1, rr[row].c2));
        S.wfull[i] = ln[n - i] / d;
        S.wpart[i] = ln[n_part - i] / d;
    }
    for (size_t u = 0; u < S.other.size(); ++u)
    {
        const int row = S.other[u];
        S.wfull[np + u] = R(1) / discount_of(std::max(rr[row].c1, rr[row].c2));
    }
    // a row of a child's ordering counts where the parent entry is positive (p >= 0: "other" rows were marked -1 above)
    auto dcg = [&](const std::vector<int>& pos, const std::vector<int>& pz, int RowRanks::*rank_c) {
        R cum = 0;
        const int cnt = static_cast<int>(pos.size());
        for (int q = 0; q < cnt; ++q)
        {
            const int rp = rr[pos[q]].p;
            if (rp < 0) continue;
            const R g = S.wpart[rp];
            cum = q == 0 ? g : cum + g / lg2[q + 1];
        }
        for (const int row : pz)
        {
            const RowRanks& x = rr[row];
            const int q = x.*rank_c;
            const R g = S.wpart[x.p];
            cum = q == 0 ? g : cum + g / lg2[q + 1];
        }
        return cum;
    };
    const R dcg1 = dcg(S.pos_1, S.pz_1, &RowRanks::c1);
    const R dcg2 = dcg(S.pos_2, S.pz_2, &RowRanks::c2);

    lap.mark(3);
    // ideal score: sorted irregular weights merged with the all-zero rows' weights, last row first
    const int nw = static_cast<int>(S.wfull.size());
    if (ctx && nw >= kDeviceSortMin)
    {
        if (smk_sort_desc(ctx, S.wfull.data(), nw) != SMK_OK) throw std::runtime_error(smk_last_error(ctx));
    }
    else std::sort(S.wfull.begin(), S.wfull.end(), std::greater<R>());
    lap.mark(4);
    const R* wf = S.wfull.data();
    const double* l2 = lg2.data();
    R ideal = 0;
    int pos = 0, head = 0;
    // all-zero rows hi, hi - 1, ..., lo with `shift` = max(n1 - positives of child 1 before them, n2 - ... of child 2):
    // rank of row i in the worse child ordering = i + shift
    // Row i of the run has weight invd[n - (i + shift)]: walking the run (i falling) walks invd upwards, and invd does not
    // increase with its index. Where at least kStretch rows come before the next irregular weight wf[head] (those with a strictly
    // larger weight; an equal irregular weight goes first), the end of that stretch is found by bisection and its rows are added
    // as one block; otherwise one row or one irregular weight at a time (large nodes: the two kinds alternate every few rows).
    constexpr int kStretch = 24;
    auto run = [&](const int hi, const int lo, const int shift) {
        int i = hi;
        while (i >= lo)
        {
            const double* seg = invd + (n - (i + shift));           // weights of rows i, i - 1, ..., lo
            const int len = i - lo + 1;
            if (head < nw)
            {
                const R wh = wf[head];
                if (wh >= seg[0]) { ideal = pos == 0 ? wh : ideal + wh / l2[pos + 1]; ++pos; ++head; continue; }
                if (pos == 0 || len < kStretch || wh >= seg[kStretch - 1])
                {
                    ideal = pos == 0 ? seg[0] : ideal + seg[0] / l2[pos + 1];
                    ++pos; --i;
                    continue;
                }
                const int c = static_cast<int>(std::partition_point(seg + kStretch, seg + len, [wh](const double v) { return !(wh >= v); }) - seg);
                ideal = add_quotients(ideal, seg, l2 + pos + 1, c);
                pos += c; i -= c;
                continue;
            }
            if (pos == 0) { ideal = seg[0]; ++pos; --i; continue; }
            ideal = add_quotients(ideal, seg, l2 + pos + 1, len);
            pos += len; i -= len;
        }
    };
    int hi = n - 1;
    for (int t = nu - 1; t >= 0; --t)
    {
        const int row = Q.u[t];
        const int shift = std::max(n1 - Q.a1[t], n2 - Q.a2[t]);
        if (hi > row) run(hi, row + 1, shift);
        if (Q.allzero[t]) run(row, row, shift);
        hi = row - 1;
    }
    if (hi >= 0) run(hi, 0, std::max(n1, n2));
    while (head < nw) { ideal = pos == 0 ? wf[head] : ideal + wf[head] / l2[pos + 1]; ++pos; ++head; }
    lap.mark(5);
    return (dcg1 / ideal) * (dcg2 / ideal);
}
} // namespace

R compute_priority_rows(smk_ctx* ctx, const R* W_parent, const R* W_child, const int n, const unsigned int* child_rows, const int n_child_rows,
                        const unsigned int* parent_rows, const int n_parent_rows)
{
    std::lock_guard<std::mutex> lock(g_priority_mutex);
    if (!child_rows) return compute_priority_on(ctx, W_parent, W_child, n);
    return priority_rows_impl(ctx, g_ps, g_rs, true, W_parent, W_child, n, child_rows, n_child_rows, parent_rows, n_parent_rows);
}

// --------------------------------------------------------------------------------------------------------------
// tree growing
// --------------------------------------------------------------------------------------------------------------
namespace {

// A priority score that may still be on its way: the tree driver lets the score of a left child be evaluated on a worker
// thread while the calling thread extracts, initialises and factors the right child (the score is host work on three
// m-vectors, ~5 ms per node at m = 320 000, during which the GPU would sit idle; the factorization is GPU work during which
// the host would). value is final once fut is no longer valid.
struct PendingPriority
{
    R value = R(0);
    std::future<R> fut;
    bool pending() const { return fut.valid(); }
    R get() { if (fut.valid()) value = fut.get(); return value; }
};

double seconds_since(const std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }

struct Stopwatch
{
    double& acc;
    std::chrono::steady_clock::time_point t0;
    explicit Stopwatch(double& a) : acc(a), t0(std::chrono::steady_clock::now()) {}
    ~Stopwatch() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// The rank-2 factors of one node. W (m x 2 in the reference) is zero outside the rows of the node's compacted matrix, so it is
// kept on those rows only: row rows[r] of W is (Wc[r], Wc[nrows + r]). The root (all_rows) holds every row. H is 2 x cols.
struct Factor
{
    std::vector<R> Wc, H;
    std::vector<unsigned int> rows;     // ascending
    bool all_rows = false;
    unsigned int cols = 0;
    size_t nrows(const unsigned int m) const { return all_rows ? m : rows.size(); }
    void release() { std::vector<R>().swap(Wc); std::vector<R>().swap(H); std::vector<unsigned int>().swap(rows); }
};

// the rows a parent vector (a column of `f`'s W) can be non-zero on, for the priority score of f's children
struct ParentRows
{
    const unsigned int* rows = nullptr;
    int count = 0;
    ParentRows() {}
    explicit ParentRows(const Factor& f) : rows(f.all_rows ? nullptr : f.rows.data()), count(f.all_rows ? 0 : static_cast<int>(f.rows.size())) {}
};

// What one thread needs to evaluate priority scores: a full-height m x 2 buffer that is zero except while a node's W is
// scattered into it, and — for the worker threads — a scratch set and a library context (stream + sort buffers) of their own.
struct Evaluator
{
    smk_ctx* ctx = nullptr;
    std::vector<R> dense;
    PriorityScratch ps;
    RowsScratch rs;
    double busy_s = 0;
};

// The worker threads' library contexts are made by the first Clust call on a device and kept for the later ones (creating and
// destroying two contexts is tens of milliseconds each way); NmfFinalize releases them (HierReleaseWorkers).
struct WorkerContexts
{
    smk_ctx* ctx[2] = {nullptr, nullptr};
    int device = -1;
    bool acquire(const int dev)
    {
        if (device == dev && ctx[0] && ctx[1]) return true;
        release();
        for (int w = 0; w < 2; ++w) if (smk_create(&ctx[w], dev) != SMK_OK) { ctx[w] = nullptr; release(); return false; }
        device = dev;
        return true;
    }
    void release()
    {
        for (int w = 0; w < 2; ++w) { if (ctx[w]) smk_destroy(ctx[w]); ctx[w] = nullptr; }
        device = -1;
    }
};
WorkerContexts g_workers;

struct MisSpeculated {};        // thrown (and caught) inside HierRun::grow: the leaf split ahead of time was not the right one

struct HierRun
{
    smk_ctx* ctx;
    const ClustOptions& opts;
    unsigned int m, n;
    Random& rng;
    ClustStats& stats;
    std::vector<unsigned int> new_to_old;      // scratch, m entries
    std::vector<R> Wsub, Hsub, Winit_full, Hinit_full, parent_scratch;
    int init_counter = 1;                      // Winit_<i>.csv / Hinit_<i>.csv, clust_hier_generic.hpp:586
    // Priority scores off the calling thread (SMK_HIER_ASYNC=0 turns it off): evaluator 0 is the calling thread's, 1 takes the
    // left children's scores, 2 the right children's.
    bool async_on = true;
    Evaluator eval[3];
    // SMK_HIER_PROF=1: seconds of the driver's own host work per kind, on stderr when the run ends
    bool prof_on = false;
    double t_tree = 0, t_buffers = 0, t_labels = 0;
    int speculations = 0, misspeculations = 0;

    HierRun(smk_ctx* c, const ClustOptions& o, Random& r, ClustStats& s)
        : ctx(c), opts(o), m(o.nmf_opts.height), n(o.nmf_opts.width), rng(r), stats(s), new_to_old(o.nmf_opts.height)
    {
        const char* e = getenv("SMK_HIER_ASYNC");
        async_on = !(e && atoi(e) == 0);
        const char* pe = getenv("SMK_HIER_PROF");
        prof_on = pe && atoi(pe) != 0;
        eval[0].ctx = ctx;
        if (async_on && !g_workers.acquire(smk_device_index(ctx))) async_on = false;
        if (async_on) { eval[1].ctx = g_workers.ctx[0]; eval[2].ctx = g_workers.ctx[1]; }
        for (int w = 0; w < (async_on ? 3 : 1); ++w) eval[w].dense.assign(static_cast<size_t>(m) * 2, R(0));
        // the log tables are grown here, on this thread, once: afterwards every thread only reads them
        g_logs.ensure(static_cast<int>(m));
        g_invd.ensure(static_cast<int>(m));
    }
    ~HierRun()
    {
        if (prof_on) fprintf(stderr, "hierclust driver: tree updates %.3f s, factor buffers %.3f s, labels + scatter %.3f s; %d leaves split ahead "
                                     "of the last score, %d of them taken back\n", t_tree, t_buffers, t_labels, speculations, misspeculations);
    }
    HierRun(const HierRun&) = delete;
    HierRun& operator=(const HierRun&) = delete;

    // LoadInitializers, clust_hier_generic.hpp:568-609
    void load_initializers()
    {
        std::ostringstream fw, fh;
        fw << opts.initdir << "Winit_" << init_counter << ".csv";
        fh << opts.initdir << "Hinit_" << init_counter << ".csv";
        unsigned int h = 0, w = 0;
        if (!smallk_io::LoadDelimitedFile(Winit_full, h, w, fw.str()) || h != m || w != 2)
            throw std::runtime_error("Load failed for file " + fw.str());
        if (!smallk_io::LoadDelimitedFile(Hinit_full, h, w, fh.str()) || h != 2 || w != n)
            throw std::runtime_error("Load failed for file " + fh.str());
        ++init_counter;
    }

    // One rank-2 NMF of the active matrix (height x width) from the initial guess in W, H. Throws what the
    // reference lets escape; returns false where NmfSolve returns false.
    bool factor(const int height, const int width, R* W, R* H)
    {
        NmfOptions o = opts.nmf_opts;
        o.height = height; o.width = width; o.k = 2; o.algorithm = NmfAlgorithm::RANK2;
        smk_nmf_options a = NmfToAbi(o);
        smk_nmf_stats st = {0, 0};
        int rc;
        { Stopwatch sw(stats.t_factor); rc = smk_nmf(ctx, &a, W, height, H, 2, &st); }
        stats.iteration_count += st.iteration_count;
        if (rc == SMK_OK)
        {
            stats.nmf_count += 1;
            if (st.iteration_count == o.max_iter) stats.max_count += 1;
            return true;
        }
        const std::string err = smk_last_error(ctx);
        NmfSetLastError(err.c_str());
        if (rc != SMK_FAILURE) throw std::runtime_error("HierNMF2: " + err);
        if (err.find("Normalize:") == 0 || err.find("ProjectedGradientNorm") == 0) throw std::runtime_error(err);
        return false;
    }

    // The priority score of `child` against the parent vector, on evaluator E: the child's W is scattered into E's full-height
    // buffer for the evaluation and taken out again (the buffer is zero between evaluations).
    R evaluate(Evaluator& E, const bool calling_thread, const R* W_parent, const ParentRows pr, const Factor& child)
    {
        const size_t cnt = child.nrows(m);
        R* D = E.dense.data();
        struct Scattered
        {
            R* D; const Factor& f; size_t cnt, m;
            Scattered(R* d, const Factor& ff, size_t c, size_t mm) : D(d), f(ff), cnt(c), m(mm)
            {
                if (f.all_rows) std::copy(f.Wc.begin(), f.Wc.end(), D);
                else for (size_t r = 0; r < cnt; ++r) { D[f.rows[r]] = f.Wc[r]; D[m + f.rows[r]] = f.Wc[cnt + r]; }
            }
            ~Scattered()
            {
                if (f.all_rows) std::fill(D, D + 2 * m, R(0));
                else for (size_t r = 0; r < cnt; ++r) { D[f.rows[r]] = R(0); D[m + f.rows[r]] = R(0); }
            }
        } guard(D, child, cnt, m);
        const unsigned int* rows = child.all_rows ? nullptr : child.rows.data();
        if (calling_thread) return compute_priority_rows(E.ctx, W_parent, D, static_cast<int>(m), rows, static_cast<int>(cnt), pr.rows, pr.count);
        const auto t0 = std::chrono::steady_clock::now();
        const R v = rows ? priority_rows_impl(E.ctx, E.ps, E.rs, false, W_parent, D, static_cast<int>(m), rows, static_cast<int>(cnt), pr.rows, pr.count)
                         : compute_priority_rows(E.ctx, W_parent, D, static_cast<int>(m), nullptr, 0);
        E.busy_s += seconds_since(t0);
        return v;
    }

    // clust_hier_generic.hpp:383-517
    // With `defer` (and evaluator `widx` free), the score of a regular split (two non-empty clusters, non-negative factors, more
    // than one positive parent entry — such a score is a product of two ratios of sums of non-negative weights: never negative)
    // is started on a worker thread and the value returned here is the placeholder 0; what the worker reads (W_parent, the rows
    // behind `pr`, `out`) must stay untouched until defer->get().
    R actual_split(const std::vector<unsigned int>& subset, const R* W_parent, const ParentRows pr, Factor& out,
                   std::vector<unsigned int>& labels, PendingPriority* defer = nullptr, const int widx = 1)
    {
        const size_t cnt = subset.size();
        out.cols = static_cast<unsigned int>(cnt);
        out.Wc.clear(); out.rows.clear(); out.all_rows = false;
        out.H.assign(cnt * 2, R(0));
        if (cnt <= 3) { labels.assign(cnt, 1u); return R(-1); }

        int new_height = 0;
        {
            Stopwatch sw(stats.t_extract);
            const int rc = smk_select_columns(ctx, subset.data(), static_cast<int>(cnt), &new_height, new_to_old.data());
            if (rc != SMK_OK) throw std::logic_error(smk_last_error(ctx));
        }

        Wsub.resize(static_cast<size_t>(new_height) * 2);
        Hsub.resize(cnt * 2);
        bool ok = false;
        for (int attempt = 0; attempt < 3 && !ok; ++attempt)
        {
            if (!opts.initdir.empty())
            {
                load_initializers();           // ExtractSubmatrices, clust_hier_generic.hpp:521-544
                for (int r = 0; r < new_height; ++r)
                {
                    Wsub[r] = Winit_full[new_to_old[r]];
                    Wsub[static_cast<size_t>(new_height) + r] = Winit_full[static_cast<size_t>(m) + new_to_old[r]];
                }
                for (size_t c = 0; c < cnt; ++c) { Hsub[2 * c] = Hinit_full[2 * static_cast<size_t>(subset[c])]; Hsub[2 * c + 1] = Hinit_full[2 * static_cast<size_t>(subset[c]) + 1]; }
            }
            else
            {
                Stopwatch sw(stats.t_init);
                RandomMatrix(Wsub.data(), new_height, new_height, 2, rng, R(0.5), R(0.5));
                RandomMatrix(Hsub.data(), 2, 2, static_cast<unsigned int>(cnt), rng, R(0.5), R(0.5));
            }
            ok = factor(new_height, static_cast<int>(cnt), Wsub.data(), Hsub.data());
            if (!ok) cout << "\nNode factorization failed, retrying with new initializers..." << endl;
        }
        if (!ok) throw std::runtime_error("HierNMF2: node factorization failed after three attempts.");

        bool has_0 = false, has_1 = false;
        {
            Stopwatch sw_labels(t_labels);
            labels.resize(cnt);
            for (size_t c = 0; c < cnt; ++c)
            {
                if (Hsub[2 * c] > Hsub[2 * c + 1]) { labels[c] = 0u; has_0 = true; }
                else { labels[c] = 1u; has_1 = true; }
            }
            out.Wc = Wsub;
            out.H = Hsub;
            out.rows.assign(new_to_old.begin(), new_to_old.begin() + new_height);
        }
        if (!(has_0 && has_1)) return R(-1);
        if (defer && async_on)
        {
            // what decides the sign of the score, and whether the streaming evaluation applies, is cheap to know now
            bool regular = true;
            int parent_pos = 0;
            auto look = [&](const R v) { if (!(v >= 0)) regular = false; parent_pos += (v > 0) ? 1 : 0; };
            if (pr.rows) { for (int t = 0; t < pr.count; ++t) look(W_parent[pr.rows[t]]); }
            else for (unsigned int r = 0; r < m; ++r) look(W_parent[r]);
            for (size_t e = 0; e < Wsub.size() && regular; ++e) if (!(Wsub[e] >= 0)) regular = false;
            if (regular && parent_pos > 1)
            {
                HierRun* self = this;
                Evaluator* E = &eval[widx];
                const Factor* child = &out;
                defer->fut = std::async(std::launch::async, [self, E, W_parent, pr, child]() { return self->evaluate(*E, false, W_parent, pr, *child); });
                return R(0);
            }
        }
        Stopwatch sw(stats.t_priority);
        return evaluate(eval[0], true, W_parent, pr, out);
    }

    // clust_hier_generic.hpp:245-376
    // min_priority() is asked for only where the reference uses the value (a small cluster's score is compared with it), which
    // lets the caller hand in a value that is still being established. With `defer` the returned score may still be pending
    // (defer->pending()): the caller takes it from defer->get().
    template <typename MinPriority>
    R trial_split(std::vector<unsigned int>& subset, MinPriority&& min_priority, const R* W_parent, const ParentRows pr, Factor& out,
                  PendingPriority* defer = nullptr, const int widx = 1)
    {
        const std::vector<unsigned int> backup(subset);
        std::vector<unsigned int> labels, small, labels_small;
        Factor tmp;
        int trial = 0;
        R priority = R(-2);
        while (trial < opts.trial_allowance)
        {
            priority = actual_split(subset, W_parent, pr, out, labels, defer, widx);
            if (priority < R(0)) break;                  // a pending score is not negative (actual_split)
            int counts[2] = {0, 0};
            for (unsigned int l : labels) counts[l] += 1;
            const int smallest = std::min(counts[0], counts[1]);
            if (!(smallest < opts.unbalanced * labels.size())) break;
            // the unbalanced path re-uses `out` and the row map: finish the pending score first
            if (defer && defer->pending()) { Stopwatch sw(stats.t_priority); priority = defer->get(); }

            const unsigned int small_label = (smallest == counts[0]) ? 0u : 1u;
            small.clear();
            for (size_t q = 0; q < labels.size(); ++q) if (labels[q] == small_label) small.push_back(subset[q]);
            // priority of the small cluster, with its own topic vector (a column of this split's W) as parent
            R priority_small;
            {
                const size_t cnt = out.rows.size();
                parent_scratch.resize(m, R(0));
                for (size_t r = 0; r < cnt; ++r) parent_scratch[out.rows[r]] = out.Wc[small_label * cnt + r];
                struct Clear
                {
                    std::vector<R>& v; const std::vector<unsigned int> rows;
                    ~Clear() { for (unsigned int r : rows) v[r] = R(0); }
                } clear{parent_scratch, out.rows};
                priority_small = actual_split(small, parent_scratch.data(), ParentRows(out), tmp, labels_small);
            }
            if (priority_small < min_priority())
            {
                trial += 1;
                if (trial < opts.trial_allowance)
                {
                    cout << "dropping " << small.size() << " items ..." << endl;
                    std::vector<unsigned int> kept;         // SetDiff(subset, small): both ascending
                    kept.reserve(subset.size() - small.size());
                    std::set_difference(subset.begin(), subset.end(), small.begin(), small.end(), std::back_inserter(kept));
                    subset.swap(kept);
                }
            }
            else break;
        }
        if (trial == opts.trial_allowance)
        {
            if (opts.verbose) cout << "recycling " << small.size() << " items ..." << endl;
            subset = backup;
            std::fill(out.Wc.begin(), out.Wc.end(), R(0));
            out.cols = static_cast<unsigned int>(subset.size());
            out.H.assign(subset.size() * 2, R(0));
            priority = R(-2);
        }
        return priority;
    }

    // clust_hier_generic.hpp:77-238
    bool grow(Tree<R>& tree)
    {
        const unsigned int num_clusters = opts.num_clusters;
        if (num_clusters <= 1) throw std::runtime_error("HierNMF2: number of clusters must be >= 2");
        const unsigned int node_count = 2 * (num_clusters - 1);
        tree.Init(num_clusters, node_count, m, n);

        // root: the whole matrix
        const auto t_grow0 = std::chrono::steady_clock::now();
        smk_select_all(ctx);
        Factor root;
        root.Wc.resize(static_cast<size_t>(m) * 2); root.H.resize(static_cast<size_t>(n) * 2); root.cols = n; root.all_rows = true;
        bool ok = false;
        for (int attempt = 0; attempt < 3 && !ok; ++attempt)
        {
            if (!opts.initdir.empty()) { load_initializers(); root.Wc = Winit_full; root.H = Hinit_full; }
            else
            {
                RandomMatrix(root.Wc.data(), m, m, 2, rng, R(0.5), R(0.5));
                RandomMatrix(root.H.data(), 2, 2, n, rng, R(0.5), R(0.5));
            }
            ok = factor(static_cast<int>(m), static_cast<int>(n), root.Wc.data(), root.H.data());
            if (!ok) cout << "\nRoot node factorization failed, retrying with new initializers..." << endl;
        }
        if (!ok) throw std::runtime_error("HierNMF2: root node factorization failed after three attempts");

        if (prof_on) fprintf(stderr, "hierclust driver: tree.Init + root factorization (initialisers included) %.3f s\n", seconds_since(t_grow0));
        const auto t_loop0 = std::chrono::steady_clock::now();
        std::vector<Factor> node_factor(node_count);
        R min_priority = std::numeric_limits<R>::infinity(), max_priority = 0;
        unsigned int split_index = 0;
        // Scores on their way: the left child's on evaluator 1 (under the right child's factorization), the right child's on
        // evaluator 2 (under the factorization of the NEXT split's left child, which is started ahead of time — see below).
        PendingPriority pend_left, pend_right;
        unsigned int pend_left_node = 0, pend_right_node = 0;
        int release_after_right = -1;          // the split node whose rows the pending right-child score still reads
        auto finish_left = [&]() {
            if (!pend_left.pending()) return;
            Stopwatch sw(stats.t_priority);
            tree.SetNodePriority(pend_left_node, pend_left.get());
        };
        auto finish_right = [&]() {
            if (pend_right.pending())
            {
                Stopwatch sw(stats.t_priority);
                tree.SetNodePriority(pend_right_node, pend_right.get());
            }
            // the factors of a node that has been split are never read again
            if (release_after_right >= 0) { node_factor[release_after_right].release(); release_after_right = -1; }
        };
        auto split_leaf = [&](const unsigned int q) {
            Stopwatch sw(t_tree);
            const Factor& f = node_factor[q];
            tree.SplitCompact(q, f.rows.data(), static_cast<unsigned int>(f.rows.size()), f.Wc.data(), f.H.data(), f.cols);
        };
        auto plain_min = [&]() { return min_priority; };

        bool ahead = false;                    // this iteration's split and left child were already done (ahead of time, confirmed)
        unsigned int i0 = 0, i1 = 0;
        for (unsigned int i = 0; i + 1 < num_clusters; ++i)
        {
            if (!ahead)
            {
                finish_left(); finish_right();
                if (0 == i) { Stopwatch sw(t_tree); tree.SplitRoot(root.Wc.data(), root.H.data(), root.cols); }
                else
                {
                    { Stopwatch sw(t_tree); tree.MinMaxLeafPriorities(min_priority, max_priority, split_index); }
                    if (max_priority < R(0)) { cout << "\nHierNMF2: no further factorization possible.\n" << endl; break; }
                    split_leaf(split_index);
                }
                i0 = tree.LeftChildIndex(); i1 = tree.RightChildIndex();
                const ParentRows pr(0 == i ? root : node_factor[split_index]);      // both topic vectors are columns of the split node's W
                const R p0 = trial_split(tree.LeftChildDocs(), plain_min, tree.LeftChildTopicVector().data(), pr, node_factor[i0], &pend_left, 1);
                if (pend_left.pending()) pend_left_node = i0; else tree.SetNodePriority(i0, p0);
            }
            ahead = false;
            {
                // the children's topic vectors live in the tree's nodes, which stay where they are
                const ParentRows pr(0 == i ? root : node_factor[split_index]);
                const R p1 = trial_split(tree.RightChildDocs(), plain_min, tree.RightChildTopicVector().data(), pr, node_factor[i1], &pend_right, 2);
                if (pend_right.pending()) pend_right_node = i1; else tree.SetNodePriority(i1, p1);
                if (i > 0) release_after_right = static_cast<int>(split_index);
            }
            finish_left();
            if (opts.verbose) { cout << "[" << (i + 1) << "] "; cout.flush(); }

            // Ahead of time: while the right child's score is evaluated, the leaf that will be split next UNLESS that score beats it
            // (the best of all the others, the reference's scan and tie rule) is split and its left child factored. When the
            // score arrives it either confirms the choice — the work is exactly what the next iteration would have done, from the same
            // generator state — or the tree, the generator and the counters are put back and the next iteration starts over.
            if (!(async_on && pend_right.pending() && i + 2 < num_clusters && opts.initdir.empty())) continue;
            R min_wo = 0, max_wo = 0;
            unsigned int best = 0;
            { Stopwatch sw(t_tree); tree.MinMaxLeafPrioritiesWithout(i1, Tree<R>::NONE, min_wo, max_wo, best); }
            if (max_wo < R(0)) continue;
            const Random saved_rng = rng;
            const int saved_nmf = stats.nmf_count, saved_max = stats.max_count;
            const long long saved_iters = stats.iteration_count;
            speculations += 1;
            split_leaf(best);
            const unsigned int s0 = tree.LeftChildIndex(), s1 = tree.RightChildIndex();
            bool confirmed = false;
            R next_min = 0;
            // the right child's score is in: was `best` the right leaf to split? if so, what MinMaxLeafPriorities would have found
            auto confirm = [&]() {
                if (confirmed) return next_min;
                finish_right();
                const R p1 = tree.NodePriority(i1);
                if (p1 > max_wo) throw MisSpeculated();
                next_min = (p1 > R(0) && p1 < min_wo) ? p1 : min_wo;
                confirmed = true;
                return next_min;
            };
            auto take_back = [&]() {
                if (pend_left.pending()) { try { pend_left.get(); } catch (...) {} }
                { Stopwatch sw(t_tree); tree.UndoSplit(best); }
                node_factor[s0] = Factor();
                rng = saved_rng;
                stats.nmf_count = saved_nmf; stats.max_count = saved_max; stats.iteration_count = saved_iters;
                misspeculations += 1;
            };
            try
            {
                const ParentRows pr(node_factor[best]);
                const R p0 = trial_split(tree.LeftChildDocs(), confirm, tree.LeftChildTopicVector().data(), pr, node_factor[s0], &pend_left, 1);
                if (pend_left.pending()) pend_left_node = s0; else tree.SetNodePriority(s0, p0);
                confirm();
            }
            catch (MisSpeculated&) { take_back(); continue; }
            catch (...)
            {
                // an error of a factorization that should never have been started is not an error
                bool wrong = false;
                if (!confirmed) { try { confirm(); } catch (MisSpeculated&) { wrong = true; } }
                if (wrong) { take_back(); continue; }
                throw;
            }
            ahead = true;
            split_index = best; i0 = s0; i1 = s1;
            min_priority = next_min; max_priority = max_wo;
        }
        finish_left(); finish_right();
        stats.t_priority_worker += eval[1].busy_s + eval[2].busy_s;
        if (prof_on) fprintf(stderr, "hierclust driver: split loop %.3f s\n", seconds_since(t_loop0));
        smk_select_all(ctx);
        { Stopwatch sw(stats.t_terms); tree.ComputeTopTerms(opts.maxterms); }
        tree.ComputeAssignments();
        cout << endl;
        return true;
    }

    // ClustFlat, hierclust/include/clust_flat_generic.hpp:33-74
    bool flat(Tree<R>& tree, R* buf_w, R* buf_h)
    {
        const unsigned int k = opts.num_clusters;
        if (!tree.FlatclustInitW(buf_w, m, m, k)) return false;
        smk_select_all(ctx);
        bool ok = false;
        for (int attempt = 0; attempt < 3 && !ok; ++attempt)
        {
            RandomMatrix(buf_h, k, k, n, rng, R(0.5), R(0.5));
            int iters = 0;
            const int rc = smk_nnls_hals(ctx, static_cast<int>(k), buf_w, static_cast<int>(m), buf_h, static_cast<int>(k),
                                         opts.nmf_opts.tol, opts.nmf_opts.max_iter, &iters);
            if (rc == SMK_OK) ok = true;
            else
            {
                const std::string err = smk_last_error(ctx);
                if (rc != SMK_FAILURE || err.find("Normalize:") == 0) throw std::runtime_error(err);
                cerr << err << endl;
            }
        }
        if (!ok) cout << "Flatclust NNLS solver failed after 3 attempts." << endl;
        return ok;
    }
};

Result check_sizes(const ClustOptions& options)
{
    if (!NmfContext())
    {
        cerr << "clustlib error: nmf_initialize() must be called prior to any clustering routine\n" << endl;
        return Result::NOTINITIALIZED;
    }
    if (!IsValid(options)) return Result::BAD_PARAM;
    const unsigned long long lim = std::numeric_limits<int>::max();
    if (2ull * options.nmf_opts.height > lim) { cerr << "W matrix size too large" << endl; return Result::SIZE_TOO_LARGE; }
    if (2ull * options.nmf_opts.width > lim) { cerr << "H matrix size too large" << endl; return Result::SIZE_TOO_LARGE; }
    return Result::OK;
}

Result run(const ClustOptions& options, R* buf_w, R* buf_h, Tree<R>& tree, ClustStats& stats, Random& rng)
{
    const auto t0 = std::chrono::steady_clock::now();
    HierRun job(NmfContext(), options, rng, stats);
    const double t_ctor = seconds_since(t0);
    const auto t1 = std::chrono::steady_clock::now();
    const bool grown = job.grow(tree);
    if (job.prof_on) fprintf(stderr, "hierclust driver: set-up %.3f s, grow() %.3f s\n", t_ctor, seconds_since(t1));
    if (!grown) return Result::FAILURE;
    if (options.flat && !job.flat(tree, buf_w, buf_h))
    {
        cerr << "Flat clustering failed." << endl;
        return Result::FLATCLUST_FAILURE;
    }
    return Result::OK;
}

} // namespace

Result Clust(const ClustOptions& options, R* buf_a, const int ldim_a, R* buf_w, R* buf_h, Tree<R>& tree, ClustStats& stats,
             Random& rng)
{
    const Result r = check_sizes(options);
    if (Result::OK != r) return r;
    if (ldim_a < options.nmf_opts.height) throw std::logic_error("invalid leading dimension for input matrix");
    const int rc = smk_load_dense(NmfContext(), buf_a, ldim_a, options.nmf_opts.height, options.nmf_opts.width);
    if (rc != SMK_OK) { NmfSetLastError(smk_last_error(NmfContext())); return NmfFromAbi(rc); }
    return run(options, buf_w, buf_h, tree, stats, rng);
}

Result ClustSparse(const ClustOptions& options, const SparseMatrix<R>& A, R* buf_w, R* buf_h, Tree<R>& tree, ClustStats& stats,
                   Random& rng)
{
    const Result r = check_sizes(options);
    if (Result::OK != r) return r;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = smk_load_csc(NmfContext(), static_cast<int>(A.Height()), static_cast<int>(A.Width()), A.Size(),
                                A.LockedColBuffer(), A.LockedRowBuffer(), A.LockedDataBuffer());
    if (rc != SMK_OK) { NmfSetLastError(smk_last_error(NmfContext())); return NmfFromAbi(rc); }
    const char* pe = getenv("SMK_HIER_PROF");
    const bool prof = pe && atoi(pe) != 0;
    if (prof) fprintf(stderr, "hierclust driver: smk_load_csc %.3f s\n", seconds_since(t0));
    const Result res = run(options, buf_w, buf_h, tree, stats, rng);
    if (prof) fprintf(stderr, "hierclust driver: ClustSparse %.3f s in all\n", seconds_since(t0));
    return res;
}

void HierReleaseWorkers() { g_workers.release(); }
