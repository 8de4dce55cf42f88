// Host sparse LU (stand-in for the un-vendored BASICLU). See sparse_lu.h.

#include "sparse_lu.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <numeric>
#include <utility>

namespace ipxb200 {

namespace {
using I = int64_t;

// Sorts the (index,value) pairs of every column by index.
void SortColumns(const std::vector<I>& ptr, std::vector<I>& idx,
                 std::vector<double>& val) {
    std::vector<std::pair<I, double>> work;
    const I ncol = static_cast<I>(ptr.size()) - 1;
    for (I k = 0; k < ncol; k++) {
        const I b = ptr[k], e = ptr[k + 1];
        bool sorted = true;
        for (I p = b + 1; p < e; p++)
            if (idx[p - 1] > idx[p]) { sorted = false; break; }
        if (sorted) continue;
        work.clear();
        for (I p = b; p < e; p++) work.emplace_back(idx[p], val[p]);
        std::sort(work.begin(), work.end());
        for (I p = b; p < e; p++) {
            idx[p] = work[p - b].first;
            val[p] = work[p - b].second;
        }
    }
}
}  // namespace

void SparseLuFactorize(I dim, const I* Bbegin, const I* Bend, const I* Bi,
                       const double* Bx, double pivottol, double abstol,
                       SparseLuResult* out) {
    SparseLuResult& R = *out;
    R = SparseLuResult();
    if (abstol <= 0.0) abstol = 1e-300;

    // Row-wise pattern of B.
    std::vector<I> Rp(dim + 1, 0);
    for (I j = 0; j < dim; j++)
        for (I p = Bbegin[j]; p < Bend[j]; p++) Rp[Bi[p] + 1]++;
    for (I i = 0; i < dim; i++) Rp[i + 1] += Rp[i];
    std::vector<I> Rj(Rp[dim]);
    {
        std::vector<I> next(Rp.begin(), Rp.end() - 1);
        for (I j = 0; j < dim; j++)
            for (I p = Bbegin[j]; p < Bend[j]; p++) Rj[next[Bi[p]]++] = j;
    }

    std::vector<I> rowpos(dim, -1), colpos(dim, -1);
    std::vector<I>& prow = R.rowperm;
    std::vector<I>& pcol = R.colperm;
    prow.reserve(dim);
    pcol.reserve(dim);
    // Factors with ORIGINAL row ids while pivoting; remapped at the end.
    std::vector<I>&Lp = R.Lp, &Li = R.Li, &Up = R.Up, &Ui = R.Ui;
    std::vector<double>&Lx = R.Lx, &Ux = R.Ux;
    Lp.assign(1, 0);
    Up.assign(1, 0);
    auto close_column = [&]() {
        Lp.push_back(static_cast<I>(Li.size()));
        Up.push_back(static_cast<I>(Ui.size()));
    };

    // ---- Phase 1: column singletons (upper triangular leading block). ----
    std::vector<I> ccount(dim);
    std::vector<I> queue;
    for (I j = 0; j < dim; j++) {
        ccount[j] = Bend[j] - Bbegin[j];
        if (ccount[j] == 1) queue.push_back(j);
    }
    for (size_t head = 0; head < queue.size(); head++) {
        const I j = queue[head];
        if (colpos[j] >= 0 || ccount[j] != 1) continue;
        I r = -1;
        double a = 0.0;
        for (I p = Bbegin[j]; p < Bend[j]; p++)
            if (rowpos[Bi[p]] < 0) { r = Bi[p]; a = Bx[p]; break; }
        if (r < 0 || !(std::abs(a) >= abstol)) continue;
        const I k = static_cast<I>(prow.size());
        for (I p = Bbegin[j]; p < Bend[j]; p++)
            if (Bi[p] != r && Bx[p] != 0.0) {
                Ui.push_back(Bi[p]);
                Ux.push_back(Bx[p]);
            }
        Ui.push_back(r);
        Ux.push_back(a);
        close_column();
        prow.push_back(r);
        pcol.push_back(j);
        rowpos[r] = k;
        colpos[j] = k;
        for (I q = Rp[r]; q < Rp[r + 1]; q++) {
            const I j2 = Rj[q];
            if (colpos[j2] < 0 && --ccount[j2] == 1) queue.push_back(j2);
        }
    }
    R.num_col_singletons = static_cast<I>(prow.size());

    // ---- Phase 2: row singletons (fill-free L columns). ----
    std::vector<I> rcount(dim, 0);
    queue.clear();
    for (I i = 0; i < dim; i++) {
        if (rowpos[i] >= 0) continue;
        I c = 0;
        for (I q = Rp[i]; q < Rp[i + 1]; q++)
            if (colpos[Rj[q]] < 0) c++;
        rcount[i] = c;
        if (c == 1) queue.push_back(i);
    }
    for (size_t head = 0; head < queue.size(); head++) {
        const I r = queue[head];
        if (rowpos[r] >= 0 || rcount[r] != 1) continue;
        I j = -1;
        for (I q = Rp[r]; q < Rp[r + 1]; q++)
            if (colpos[Rj[q]] < 0) { j = Rj[q]; break; }
        if (j < 0) continue;
        double a = 0.0, cmax = 0.0;
        for (I p = Bbegin[j]; p < Bend[j]; p++) {
            if (rowpos[Bi[p]] >= 0) continue;
            cmax = std::max(cmax, std::abs(Bx[p]));
            if (Bi[p] == r) a = Bx[p];
        }
        if (!(std::abs(a) >= abstol) || std::abs(a) < pivottol * cmax)
            continue;  // unstable pivot: leave row and column to the bump
        const I k = static_cast<I>(prow.size());
        for (I p = Bbegin[j]; p < Bend[j]; p++) {
            const I i = Bi[p];
            if (i == r || Bx[p] == 0.0) continue;
            if (rowpos[i] >= 0) {
                Ui.push_back(i);
                Ux.push_back(Bx[p]);
            } else {
                Li.push_back(i);
                Lx.push_back(Bx[p] / a);
            }
        }
        Ui.push_back(r);
        Ux.push_back(a);
        close_column();
        prow.push_back(r);
        pcol.push_back(j);
        rowpos[r] = k;
        colpos[j] = k;
        for (I p = Bbegin[j]; p < Bend[j]; p++) {
            const I i = Bi[p];
            if (rowpos[i] < 0 && --rcount[i] == 1) queue.push_back(i);
        }
    }
    const I n2 = static_cast<I>(prow.size());
    R.num_row_singletons = n2 - R.num_col_singletons;

    // ---- Phase 3: left-looking Gilbert-Peierls on the bump. ----
    std::vector<I> bump_cols;
    for (I j = 0; j < dim; j++)
        if (colpos[j] < 0) bump_cols.push_back(j);
    R.bump_size = static_cast<I>(bump_cols.size());
    {
        std::vector<I> cnt(dim, 0);
        for (I j : bump_cols)
            for (I p = Bbegin[j]; p < Bend[j]; p++)
                if (rowpos[Bi[p]] < 0) cnt[j]++;
        std::stable_sort(bump_cols.begin(), bump_cols.end(),
                         [&](I a, I b) { return cnt[a] < cnt[b]; });
    }
    std::vector<double> x(dim, 0.0);
    std::vector<I> mark(dim, -1), topo, stack_node, stack_ptr, deferred;
    I stamp = 0;
    for (I j : bump_cols) {
        stamp++;
        topo.clear();
        const size_t u_begin = Ui.size();
        for (I p = Bbegin[j]; p < Bend[j]; p++) {
            const I i0 = Bi[p];
            const I k0 = rowpos[i0];
            if (k0 >= 0 && k0 < n2) {
                // Pivot row of a singleton stage: no elimination applies.
                if (Bx[p] != 0.0) {
                    Ui.push_back(i0);
                    Ux.push_back(Bx[p]);
                }
                continue;
            }
            x[i0] = Bx[p];
            if (mark[i0] == stamp) continue;
            mark[i0] = stamp;
            stack_node.push_back(i0);
            stack_ptr.push_back(k0 >= 0 ? Lp[k0] : 0);
            while (!stack_node.empty()) {
                const I i = stack_node.back();
                const I k = rowpos[i];
                bool descended = false;
                if (k >= 0) {
                    I q = stack_ptr.back();
                    const I qend = Lp[k + 1];
                    while (q < qend) {
                        const I i2 = Li[q++];
                        if (mark[i2] != stamp) {
                            mark[i2] = stamp;
                            stack_ptr.back() = q;
                            stack_node.push_back(i2);
                            const I k2 = rowpos[i2];
                            stack_ptr.push_back(k2 >= 0 ? Lp[k2] : 0);
                            descended = true;
                            break;
                        }
                    }
                }
                if (!descended) {
                    stack_node.pop_back();
                    stack_ptr.pop_back();
                    topo.push_back(i);
                }
            }
        }
        // Numeric solve in topological order (reverse finishing order).
        for (size_t t = topo.size(); t-- > 0;) {
            const I i = topo[t];
            const I k = rowpos[i];
            if (k < 0) continue;
            const double xk = x[i];
            if (xk == 0.0) continue;
            for (I q = Lp[k]; q < Lp[k + 1]; q++) x[Li[q]] -= Lx[q] * xk;
        }
        // Pivot search among non-pivotal rows.
        double xmax = 0.0;
        for (I i : topo)
            if (rowpos[i] < 0) xmax = std::max(xmax, std::abs(x[i]));
        if (!(xmax >= abstol)) {
            deferred.push_back(j);
            Ui.resize(u_begin);
            Ux.resize(u_begin);
            for (I i : topo) x[i] = 0.0;
            continue;
        }
        I r = -1;
        {
            const double thresh = pivottol * xmax;
            I best_count = 0;
            double best_abs = 0.0;
            for (I i : topo) {
                if (rowpos[i] >= 0) continue;
                const double ax = std::abs(x[i]);
                if (ax < thresh || ax < abstol) continue;
                if (r < 0 || rcount[i] < best_count ||
                    (rcount[i] == best_count && ax > best_abs)) {
                    r = i;
                    best_count = rcount[i];
                    best_abs = ax;
                }
            }
        }
        assert(r >= 0);
        const double pivot = x[r];
        const I k = static_cast<I>(prow.size());
        for (I i : topo) {
            const double xi = x[i];
            x[i] = 0.0;
            if (i == r || xi == 0.0) continue;
            if (rowpos[i] >= 0) {
                Ui.push_back(i);
                Ux.push_back(xi);
            } else {
                Li.push_back(i);
                Lx.push_back(xi / pivot);
            }
        }
        Ui.push_back(r);
        Ux.push_back(pivot);
        close_column();
        prow.push_back(r);
        pcol.push_back(j);
        rowpos[r] = k;
        colpos[j] = k;
    }

    // ---- Phase 4: dependent columns go last, paired with leftover rows. ----
    {
        size_t d = 0;
        for (I i = 0; i < dim && d < deferred.size(); i++) {
            if (rowpos[i] >= 0) continue;
            const I k = static_cast<I>(prow.size());
            const I j = deferred[d++];
            Ui.push_back(i);
            Ux.push_back(1.0);
            close_column();
            prow.push_back(i);
            pcol.push_back(j);
            rowpos[i] = k;
            colpos[j] = k;
            R.dependent_cols.push_back(k);
        }
        assert(d == deferred.size());
        assert(static_cast<I>(prow.size()) == dim);
    }

    // ---- Phase 5: permuted row indices, sorted columns. ----
    for (I& i : Li) i = rowpos[i];
    for (I& i : Ui) i = rowpos[i];
    SortColumns(Lp, Li, Lx);
    SortColumns(Up, Ui, Ux);
#ifndef NDEBUG
    for (I k = 0; k < dim; k++) {
        for (I p = Lp[k]; p < Lp[k + 1]; p++) assert(Li[p] > k);
        for (I p = Up[k]; p < Up[k + 1]; p++) assert(Ui[p] <= k);
        assert(Up[k + 1] > Up[k] && Ui[Up[k + 1] - 1] == k);
    }
#endif
}

}  // namespace ipxb200
