// Host sparse LU (stand-in for the un-vendored BASICLU). See sparse_lu.h.
//
// Right-looking factorization with Markowitz pivot search and threshold
// partial pivoting on an explicitly held active submatrix (the scheme BASICLU
// documents: columns/rows are searched in order of increasing count, a pivot
// must reach pivottol times the largest entry of its column), followed by a
// dense, blocked and threaded factorization of the trailing block once the
// active submatrix has filled in (linking rows of block-angular LPs end there).

#include "sparse_lu.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <utility>

namespace ipxb200 {

namespace {
using I = int64_t;

struct Ent {
    I i;
    double v;
};

// Doubly linked bucket lists over nodes 0..n-1 keyed by count 0..n.
struct Buckets {
    std::vector<I> head, next, prev, key;
    void Init(I n) {
        head.assign((size_t)n + 2, -1);
        next.assign((size_t)n, -1);
        prev.assign((size_t)n, -1);
        key.assign((size_t)n, -1);
    }
    void Insert(I node, I k) {
        key[node] = k;
        prev[node] = -1;
        next[node] = head[k];
        if (head[k] >= 0) prev[head[k]] = node;
        head[k] = node;
    }
    void Remove(I node) {
        const I k = key[node];
        if (k < 0) return;
        if (prev[node] >= 0) next[prev[node]] = next[node];
        else head[k] = next[node];
        if (next[node] >= 0) prev[next[node]] = prev[node];
        key[node] = -1;
    }
    void Move(I node, I k) {
        if (key[node] == k) return;
        Remove(node);
        Insert(node, k);
    }
};

// ---------------------------------------------------------------- dense tail

typedef double v4d __attribute__((vector_size(32), aligned(8)));

// C[0:8, 0:4] -= L[0:8, 0:nb] * U[0:nb, 0:4]; column-major, leading dimension ld for all three
// (they are windows of one array).
#define IPXB200_DENSE_KERNEL_BODY                                               \
    v4d c00 = *(const v4d*)(C), c10 = *(const v4d*)(C + 4);                     \
    v4d c01 = *(const v4d*)(C + ld), c11 = *(const v4d*)(C + ld + 4);           \
    v4d c02 = *(const v4d*)(C + 2 * ld), c12 = *(const v4d*)(C + 2 * ld + 4);   \
    v4d c03 = *(const v4d*)(C + 3 * ld), c13 = *(const v4d*)(C + 3 * ld + 4);   \
    for (I k = 0; k < nb; k++) {                                                \
        const v4d l0 = *(const v4d*)(L + k * ld), l1 = *(const v4d*)(L + k * ld + 4); \
        const double u0 = U[k], u1 = U[k + ld], u2 = U[k + 2 * ld], u3 = U[k + 3 * ld]; \
        c00 -= l0 * u0; c10 -= l1 * u0;                                         \
        c01 -= l0 * u1; c11 -= l1 * u1;                                         \
        c02 -= l0 * u2; c12 -= l1 * u2;                                         \
        c03 -= l0 * u3; c13 -= l1 * u3;                                         \
    }                                                                           \
    *(v4d*)(C) = c00; *(v4d*)(C + 4) = c10;                                     \
    *(v4d*)(C + ld) = c01; *(v4d*)(C + ld + 4) = c11;                           \
    *(v4d*)(C + 2 * ld) = c02; *(v4d*)(C + 2 * ld + 4) = c12;                   \
    *(v4d*)(C + 3 * ld) = c03; *(v4d*)(C + 3 * ld + 4) = c13;

#if defined(__x86_64__)
__attribute__((target("avx2,fma"))) void Kernel8x4Avx2(double* C, const double* L,
                                                       const double* U, I nb, I ld) {
    IPXB200_DENSE_KERNEL_BODY
}
#endif
void Kernel8x4Generic(double* C, const double* L, const double* U, I nb, I ld) {
    IPXB200_DENSE_KERNEL_BODY
}

typedef void (*Kernel8x4)(double*, const double*, const double*, I, I);

Kernel8x4 PickKernel() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return Kernel8x4Avx2;
#endif
    return Kernel8x4Generic;
}

// A22[r0:r1, c0:c1] -= L21[r0:r1, 0:nb] * U12[0:nb, c0:c1], where L21 starts at column lcol and
// U12 at row urow of the same column-major array D (leading dimension ld).
void TrailingUpdate(double* D, I ld, I lcol, I urow, I nb, I r0, I r1, I c0, I c1,
                    Kernel8x4 kernel) {
    const I kRowTile = 256;
    for (I rt = r0; rt < r1; rt += kRowTile) {
        const I re = std::min(r1, rt + kRowTile);
        I c = c0;
        for (; c + 4 <= c1; c += 4) {
            const double* U = D + urow + c * ld;
            I r = rt;
            for (; r + 8 <= re; r += 8) kernel(D + r + c * ld, D + r + lcol * ld, U, nb, ld);
            for (; r < re; r++) {
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                const double* L = D + r + lcol * ld;
                for (I k = 0; k < nb; k++) {
                    const double l = L[k * ld];
                    s0 += l * U[k];
                    s1 += l * U[k + ld];
                    s2 += l * U[k + 2 * ld];
                    s3 += l * U[k + 3 * ld];
                }
                D[r + c * ld] -= s0;
                D[r + (c + 1) * ld] -= s1;
                D[r + (c + 2) * ld] -= s2;
                D[r + (c + 3) * ld] -= s3;
            }
        }
        for (; c < c1; c++) {
            const double* U = D + urow + c * ld;
            for (I k = 0; k < nb; k++) {
                const double u = U[k];
                if (u == 0.0) continue;
                const double* L = D + lcol * ld + k * ld;
                double* Cc = D + c * ld;
                for (I r = rt; r < re; r++) Cc[r] -= L[r] * u;
            }
        }
    }
}

int DenseThreads() {
    if (const char* env = std::getenv("IPXB200_LU_THREADS")) return std::max(1, std::atoi(env));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(16u, std::max(1u, hw));
}

// In-place LU of the nr x nc column-major matrix D (ld = nr) with row partial pivoting.
// Columns without an acceptable pivot (all eligible entries below abstol) are moved behind the
// pivotal ones. On return rows [0, npiv) / columns [0, npiv) hold the factors (unit lower part
// below the diagonal, upper part on and above it), rows_of[r] / cols_of[c] give the original
// row / column held at position r / c.
I DenseLu(double* D, I nr, I nc, double abstol, std::vector<I>* rows_of, std::vector<I>* cols_of) {
    const I ld = nr;
    rows_of->resize((size_t)nr);
    cols_of->resize((size_t)nc);
    std::iota(rows_of->begin(), rows_of->end(), (I)0);
    std::iota(cols_of->begin(), cols_of->end(), (I)0);
    const Kernel8x4 kernel = PickKernel();
    const int nthreads = DenseThreads();
    const I kPanel = 64;
    I npiv = 0;       // pivots found = next pivot row / column position
    I nact = nc;      // columns [npiv, nact) are still candidates
    while (npiv < nact && npiv < nr) {
        const I p0 = npiv;
        // ---- panel: unblocked elimination of up to kPanel columns ----
        I k = p0;
        while (k < std::min(nact, p0 + kPanel) && k < nr) {
            double* col = D + k * ld;
            I best = -1;
            double bmax = 0.0;
            for (I r = k; r < nr; r++) {
                const double a = std::abs(col[r]);
                if (a > bmax) {
                    bmax = a;
                    best = r;
                }
            }
            if (best < 0 || !(bmax >= abstol)) {
                // dependent column: swap it behind the candidates (its updated rows above k
                // are dropped by the caller)
                --nact;
                if (k != nact) {
                    for (I r = 0; r < nr; r++) std::swap(D[r + k * ld], D[r + nact * ld]);
                    std::swap((*cols_of)[k], (*cols_of)[nact]);
                    // the swapped-in column has not seen the eliminations of this panel yet
                    double* c2 = D + k * ld;
                    for (I q = p0; q < k; q++) {
                        const double u = c2[q];
                        if (u == 0.0) continue;
                        const double* lq = D + q * ld;
                        for (I r = q + 1; r < nr; r++) c2[r] -= lq[r] * u;
                    }
                }
                continue;
            }
            if (best != k) {
                for (I c = 0; c < nc; c++) std::swap(D[k + c * ld], D[best + c * ld]);
                std::swap((*rows_of)[k], (*rows_of)[best]);
            }
            const double piv = col[k];
            for (I r = k + 1; r < nr; r++) col[r] /= piv;
            // rank-1 update of the remaining panel columns
            const I pend = std::min(nact, p0 + kPanel);
            for (I c = k + 1; c < pend; c++) {
                double* cc = D + c * ld;
                const double u = cc[k];
                if (u == 0.0) continue;
                for (I r = k + 1; r < nr; r++) cc[r] -= col[r] * u;
            }
            k++;
        }
        const I nb = k - p0;
        npiv = k;
        if (nb == 0) break;
        if (npiv >= nact) break;
        // ---- U12 = L11^{-1} A12 and A22 -= L21 U12 over the columns right of the panel ----
        const I c_begin = npiv, c_end = nc;  // dependent columns keep being updated: harmless
        const I ncols = c_end - c_begin;
        auto work = [&](I ca, I cb) {
            for (I c = ca; c < cb; c++) {
                double* cc = D + c * ld;
                for (I q = p0; q < npiv; q++) {
                    const double u = cc[q];
                    if (u == 0.0) continue;
                    const double* lq = D + q * ld;
                    for (I r = q + 1; r < npiv; r++) cc[r] -= lq[r] * u;
                }
            }
            if (npiv < nr) TrailingUpdate(D, ld, p0, p0, nb, npiv, nr, ca, cb, kernel);
        };
        const double flops = 2.0 * (double)nb * (double)(nr - p0) * (double)ncols;
        const int nt = flops < 4e6 ? 1 : (int)std::min<I>(nthreads, std::max<I>(1, ncols / 8));
        if (nt <= 1) {
            work(c_begin, c_end);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; t++) {
                I ca = c_begin + ncols * t / nt, cb = c_begin + ncols * (t + 1) / nt;
                ca = c_begin + ((ca - c_begin) & ~(I)3);  // kernel works on 4 columns
                if (t + 1 < nt) cb = c_begin + ((cb - c_begin) & ~(I)3);
                if (ca < cb) pool.emplace_back(work, ca, cb);
            }
            for (std::thread& th : pool) th.join();
        }
    }
    return npiv;
}

// Sorts the (index,value) pairs of every column by index.
void SortColumns(const std::vector<I>& ptr, std::vector<I>& idx, std::vector<double>& val) {
    std::vector<std::pair<I, double>> work;
    const I ncol = static_cast<I>(ptr.size()) - 1;
    for (I k = 0; k < ncol; k++) {
        const I b = ptr[k], e = ptr[k + 1];
        bool sorted = true;
        for (I p = b + 1; p < e; p++)
            if (idx[p - 1] > idx[p]) {
                sorted = false;
                break;
            }
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

void SparseLuFactorize(I dim, const I* Bbegin, const I* Bend, const I* Bi, const double* Bx,
                       double pivottol, double abstol, SparseLuResult* out) {
    SparseLuResult& R = *out;
    R = SparseLuResult();
    static const bool timing = std::getenv("IPXB200_LU_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto seconds = [&]() {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    };
    double t_sparse = 0.0, t_dense = 0.0;
    I dense_rows = 0, dense_cols = 0;
    if (abstol <= 0.0) abstol = 1e-300;
    if (!(pivottol > 0.0)) pivottol = 0.1;
    if (pivottol > 1.0) pivottol = 1.0;

    // ---- active submatrix: columns with values, rows as patterns ----
    std::vector<std::vector<Ent>> col((size_t)dim);
    std::vector<std::vector<I>> row((size_t)dim);
    {
        std::vector<I> rc((size_t)dim, 0);
        for (I j = 0; j < dim; j++)
            for (I p = Bbegin[j]; p < Bend[j]; p++) rc[Bi[p]]++;
        for (I i = 0; i < dim; i++) row[i].reserve((size_t)rc[i] + 4);
        for (I j = 0; j < dim; j++) {
            col[j].reserve((size_t)(Bend[j] - Bbegin[j]) + 4);
            for (I p = Bbegin[j]; p < Bend[j]; p++) {
                if (Bx[p] == 0.0) continue;
                col[j].push_back(Ent{Bi[p], Bx[p]});
                row[Bi[p]].push_back(j);
            }
        }
    }
    long long act_nnz = 0;
    for (I j = 0; j < dim; j++) act_nnz += (long long)col[j].size();
    Buckets cb, rb;
    cb.Init(dim);
    rb.Init(dim);
    for (I j = 0; j < dim; j++) cb.Insert(j, (I)col[j].size());
    for (I i = 0; i < dim; i++) rb.Insert(i, (I)row[i].size());
    std::vector<double> colmax((size_t)dim, -1.0);  // < 0: not known
    auto col_max = [&](I j) {
        if (colmax[j] < 0.0) {
            double mx = 0.0;
            for (const Ent& e : col[j]) mx = std::max(mx, std::abs(e.v));
            colmax[j] = mx;
        }
        return colmax[j];
    };

    std::vector<I> rowpos((size_t)dim, -1), colpos((size_t)dim, -1);
    std::vector<I>&prow = R.rowperm, &pcol = R.colperm;
    prow.reserve((size_t)dim);
    pcol.reserve((size_t)dim);
    // L: one column per pivot, ORIGINAL row ids until the end. U: per ORIGINAL column the
    // entries (pivot position, value) in pivot order, i.e. already sorted, diagonal last.
    std::vector<I>&Lp = R.Lp, &Li = R.Li;
    std::vector<double>& Lx = R.Lx;
    Lp.assign(1, 0);
    std::vector<std::vector<Ent>> ucol((size_t)dim);
    std::vector<I> deferred;  // columns without a pivot
    std::vector<char> col_dead((size_t)dim, 0);

    std::vector<double> mult((size_t)dim, 0.0);
    std::vector<I> mark((size_t)dim, -1), hit((size_t)dim, -1), clist;
    I stamp = 0, jstamp = 0;
    I nact_rows = dim, nact_cols = dim;

    // Row patterns are pruned lazily: a column that left the active submatrix (pivotal or dead)
    // stays in the lists of its rows until they are compacted; rb.key holds the exact count.
    // (Searching a linking row of a block-angular basis for the entry costs its whole length.)
    auto gone = [&](I j) { return colpos[j] >= 0 || col_dead[j]; };
    auto drop_from_row = [&](I i) {
        const I cnt = rb.key[i] - 1;
        std::vector<I>& r = row[i];
        if ((I)r.size() > 2 * cnt + 8) {
            size_t put = 0;
            for (size_t t = 0; t < r.size(); t++)
                if (!gone(r[t])) r[put++] = r[t];
            r.resize(put);
        }
        rb.Move(i, cnt);
    };
    // Removes column j from the active submatrix without a pivot.
    auto kill_column = [&](I j) {
        col_dead[j] = 1;
        for (const Ent& e : col[j]) drop_from_row(e.i);
        act_nnz -= (long long)col[j].size();
        std::vector<Ent>().swap(col[j]);
        cb.Remove(j);
        deferred.push_back(j);
        nact_cols--;
    };

    const int kMaxSearch = 4;
    const double kDenseFill = 0.30;
    const I kDenseMin = 48;

    for (;;) {
        if (nact_cols == 0 || nact_rows == 0) break;
        // ---- switch to the dense kernel once the active submatrix has filled in ----
        if (nact_cols >= kDenseMin && (double)nact_rows * (double)nact_cols <= 6e8 &&
            (double)act_nnz >= kDenseFill * (double)nact_rows * (double)nact_cols) {
            break;
        }
        // ---- Markowitz search ----
        I bp = -1, bq = -1;
        double bcost = -1.0, babs = 0.0;
        int searched = 0;
        bool done = false;
        for (I nz = 1; nz <= dim && !done; nz++) {
            // columns with nz entries
            for (I j = cb.head[nz]; j >= 0 && !done;) {
                const I jn = cb.next[j];
                const double cmax = col_max(j);
                if (!(cmax >= abstol)) {
                    kill_column(j);
                    j = jn;
                    continue;
                }
                const double thr = std::max(abstol, pivottol * cmax);
                for (const Ent& e : col[j]) {
                    const double a = std::abs(e.v);
                    if (a < thr) continue;
                    const double cost = (double)(rb.key[e.i] - 1) * (double)(nz - 1);
                    if (bcost < 0.0 || cost < bcost || (cost == bcost && a > babs)) {
                        bcost = cost;
                        babs = a;
                        bp = e.i;
                        bq = j;
                    }
                }
                searched++;
                if (bcost >= 0.0 && (bcost <= (double)(nz - 1) * (double)(nz - 1) ||
                                     searched >= kMaxSearch))
                    done = true;
                j = jn;
            }
            if (done) break;
            // rows with nz entries
            for (I i = rb.head[nz]; i >= 0 && !done; i = rb.next[i]) {
                for (I j : row[i]) {
                    if (gone(j)) continue;
                    const double cmax = col_max(j);
                    double a = 0.0;
                    for (const Ent& e : col[j])
                        if (e.i == i) {
                            a = std::abs(e.v);
                            break;
                        }
                    if (a < abstol || a < pivottol * cmax) continue;
                    const double cost = (double)(nz - 1) * (double)(cb.key[j] - 1);
                    if (bcost < 0.0 || cost < bcost || (cost == bcost && a > babs)) {
                        bcost = cost;
                        babs = a;
                        bp = i;
                        bq = j;
                    }
                }
                searched++;
                if (bcost >= 0.0 && (bcost <= (double)nz * (double)(nz - 1) ||
                                     searched >= kMaxSearch))
                    done = true;
            }
            if (nz >= nact_rows && nz >= nact_cols) break;
        }
        if (bp < 0) {
            // no acceptable pivot anywhere: every remaining column is dependent
            for (I j = 0; j < dim; j++)
                if (colpos[j] < 0 && !col_dead[j]) kill_column(j);
            break;
        }

        // ---- eliminate with pivot (bp, bq) ----
        const I p = bp, q = bq, k = (I)prow.size();
        stamp++;
        clist.clear();
        double piv = 0.0;
        for (const Ent& e : col[q]) {
            if (e.i == p) piv = e.v;
        }
        for (const Ent& e : col[q]) {
            if (e.i == p) continue;
            const double l = e.v / piv;
            mult[e.i] = l;
            mark[e.i] = stamp;
            clist.push_back(e.i);
            if (l != 0.0) {
                Li.push_back(e.i);
                Lx.push_back(l);
            }
        }
        Lp.push_back((I)Li.size());
        // the pivot row leaves: its entries go to U, the columns lose them
        for (I j : row[p]) {
            if (j == q || gone(j)) continue;
            std::vector<Ent>& cj = col[j];
            jstamp++;
            double apj = 0.0;
            for (size_t t = 0; t < cj.size();) {
                if (cj[t].i == p) {
                    apj = cj[t].v;
                    cj[t] = cj.back();
                    cj.pop_back();
                    act_nnz--;
                    continue;
                }
                t++;
            }
            if (apj != 0.0) ucol[j].push_back(Ent{k, apj});
            if (!clist.empty() && apj != 0.0) {
                for (Ent& e : cj)
                    if (mark[e.i] == stamp) {
                        e.v -= mult[e.i] * apj;
                        hit[e.i] = jstamp;
                    }
                for (I i : clist) {
                    if (hit[i] == jstamp) continue;
                    const double v = -mult[i] * apj;
                    if (v == 0.0) continue;
                    cj.push_back(Ent{i, v});
                    row[i].push_back(j);
                    rb.Move(i, rb.key[i] + 1);
                    act_nnz++;
                }
            }
            colmax[j] = -1.0;
            cb.Move(j, (I)cj.size());
            if (cj.empty()) kill_column(j);  // nothing left to pivot on: dependent
        }
        ucol[q].push_back(Ent{k, piv});
        // the pivot column leaves
        rowpos[p] = k;
        colpos[q] = k;
        for (I i : clist) drop_from_row(i);
        act_nnz -= (long long)col[q].size();
        std::vector<Ent>().swap(col[q]);
        std::vector<I>().swap(row[p]);
        cb.Remove(q);
        rb.Remove(p);
        prow.push_back(p);
        pcol.push_back(q);
        nact_rows--;
        nact_cols--;
    }
    R.bump_size = nact_cols;
    t_sparse = seconds();

    // ---- dense trailing block ----
    if (nact_cols > 0 && nact_rows > 0) {
        std::vector<I> arow, acol, rloc((size_t)dim, -1);
        for (I i = 0; i < dim; i++)
            if (rowpos[i] < 0) {
                rloc[i] = (I)arow.size();
                arow.push_back(i);
            }
        for (I j = 0; j < dim; j++)
            if (colpos[j] < 0 && !col_dead[j]) acol.push_back(j);
        const I nr = (I)arow.size(), nc = (I)acol.size();
        std::vector<double> D((size_t)nr * (size_t)nc, 0.0);
        for (I c = 0; c < nc; c++) {
            for (const Ent& e : col[acol[c]]) D[(size_t)rloc[e.i] + (size_t)c * nr] = e.v;
            std::vector<Ent>().swap(col[acol[c]]);
        }
        std::vector<I> rows_of, cols_of;
        dense_rows = nr;
        dense_cols = nc;
        const I npiv = DenseLu(D.data(), nr, nc, abstol, &rows_of, &cols_of);
        t_dense = seconds() - t_sparse;
        const I k0 = (I)prow.size();
        for (I t = 0; t < npiv; t++) {
            const I k = k0 + t;
            const I p = arow[rows_of[t]], q = acol[cols_of[t]];
            prow.push_back(p);
            pcol.push_back(q);
            rowpos[p] = k;
            colpos[q] = k;
        }
        // U: column c of the dense block holds U(0..c, c); L: column t holds the multipliers
        // of all rows below t (also of rows that end without a pivot).
        for (I c = 0; c < npiv; c++) {
            const double* dc = D.data() + (size_t)c * nr;
            std::vector<Ent>& uc = ucol[acol[cols_of[c]]];
            for (I t = 0; t <= c; t++)
                if (dc[t] != 0.0 || t == c) uc.push_back(Ent{k0 + t, dc[t]});
        }
        for (I t = 0; t < npiv; t++) {
            const double* dc = D.data() + (size_t)t * nr;
            for (I r = t + 1; r < nr; r++)
                if (dc[r] != 0.0) {
                    Li.push_back(arow[rows_of[r]]);
                    Lx.push_back(dc[r]);
                }
            Lp.push_back((I)Li.size());
        }
        for (I c = npiv; c < nc; c++) {
            const I j = acol[cols_of[c]];
            col_dead[j] = 1;
            deferred.push_back(j);
        }
    }

    // ---- dependent columns go last, paired with leftover rows ----
    {
        std::sort(deferred.begin(), deferred.end());
        size_t d = 0;
        for (I i = 0; i < dim && d < deferred.size(); i++) {
            if (rowpos[i] >= 0) continue;
            const I k = (I)prow.size();
            const I j = deferred[d++];
            ucol[j].clear();
            ucol[j].push_back(Ent{k, 1.0});
            Lp.push_back((I)Li.size());
            prow.push_back(i);
            pcol.push_back(j);
            rowpos[i] = k;
            colpos[j] = k;
            R.dependent_cols.push_back(k);
        }
        assert(d == deferred.size());
        assert((I)prow.size() == dim);
    }

    // ---- assemble: permuted row indices in L, U by pivot column ----
    for (I& i : Li) i = rowpos[i];
    SortColumns(Lp, Li, Lx);
    std::vector<I>&Up = R.Up, &Ui = R.Ui;
    std::vector<double>& Ux = R.Ux;
    Up.assign((size_t)dim + 1, 0);
    for (I k = 0; k < dim; k++) Up[k + 1] = Up[k] + (I)ucol[pcol[k]].size();
    Ui.resize((size_t)Up[dim]);
    Ux.resize((size_t)Up[dim]);
    for (I k = 0; k < dim; k++) {
        I put = Up[k];
        for (const Ent& e : ucol[pcol[k]]) {
            Ui[put] = e.i;
            Ux[put] = e.v;
            put++;
        }
    }
    if (timing)
        std::fprintf(stderr,
                     "[sparse_lu] dim %lld: sparse pivots %lld in %.3f s, dense tail %lld x %lld in "
                     "%.3f s, total %.3f s, nnz(L) %lld nnz(U) %lld, dependent %lld\n",
                     (long long)dim, (long long)(dim - dense_rows), t_sparse, (long long)dense_rows,
                     (long long)dense_cols, t_dense, seconds(), (long long)Lp[dim],
                     (long long)Up[dim], (long long)R.dependent_cols.size());
#ifndef NDEBUG
    for (I k = 0; k < dim; k++) {
        for (I p = Lp[k]; p < Lp[k + 1]; p++) assert(Li[p] > k);
        for (I p = Up[k]; p < Up[k + 1]; p++) assert(Ui[p] <= k);
        assert(Up[k + 1] > Up[k] && Ui[Up[k + 1] - 1] == k);
    }
#endif
}

}  // namespace ipxb200
