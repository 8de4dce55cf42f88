// GPU drop-in for ipx::Maxvolume: defines the members declared in the UNMODIFIED reference
// header src/maxvolume.h:11-52 and is linked instead of src/maxvolume.cc.
//
// RunHeuristic (reference src/maxvolume.cc:109-152 with its Driver, :202-320) keeps one weight
// per column of AI and, per basis update, scans all of them for the largest two, forms the
// tableau row of the leaving variable over all nonbasic columns and folds it into the weights -
// three passes over n+m columns around two solves with the basis factorization. Here the
// weights and scaling factors of the columns live on the device for the whole run
// (ipxgpu_maxvol_*): the host keeps the pivoting (Basis::SolveForUpdate, ExchangeIfStable, the
// scaled pivot search on the FTRAN column), sends btran (m doubles) per update and receives the
// next two candidates. The tableau row itself is never materialised: only its entry in the
// entering column is needed by value (for the stability test of the exchange), and that one is
// a single column dot product.
//
// RunSequential (:14-95) touches one FTRAN column per candidate and no column sweep; it is the
// reference's procedure on the host's Basis object.

#include "maxvolume.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <vector>

#include "gpu_bridge.h"
#include "timer.h"
#include "utils.h"

namespace ipx {

using ipxb200::Check;

Maxvolume::Maxvolume(const Control& control) : control_(control) {}

// Scaling state of one run; Slice is declared (not defined) by the reference header.
struct Maxvolume::Slice {
    Slice(Int m, Int n) : colscale(n + m), invscale_basic(m), tblrow_used(m), lhs(m), work(m) {}
    Vector colscale;        // host mirror of the device copy (0: basic, fixed or given up)
    Vector invscale_basic;  // 1 / scaling factor of the variable at each basis position
    std::vector<bool> tblrow_used;
    IndexedVector lhs;      // FTRAN column, then the BTRAN row of an update
    Vector work;
    ipxgpu_ctx* ctx = nullptr;
    bool scale_resident = false;  // the device holds colscale (set by the first Driver call)
};

namespace {

// 1 / colscale[j] for the BASIC variable at every position, 0 for BASIC_FREE ones (which are
// thereby never chosen to leave, :24-34, :121-128).
void FillInverseBasicScales(const Basis& basis, const double* colscale, Vector& inv) {
    const Int m = basis.model().rows();
    for (Int p = 0; p < m; p++) {
        const Int j = basis[p];
        if (basis.StatusOf(j) != Basis::BASIC) continue;
        inv[p] = colscale ? 1.0 / colscale[j] : 1.0;
        assert(std::isfinite(inv[p]));
    }
}

// AI[:,j]' * x in storage order.
double ColumnDot(const SparseMatrix& AI, Int j, const double* x) {
    double sum = 0.0;
    for (Int p = AI.begin(j); p < AI.end(j); p++) sum += x[AI.index(p)] * AI.value(p);
    return sum;
}

}  // namespace

Int Maxvolume::RunSequential(const double* colscale, Basis& basis) {
    const Int m = basis.model().rows();
    const Int n = basis.model().cols();
    const Int maxpasses = control_.maxpasses();
    const double volumetol = std::max(control_.volume_tol(), 1.0);
    IndexedVector ftran(m);
    Vector invscale_basic(m);
    FillInverseBasicScales(basis, colscale, invscale_basic);
    Timer timer;
    Int errflag = 0;
    Reset();

    for (bool again = true; again && (passes_ < maxpasses || maxpasses < 0);) {
        tblnnz_ = 0;
        tblmax_ = 0.0;
        frobnorm_squared_ = 0.0;
        Int updates_in_pass = 0;
        // columns by ascending scaling factor: the back is tried first
        std::vector<Int> order = Sortperm(n + m, colscale, false);
        while (!order.empty()) {
            const Int j = order.back();
            const double dj = colscale ? colscale[j] : 1.0;
            if (dj == 0.0) break;  // only zero factors remain
            if (basis.StatusOf(j) != Basis::NONBASIC) {
                order.pop_back();
                continue;
            }
            if ((errflag = control_.InterruptCheck()) != 0) break;
            basis.SolveForUpdate(j, ftran);
            // largest scaled entry of the tableau column and the statistics of the pass
            Int pmax = -1;
            double vmax = 0.0;
            auto visit = [&](Int p, double x) {
                const double v = std::abs(x) * invscale_basic[p] * dj;
                if (v > vmax) {
                    vmax = v;
                    pmax = p;
                }
                tblnnz_ += v != 0;
                frobnorm_squared_ += v * v;
            };
            for_each_nonzero(ftran, visit);
            tblmax_ = std::max(tblmax_, vmax);
            if (vmax <= volumetol) {
                skipped_++;
                order.pop_back();
                continue;
            }
            const Int jb = basis[pmax];
            assert(basis.StatusOf(jb) == Basis::BASIC);
            bool exchanged = false;
            errflag = basis.ExchangeIfStable(jb, j, ftran[pmax], -1, &exchanged);
            if (errflag) break;
            if (!exchanged) continue;  // refactorized: same column once more
            invscale_basic[pmax] = 1.0 / dj;
            updates_in_pass++;
            volinc_ += std::log2(vmax);
            order.pop_back();
        }
        updates_ += updates_in_pass;
        passes_++;
        again = updates_in_pass > 0 && errflag == 0;
    }
    time_ = timer.Elapsed();
    return errflag;
}

Int Maxvolume::RunHeuristic(const double* colscale, Basis& basis) {
    const Model& model = basis.model();
    const Int m = model.rows();
    const Int n = model.cols();
    Timer timer;
    Reset();
    Slice slice(m, n);

    const Int num_slices =
        std::min<Int>(5 + std::max((long)(m / control_.rows_per_slice()), 0l), m);
    FillInverseBasicScales(basis, colscale, slice.invscale_basic);
    // Working copy of the scaling factors of the NONBASIC columns; a column the heuristic gives
    // up on is zeroed there and not looked at again (:130-135).
    for (Int j = 0; j < n + m; j++)
        if (basis.StatusOf(j) == Basis::NONBASIC) slice.colscale[j] = colscale ? colscale[j] : 1.0;

    // Row slices of the tableau matrix: every Driver call has one row of each slice in use,
    // dealt out round robin along the order of the inverse scaling factors (:137-147).
    const std::vector<Int> perm = Sortperm(m, &slice.invscale_basic[0], false);
    Int errflag = 0;
    try {
        for (Int s = 0; s < num_slices && errflag == 0; s++) {
            for (Int i = 0; i < m; i++) slice.tblrow_used[perm[i]] = i % num_slices == s;
            errflag = Driver(basis, slice);
        }
    } catch (...) {
        if (slice.ctx) ipxgpu_maxvol_release(slice.ctx);
        throw;
    }
    if (slice.ctx) Check(ipxgpu_maxvol_release(slice.ctx));
    time_ = timer.Elapsed();
    passes_ = -1;
    slices_ = num_slices;
    return errflag;
}

Int Maxvolume::updates() const { return updates_; }
Int Maxvolume::skipped() const { return skipped_; }
Int Maxvolume::passes() const { return passes_; }
Int Maxvolume::slices() const { return slices_; }
double Maxvolume::volinc() const { return volinc_; }
double Maxvolume::time() const { return time_; }
Int Maxvolume::tblnnz() const { return tblnnz_; }
double Maxvolume::tblmax() const { return tblmax_; }
double Maxvolume::frobnorm_squared() const { return frobnorm_squared_; }

void Maxvolume::Reset() {
    updates_ = skipped_ = passes_ = slices_ = tblnnz_ = 0;
    volinc_ = time_ = tblmax_ = frobnorm_squared_ = 0.0;
}

Int Maxvolume::Driver(Basis& basis, Slice& slice) {
    const Model& model = basis.model();
    const Int m = model.rows();
    const SparseMatrix& AI = model.AI();
    const double volumetol = std::max(control_.volume_tol(), 1.0);
    const Int maxskip = control_.maxskip_updates();
    Vector& colscale = slice.colscale;
    Vector& invscale_basic = slice.invscale_basic;
    const std::vector<bool>& used = slice.tblrow_used;
    IndexedVector& lhs = slice.lhs;
    if (!slice.ctx) slice.ctx = ipxb200::ContextFor(model).ctx;
    ipxgpu_ctx* ctx = slice.ctx;

    // Column weights (:220-231): the rows in use, scaled, through inverse(B'); then one sweep
    // over the columns on the device, which also finds the first two candidates.
    for (Int p = 0; p < m; p++) slice.work[p] = used[p] ? invscale_basic[p] : 0.0;
    basis.SolveDense(slice.work, slice.work, 'T');
    ipxgpu_maxvol_top top{};
    Check(ipxgpu_maxvol_weights(ctx, slice.scale_resident ? nullptr : &colscale[0],
                                &slice.work[0], &top));
    slice.scale_resident = true;

    // Candidates as FindLargest leaves them: the largest weight at the back. Their weights do
    // not change while they wait (only an update changes weights, and it clears the list).
    struct Candidate {
        Int j;
        double weight;
    };
    std::vector<Candidate> candidates;
    auto refill = [&](const ipxgpu_maxvol_top& t) {
        candidates.clear();
        candidates.push_back({(Int)t.jmax2, t.wmax2});
        candidates.push_back({(Int)t.jmax, t.wmax});
    };
    refill(top);

    Int errflag = 0, skipped = 0;
    while (true) {
        if (candidates.empty()) {
            Check(ipxgpu_maxvol_skip(ctx, -1, &top));  // search only
            refill(top);
        }
        const Int jn = candidates.back().j;
        if (candidates.back().weight == 0.0) break;
        assert(basis.StatusOf(jn) == Basis::NONBASIC && colscale[jn] > 0.0);
        if ((errflag = control_.InterruptCheck()) != 0) break;

        // Largest scaled entry of the tableau column of jn (:249-252).
        basis.SolveForUpdate(jn, lhs);
        const Int pmax = ScaleFtran(colscale[jn], invscale_basic, lhs);
        const double vmax = std::abs(lhs[pmax]);
        if (vmax <= volumetol) {  // the exchange would not gain enough volume: give jn up
            colscale[jn] = 0.0;
            Check(ipxgpu_maxvol_skip(ctx, jn, nullptr));
            candidates.pop_back();
            if (++skipped > maxskip && maxskip >= 0) break;
            continue;
        }
        // The weight of jn once more, from the scaled column (:267-274).
        double weight_recomp = 0.0;
        auto add_used = [&](Int p, double x) {
            if (used[p]) weight_recomp += x;
        };
        for_each_nonzero(lhs, add_used);
        assert(std::isfinite(weight_recomp));

        // Row pmax of inverse(B); the tableau row's entry in column jn is the pivot (:277-283).
        const Int jb = basis[pmax];
        basis.SolveForUpdate(jb, lhs);
        const double pivot = ColumnDot(AI, jn, lhs.elements());
        if (std::abs(pivot) < 1e-3)
            control_.Debug(3) << " |pivot| " << sci2(std::abs(pivot)) << "(maxvolume)\n";
        bool exchanged = false;
        errflag = basis.ExchangeIfStable(jb, jn, pivot, 0, &exchanged);
        if (errflag) break;
        if (!exchanged) continue;  // refactorized: same column once more
        updates_++;
        volinc_ += std::log2(vmax);

        // jb takes over the scaling factor slot of jn and vice versa (:291-299) ...
        const double dn = colscale[jn];
        const double dbinv = invscale_basic[pmax];
        assert(colscale[jb] == 0.0);
        colscale[jb] = 1.0 / dbinv;
        invscale_basic[pmax] = 1.0 / dn;
        colscale[jn] = 0.0;
        assert(std::isfinite(colscale[jb]) && std::isfinite(invscale_basic[pmax]));
        // ... and every weight moves along the tableau row (:301-308), on the device.
        const double alpha = (used[pmax] - weight_recomp) / (dn * pivot);
        assert(std::isfinite(alpha));
        Check(ipxgpu_maxvol_update(ctx, lhs.elements(), alpha, jb, colscale[jb],
                                   used[pmax] + alpha / dbinv, jn, &top));
        refill(top);
    }
    skipped_ += skipped;
    return errflag;
}

Int Maxvolume::ScaleFtran(double colscale_jn, const Vector& invscale_basic,
                          IndexedVector& ftran) {
    // Scales the column in place and returns the position of its largest scaled entry among
    // those whose unscaled value is a usable pivot.
    Int pmax = 0;
    double vmax = 0.0;
    auto scale = [&](Int p, double& entry) {
        const double raw = entry;
        entry = raw * colscale_jn * invscale_basic[p];
        const double v = std::abs(entry);
        if (v > vmax && std::abs(raw) > kPivotZeroTol) {
            vmax = v;
            pmax = p;
        }
    };
    for_each_nonzero(ftran, scale);
    return pmax;
}

}  // namespace ipx
