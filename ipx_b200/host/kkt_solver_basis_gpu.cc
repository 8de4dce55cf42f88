// GPU drop-in for ipx::KKTSolverBasis: defines the members declared in the
// UNMODIFIED reference header src/kkt_solver_basis.h:24-72 and is linked instead
// of src/kkt_solver_basis.cc.
//
// _Solve (reference src/kkt_solver_basis.cc:75-194) runs on the device in one
// call: the two masked sweeps over the nonbasic columns of AI (right-hand side
// and solution recovery), the Basis::SolveDense steps (as permutation + the
// sparse triangular solves on the factors SplittedNormalMatrix::Prepare loaded)
// and the CR loop on the split operator. The host uploads (a, b) and receives
// (x, y).
//
// _Factorize keeps the reference's sequence (:20-67): scaling factors from the
// iterate, DropPrimal / DropDual, Maxvolume, refactorization, Prepare. Basis
// maintenance (tableau rows, exchanges, Maxvolume) is host work of the
// reference's own Basis / Maxvolume classes.

#include "kkt_solver_basis.h"

#include <cassert>
#include <cmath>
#include <stdexcept>
#include <vector>

#include "gpu_bridge.h"
#include "maxvolume.h"

namespace ipx {

using ipxb200::Check;

namespace {

int64_t InterruptThunk(void* user) {
    return static_cast<const Control*>(user)->InterruptCheck();
}

// A variable's distance to its nearer bound (primal side) or its larger dual (dual side),
// as DropPrimal / DropDual judge it (reference :213-221, :311-319).
struct BoundPair {
    double x, z;
};

BoundPair NearerBound(const Iterate& it, Int j) {
    if (it.xl()[j] <= it.xu()[j]) return {it.xl()[j], it.zl()[j]};
    return {it.xu()[j], it.zu()[j]};
}

BoundPair LargerDual(const Iterate& it, Int j) {
    if (it.zl()[j] >= it.zu()[j]) return {it.xl()[j], it.zl()[j]};
    return {it.xu()[j], it.zu()[j]};
}

// 1 / colscale of the variable at every basis position.
Vector InverseBasicScales(const Basis& basis, const Vector& colscale, Int m) {
    Vector inv(m);
    for (Int p = 0; p < m; p++) {
        inv[p] = 1.0 / colscale[basis[p]];
        assert(std::isfinite(inv[p]) && inv[p] >= 0.0);
    }
    return inv;
}

}  // namespace

KKTSolverBasis::KKTSolverBasis(const Control& control, Basis& basis)
    : control_(control), model_(basis.model()), basis_(basis), splitted_normal_matrix_(model_) {
    colscale_.resize(model_.rows() + model_.cols());
}

void KKTSolverBasis::_Factorize(Iterate* iterate, Info* info) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    info->errflag = 0;
    factorized_ = false;
    iter_ = 0;
    basis_changes_ = 0;

    for (Int j = 0; j < n + m; j++) colscale_[j] = iterate->ScalingFactor(j);

    // Degenerate variables are removed only while the model looks feasible
    // (reference :31-43).
    if (iterate->pobjective() >= iterate->dobjective()) {
        DropPrimal(iterate, info);
        if (info->errflag) return;
        DropDual(iterate, info);
        if (info->errflag) return;
    }

    Maxvolume maxvol(control_);
    info->errflag = control_.update_heuristic() == 0
                        ? maxvol.RunSequential(&colscale_[0], basis_)
                        : maxvol.RunHeuristic(&colscale_[0], basis_);
    info->updates_ipm += maxvol.updates();
    info->time_maxvol += maxvol.time();
    basis_changes_ += maxvol.updates();
    if (info->errflag) return;

    if (!basis_.FactorizationIsFresh()) {
        info->errflag = basis_.Factorize();
        if (info->errflag) return;
    }
    // Loads L, U (column-scaled), the permutations and the masked nonbasic scales.
    splitted_normal_matrix_.Prepare(basis_, &colscale_[0]);

    // What _Solve needs on top: which variable sits at each pivot position and its scale.
    const Int* colperm = splitted_normal_matrix_.colperm();
    std::vector<Int> basic_var(m);
    Vector basic_scale(m);
    for (Int k = 0; k < m; k++) {
        const Int j = basis_[colperm[k]];
        basic_var[k] = j;
        basic_scale[k] = basis_.StatusOf(j) == Basis::BASIC ? colscale_[j] : 1.0;
    }
    if (m > 0) {
        const ipxb200::ContextRef ref = ipxb200::ContextFor(model_);
        Check(ipxgpu_kktbasis_prepare(ref.ctx, basic_var.data(), colperm, &basic_scale[0]));
    }
    factorized_ = true;
}

void KKTSolverBasis::_Solve(const Vector& a, const Vector& b, double tol, Vector& x, Vector& y,
                            Info* info) {
    const Int m = model_.rows();
    info->errflag = 0;
    assert(factorized_);
    if (m == 0) {
        x = 0.0;
        return;
    }
    const ipxb200::ContextRef ref = ipxb200::CurrentContext(model_);
    if (!ref.ctx) throw std::logic_error("KKTSolverBasis: device context lost; call Factorize again");
    ipxgpu_cr_result res{};
    if (ipxb200::OperatorRecord* op = ipxb200::FindRecord(&splitted_normal_matrix_))
        ipxb200::EnsurePrimed(*op, &splitted_normal_matrix_);  // throws if another basis took over
    const int rc = ipxgpu_kktbasis_solve(ref.ctx, &a[0], &b[0], tol, maxiter_, &x[0], &y[0], &res,
                                         InterruptThunk, const_cast<Control*>(&control_));
    if (rc == IPXGPU_ERR_STATE)
        throw std::logic_error("KKTSolverBasis: device state lost; call Factorize again");
    Check(rc);
    if (res.errflag == IPX_ERROR_cr_iter_limit)
        control_.Debug(3) << " CR method not converged in " << res.iter << " iterations."
                          << " residual = " << sci2(res.resnorm) << ','
                          << " tolerance = " << sci2(tol) << '\n';
    info->errflag = res.errflag;
    info->kktiter2 += res.iter;
    info->time_cr2 += res.time;
    info->time_cr2_NNt += res.time_NNt;
    info->time_cr2_B += res.time_B;
    info->time_cr2_Bt += res.time_Bt;
    iter_ += res.iter;
}

// Basic variables close to a bound either leave the basis (if some nonbasic
// variable increases the volume enough) or become "implied" (reference :196-291).
void KKTSolverBasis::DropPrimal(Iterate* iterate, Info* info) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    const double drop_tol = control_.ipm_drop_primal();
    const double kVolumeTol = 2.0;
    info->errflag = 0;

    std::vector<Int> todo;
    for (Int p = 0; p < m; p++) {
        const Int jb = basis_[p];
        if (basis_.StatusOf(jb) != Basis::BASIC) continue;  // free variables stay
        const BoundPair near = NearerBound(*iterate, jb);
        if (near.x < 0.01 * near.z && near.x <= drop_tol) todo.push_back(jb);
    }
    if (todo.empty()) return;

    Vector inv_scale = InverseBasicScales(basis_, colscale_, m);
    IndexedVector btran(m), row(n + m);
    while (!todo.empty()) {
        const Int jb = todo.back();
        const Int p = basis_.PositionOf(jb);
        assert(p >= 0);
        basis_.TableauRow(jb, btran, row, true);
        // entering candidate with the largest scaled pivot above the volume tolerance
        const double s = inv_scale[p];
        Int jenter = -1;
        double best = kVolumeTol;
        auto consider = [&](Int j, double pivot) {
            const double size = std::abs(pivot);
            if (size <= kPivotZeroTol) return;
            const double volume = size * colscale_[j] * s;
            if (volume > best) {
                best = volume;
                jenter = j;
            }
        };
        for_each_nonzero(row, consider);
        if (jenter < 0) {
            // nothing worth pivoting on: the variable becomes implied at the bound whose dual
            // pull is stronger
            const Vector &xl = iterate->xl(), &xu = iterate->xu();
            const Vector &zl = iterate->zl(), &zu = iterate->zu();
            if (zl[jb] / xl[jb] > zu[jb] / xu[jb]) iterate->make_implied_lb(jb);
            else iterate->make_implied_ub(jb);
            basis_.FreeBasicVariable(jb);
            inv_scale[p] = 0.0;
            colscale_[jb] = INFINITY;
            info->primal_dropped++;
            todo.pop_back();
            continue;
        }
        const double pivot = row[jenter];
        if (std::abs(pivot) < 1e-3)
            control_.Debug(3) << " |pivot| = " << sci2(std::abs(pivot))
                              << " (primal basic variable close to bound)\n";
        assert(basis_.StatusOf(jenter) == Basis::NONBASIC);
        bool exchanged = false;
        info->errflag = basis_.ExchangeIfStable(jb, jenter, pivot, 1, &exchanged);
        if (info->errflag) return;
        if (!exchanged) continue;  // unstable factorization update: refactorized, try again
        inv_scale[p] = 1.0 / colscale_[jenter];
        assert(std::isfinite(inv_scale[p]) && inv_scale[p] >= 0.0);
        info->updates_ipm++;
        basis_changes_++;
        todo.pop_back();
    }
}

// Nonbasic variables whose dual is close to zero either enter the basis or are
// fixed at their current value (reference :293-388).
void KKTSolverBasis::DropDual(Iterate* iterate, Info* info) {
    const Int m = model_.rows();
    const Int n = model_.cols();
    const double drop_tol = control_.ipm_drop_dual();
    const double kVolumeTol = 2.0;
    info->errflag = 0;

    std::vector<Int> todo;
    for (Int jn = 0; jn < n + m; jn++) {
        if (basis_.StatusOf(jn) != Basis::NONBASIC) continue;
        const BoundPair big = LargerDual(*iterate, jn);
        if (big.z < 0.01 * big.x && big.z <= drop_tol) todo.push_back(jn);
    }
    if (todo.empty()) return;

    Vector inv_scale = InverseBasicScales(basis_, colscale_, m);
    IndexedVector ftran(m);
    while (!todo.empty()) {
        const Int jn = todo.back();
        basis_.SolveForUpdate(jn, ftran);
        // leaving position with the largest scaled pivot above the volume tolerance
        const double s = colscale_[jn];
        Int pleave = -1;
        double best = kVolumeTol;
        auto consider = [&](Int p, double pivot) {
            const double size = std::abs(pivot);
            if (size <= kPivotZeroTol) return;
            const double volume = size * inv_scale[p] * s;
            if (volume > best) {
                best = volume;
                pleave = p;
            }
        };
        for_each_nonzero(ftran, consider);
        if (pleave < 0) {
            iterate->make_fixed(jn);
            basis_.FixNonbasicVariable(jn);
            colscale_[jn] = 0.0;
            info->dual_dropped++;
            todo.pop_back();
            continue;
        }
        const double pivot = ftran[pleave];
        if (std::abs(pivot) < 1e-3)
            control_.Debug(3) << " |pivot| = " << sci2(std::abs(pivot))
                              << " (dual nonbasic variable close to zero)\n";
        const Int jb = basis_[pleave];
        assert(basis_.StatusOf(jb) == Basis::BASIC);
        bool exchanged = false;
        info->errflag = basis_.ExchangeIfStable(jb, jn, pivot, -1, &exchanged);
        if (info->errflag) return;
        if (!exchanged) continue;
        inv_scale[pleave] = 1.0 / colscale_[jn];
        assert(std::isfinite(inv_scale[pleave]) && inv_scale[pleave] >= 0.0);
        info->updates_ipm++;
        basis_changes_++;
        todo.pop_back();
    }
}

}  // namespace ipx
