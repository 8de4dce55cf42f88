// Host LU provider for IPX builds without BASICLU.
//
// The reference links the external BASICLU library through two adapter TUs
// (src/basiclu_kernel.cc, src/basiclu_wrapper.cc). BASICLU is not vendored, so
// this TU defines the members of the two adapter classes declared in the
// UNMODIFIED reference headers (src/basiclu_kernel.h:10-18,
// src/basiclu_wrapper.h:11-47) on top of ipxb200::SparseLuFactorize:
//
//  * BasicLuKernel::_Factorize honours the LuFactorization contract
//    (src/lu_factorization.h:22-59).
//  * BasicLu (the lu_kernel<=0 default, src/basis.cc:24-29) delegates every
//    LuUpdate virtual to the reference's own ForrestTomlin updater
//    (src/forrest_tomlin.h) running on that kernel. BasicLu's layout is fixed
//    by the header, so the delegate lives in a side table keyed by `this`.
//
// It is linked into BOTH the CPU reference build of the test suite and the GPU
// drop-in build, so the two arms see identical L, U and permutations.

#include <memory>
#include <mutex>
#include <unordered_map>

#include "basiclu_kernel.h"
#include "basiclu_wrapper.h"
#include "forrest_tomlin.h"
#include "sparse_lu.h"

namespace ipx {

void BasicLuKernel::_Factorize(Int dim, const Int* Bbegin, const Int* Bend,
                               const Int* Bi, const double* Bx, double pivottol,
                               bool strict_abs_pivottol, SparseMatrix* L,
                               SparseMatrix* U, std::vector<Int>* rowperm,
                               std::vector<Int>* colperm,
                               std::vector<Int>* dependent_cols) {
    if (dim == 0) {
        L->clear();
        U->clear();
        rowperm->clear();
        colperm->clear();
        dependent_cols->clear();
        return;
    }
    static_assert(sizeof(Int) == sizeof(int64_t), "IPX Int must be 64-bit");
    ipxb200::SparseLuResult lu;
    // BASICLU's default absolute pivot tolerance is 1e-14
    // (src/basiclu_wrapper.cc:55); strict mode uses kLuDependencyTol.
    const double abstol = strict_abs_pivottol ? kLuDependencyTol : 1e-14;
    ipxb200::SparseLuFactorize(dim, Bbegin, Bend, Bi, Bx, pivottol, abstol,
                               &lu);
    const Int lnz = lu.Lp[dim], unz = lu.Up[dim];
    L->resize(dim, dim, lnz);
    U->resize(dim, dim, unz);
    std::copy(lu.Lp.begin(), lu.Lp.end(), L->colptr());
    std::copy(lu.Li.begin(), lu.Li.end(), L->rowidx());
    std::copy(lu.Lx.begin(), lu.Lx.end(), L->values());
    std::copy(lu.Up.begin(), lu.Up.end(), U->colptr());
    std::copy(lu.Ui.begin(), lu.Ui.end(), U->rowidx());
    std::copy(lu.Ux.begin(), lu.Ux.end(), U->values());
    rowperm->assign(lu.rowperm.begin(), lu.rowperm.end());
    colperm->assign(lu.colperm.begin(), lu.colperm.end());
    dependent_cols->assign(lu.dependent_cols.begin(), lu.dependent_cols.end());
}

namespace {
using Table = std::unordered_map<const BasicLu*, std::unique_ptr<ForrestTomlin>>;
Table& table() {
    static Table t;
    return t;
}
std::mutex& table_mutex() {
    static std::mutex m;
    return m;
}
ForrestTomlin& delegate(const BasicLu* self) {
    std::lock_guard<std::mutex> lock(table_mutex());
    return *table().at(self);
}
}  // namespace

BasicLu::BasicLu(const Control& control, Int dim) : control_(control) {
    dim_ = dim;
    std::unique_ptr<LuFactorization> kernel(new BasicLuKernel);
    std::unique_ptr<ForrestTomlin> ft(new ForrestTomlin(control, dim, kernel));
    std::lock_guard<std::mutex> lock(table_mutex());
    // An entry left by a destroyed BasicLu at the same address is replaced
    // (BasicLu has a defaulted destructor, so there is no hook to erase it).
    table()[this] = std::move(ft);
}

Int BasicLu::_Factorize(const Int* Bbegin, const Int* Bend, const Int* Bi,
                        const double* Bx, bool strict_abs_pivottol) {
    ForrestTomlin& ft = delegate(this);
    Int ret = ft.Factorize(Bbegin, Bend, Bi, Bx, strict_abs_pivottol);
    fill_factor_ = ft.fill_factor();
    return ret;
}

void BasicLu::_GetFactors(SparseMatrix* L, SparseMatrix* U, Int* rowperm,
                          Int* colperm, std::vector<Int>* dependent_cols) {
    delegate(this).GetFactors(L, U, rowperm, colperm, dependent_cols);
}

void BasicLu::_SolveDense(const Vector& rhs, Vector& lhs, char trans) {
    delegate(this).SolveDense(rhs, lhs, trans);
}

void BasicLu::_FtranForUpdate(Int nz, const Int* bi, const double* bx) {
    delegate(this).FtranForUpdate(nz, bi, bx);
}

void BasicLu::_FtranForUpdate(Int nz, const Int* bi, const double* bx,
                              IndexedVector& lhs) {
    delegate(this).FtranForUpdate(nz, bi, bx, lhs);
}

void BasicLu::_BtranForUpdate(Int j) { delegate(this).BtranForUpdate(j); }

void BasicLu::_BtranForUpdate(Int j, IndexedVector& lhs) {
    delegate(this).BtranForUpdate(j, lhs);
}

Int BasicLu::_Update(double pivot) { return delegate(this).Update(pivot); }

bool BasicLu::_NeedFreshFactorization() {
    return delegate(this).NeedFreshFactorization();
}

double BasicLu::_fill_factor() const { return fill_factor_; }

double BasicLu::_pivottol() const { return pivottol_; }

void BasicLu::_pivottol(double new_pivottol) {
    pivottol_ = new_pivottol;
    delegate(this).pivottol(new_pivottol);
}

void BasicLu::Reallocate() {}

}  // namespace ipx
