// Host sparse LU used where the reference links the (un-vendored) BASICLU
// library. Stand-alone: plain int64/double arrays, no IPX types.
//
// Contract (reference src/lu_factorization.h:22-59, src/lu_update.h:43-60):
//   B[rowperm, colperm] = (L + I) * U
// with L strictly lower triangular (unit diagonal not stored), U upper
// triangular with the diagonal entry LAST in each column (indices sorted),
// dependent columns replaced by unit columns and listed in dependent_cols.
//
// Algorithm: column-singleton pass, row-singleton pass (both fill-free), then a
// left-looking Gilbert-Peierls factorization of the remaining bump with
// threshold partial pivoting and a sparsest-row tie break.

#ifndef IPXB200_SPARSE_LU_H_
#define IPXB200_SPARSE_LU_H_

#include <cstdint>
#include <vector>

namespace ipxb200 {

struct SparseLuResult {
    std::vector<int64_t> Lp, Li;  // CSC, dim columns, strict lower, permuted
    std::vector<double> Lx;
    std::vector<int64_t> Up, Ui;  // CSC, dim columns, upper, diagonal last
    std::vector<double> Ux;
    std::vector<int64_t> rowperm, colperm;
    std::vector<int64_t> dependent_cols;  // positions k in the pivot sequence
    int64_t num_col_singletons{0};
    int64_t num_row_singletons{0};
    int64_t bump_size{0};
};

// @pivottol  relative threshold in (0,1]
// @abstol    absolute pivot tolerance; a column whose eligible entries are all
//            smaller is declared dependent
void SparseLuFactorize(int64_t dim, const int64_t* Bbegin, const int64_t* Bend,
                       const int64_t* Bi, const double* Bx, double pivottol,
                       double abstol, SparseLuResult* out);

}  // namespace ipxb200

#endif  // IPXB200_SPARSE_LU_H_
