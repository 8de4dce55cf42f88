// Vector kernels of the (preconditioned) Conjugate Residuals loop.
//
// The loop of reference src/conjugate_residuals.cc:14-88 / :90-213 is cut at
// its global reductions into three kernels; the scalar logic (alpha, beta,
// termination tests, in the reference's order) runs in the CTA that finishes a
// kernel's reduction, so no scalar travels to the host inside the loop:
//
//   cr_update     y += a p; r -= a Cp; [s -= a q];  resnorm = max|resscale.*r|
//   <C.Apply>     Cs = C s (or C r), cdotnew -> beta, cdot, iter++   (after_apply)
//   cr_direction  p = s + b p; Cp = Cs + b Cp; [q = Cp./diag]; pdot;
//                 [every 5th pass: s = r./diag, monotonicity test];
//                 then the top-of-loop tests and alpha for the next pass.
//
// Every kernel returns at once when st->done is set, so the host may enqueue
// passes ahead of the device without overshooting the iterate.
#pragma once

#include "common.cuh"
#include "smw.cuh"

namespace ipxgpu {

struct CrVectors {
    int m;
    double* y;        // lhs
    double* r;        // residual
    double* s;        // preconditioned residual (PCR only)
    double* p;        // step
    double* Cp;       // Cstep
    double* Cs;       // C*sresidual (PCR) or C*residual (CR); m+1 entries
    double* q;        // P*Cstep (PCR only)
    const double* diag;      // PCR only
    const double* resscale;  // may be nullptr
};

// r = rhs - Cy (Cy == nullptr: r = rhs); s = r./diag; rsdot; resnorm; p = Cp = 0.
// Reference src/conjugate_residuals.cc:33-39, :118-125 and the first residual
// norm of :44-49 / :131-136.
__global__ void __launch_bounds__(kBlock)
cr_init_kernel(CrVectors v, const double* __restrict__ rhs, const double* __restrict__ Cy,
               Reduce red, CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    double rs = 0.0, mx = 0.0;
    const bool precond = st->precond != 0;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < v.m; i += gridDim.x * kBlock) {
        const double ri = Cy ? rhs[i] - Cy[i] : rhs[i];
        v.r[i] = ri;
        v.p[i] = 0.0;
        v.Cp[i] = 0.0;
        if (precond) {
            const double si = ri / v.diag[i];
            v.s[i] = si;
            rs += __dmul_rn(si, ri);
        }
        const double sc = v.resscale ? __dmul_rn(v.resscale[i], ri) : ri;
        mx = fmax(mx, fabs(sc));
    }
    const double bs = block_sum(rs, s_red);
    const double bm = block_max(mx, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, bs, 0.0, bm, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0) {
        st->rsdot_prev = ts;
        st->resnorm = tm;
        stamp(st, precond ? kSlotPre : kSlotVec);
    }
}

// Reference src/conjugate_residuals.cc:72-73 / :173-175, and the residual norm
// tested at the top of the next pass (:44-49 / :131-136).
__global__ void __launch_bounds__(kBlock)
cr_update_kernel(CrVectors v, Reduce red, CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st->done) return;
    const double alpha = st->alpha;
    const bool precond = st->precond != 0;
    double mx = 0.0;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < v.m; i += gridDim.x * kBlock) {
        v.y[i] = v.y[i] + __dmul_rn(alpha, v.p[i]);
        const double ri = v.r[i] - __dmul_rn(alpha, v.Cp[i]);
        v.r[i] = ri;
        if (precond) v.s[i] = v.s[i] - __dmul_rn(alpha, v.q[i]);
        const double sc = v.resscale ? __dmul_rn(v.resscale[i], ri) : ri;
        mx = fmax(mx, fabs(sc));
    }
    const double bm = block_max(mx, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, 0.0, 0.0, bm, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0) {
        st->resnorm = tm;
        stamp(st, kSlotVec);
    }
}

// The monotonicity test of the recomputed preconditioned residual (reference
// src/conjugate_residuals.cc:187-207) and the tests at the top of the loop (:50-71 / :137-172),
// in the reference's order; alpha for the next pass. One thread, after the reductions.
__device__ __forceinline__ void cr_direction_tests(CrState* st, bool precond, bool recompute,
                                                   long long iter, double tpd, double trs) {
    int done = 0, err = 0;
    if (recompute) {  // :187-207
        if (trs >= st->rsdot_prev) {
            err = 204;
            done = 1;
        } else {
            st->rsdot_prev = trs;
        }
    }
    if (!done) {
        const double resnorm = st->resnorm;
        if (st->hist && iter < st->hist_cap) st->hist[iter] = resnorm;
        st->pdot = tpd;
        if (resnorm <= st->tol) {
            done = 1;
        } else if (iter == st->maxiter) {
            err = 201;
            done = 1;
        } else if (st->cdot <= 0.0) {
            err = 202;
            done = 1;
        } else if (precond && tpd <= 0.0) {
            err = 203;
            done = 1;
        } else {
            const double alpha = st->cdot / tpd;
            if (!isfinite(alpha)) {
                err = 205;
                done = 1;
            }
            st->alpha = alpha;
        }
    }
    st->errflag = err;
    st->done = done;
    stamp(st, precond ? kSlotPre : kSlotVec);
    publish(st);
}

// Reference src/conjugate_residuals.cc:79-80 / :181-207 followed by the tests
// at the top of the loop (:50-71 / :137-172), in the reference's order.
// vec_only: p and Cp only - the preconditioner has a dense-column part and is applied by the
// kernels that follow (smw_gather, smw_solve, cr_direction_smw_kernel).
__global__ void __launch_bounds__(kBlock)
cr_direction_kernel(CrVectors v, Reduce red, CrState* st, int vec_only) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st->done) return;
    const double beta = st->beta;
    const bool precond = st->precond != 0;
    const long long iter = st->iter;
    const bool recompute = precond && iter > 0 && (iter % 5 == 0);
    const double* sv = precond ? v.s : v.r;
    double pd = 0.0, rs = 0.0;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < v.m; i += gridDim.x * kBlock) {
        const double pn = sv[i] + __dmul_rn(beta, v.p[i]);
        const double cpn = v.Cs[i] + __dmul_rn(beta, v.Cp[i]);
        v.p[i] = pn;
        v.Cp[i] = cpn;
        if (vec_only) continue;
        if (precond) {
            const double d = v.diag[i];
            const double qi = cpn / d;
            v.q[i] = qi;
            pd += __dmul_rn(qi, cpn);
            if (recompute) {
                const double ri = v.r[i];
                const double sn = ri / d;
                v.s[i] = sn;
                rs += __dmul_rn(sn, ri);
            }
        } else {
            pd += __dmul_rn(cpn, cpn);
        }
    }
    if (vec_only) return;
    const double bpd = block_sum(pd, s_red);
    const double brs = block_sum(rs, s_red);
    double tpd, trs, tm;
    if (!grid_reduce(red, bpd, brs, 0.0, s_red, &s_flag, &tpd, &trs, &tm)) return;
    if (threadIdx.x != 0) return;
    cr_direction_tests(st, precond, recompute, iter, tpd, trs);
}

// Second half of the direction stage when the preconditioner has a dense-column part:
// q = inv(P) Cp (and s = inv(P) r on recompute iterations) from the solved z, the two dots,
// then the tests.
__global__ void __launch_bounds__(kBlock)
cr_direction_smw_kernel(CrVectors v, SmwDev S, Reduce red, CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    if (st->done) return;
    const long long iter = st->iter;
    const bool recompute = iter > 0 && (iter % 5 == 0);
    double pd = 0.0, rs = 0.0;
    smw_rows(S, v.diag, v.Cp, v.r, recompute, [&](int i, double qi, double sn) {
        v.q[i] = qi;
        pd += __dmul_rn(qi, v.Cp[i]);
        if (recompute) {
            v.s[i] = sn;
            rs += __dmul_rn(sn, v.r[i]);
        }
    });
    const double bpd = block_sum(pd, s_red);
    const double brs = block_sum(rs, s_red);
    double tpd, trs, tm;
    if (!grid_reduce(red, bpd, brs, 0.0, s_red, &s_flag, &tpd, &trs, &tm)) return;
    if (threadIdx.x != 0) return;
    cr_direction_tests(st, true, recompute, iter, tpd, trs);
}

// s = inv(P) r and r's' at the start of a solve (reference :124-125) for the same case.
__global__ void __launch_bounds__(kBlock)
cr_init_smw_kernel(CrVectors v, SmwDev S, Reduce red, CrState* st) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    double rs = 0.0;
    smw_rows(S, v.diag, v.r, nullptr, false, [&](int i, double si, double) {
        v.s[i] = si;
        rs += __dmul_rn(si, v.r[i]);
    });
    const double bs = block_sum(rs, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, bs, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0) {
        st->rsdot_prev = ts;
        stamp(st, kSlotPre);
    }
}

// After a sharded C.Apply the dot lives in Cs[m] only once the allreduce has
// completed; this single-thread kernel then performs after_apply.
__global__ void cr_after_apply_kernel(const double* dot, int mode, int slot, CrState* st) {
    if (st->done) return;
    after_apply(st, mode, *dot, slot);
}

// ---- small elementwise helpers ----

// lhs = rhs ./ diag with the fused dot (reference src/diagonal_precond.cc:150-157).
__global__ void __launch_bounds__(kBlock)
diag_apply_kernel(int m, const double* __restrict__ diag, const double* __restrict__ rhs,
                  double* __restrict__ lhs, Reduce red, double* dot_out) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    double acc = 0.0;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < m; i += gridDim.x * kBlock) {
        const double l = rhs[i] / diag[i];
        lhs[i] = l;
        acc += __dmul_rn(l, rhs[i]);
    }
    const double b = block_sum(acc, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, b, 0.0, 0.0, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0)
        *dot_out = ts;
}

// KKTSolverDiag::_Factorize pass 1 (reference src/kkt_solver_diag.cc:34-41):
// W = 1/(zl/xl + zu/xu), and max of the finite W = 1/(smallest nonzero g)
// (rounding 1/g is monotone, so the maximum of the rounded reciprocals is the
// rounded reciprocal of the minimum).
__global__ void __launch_bounds__(kBlock)
kkt_weights_kernel(long long nm, const double* __restrict__ xl, const double* __restrict__ xu,
                   const double* __restrict__ zl, const double* __restrict__ zu,
                   double* __restrict__ W, Reduce red, double* maxw_out) {
    __shared__ double s_red[kWarps];
    __shared__ int s_flag;
    double best = 0.0;
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < nm;
         j += (long long)gridDim.x * kBlock) {
        const double g = zl[j] / xl[j] + zu[j] / xu[j];
        const double w = 1.0 / g;
        W[j] = w;
        if (g != 0.0 && isfinite(w)) best = fmax(best, w);
    }
    const double bm = block_max(best, s_red);
    double ts, ts2, tm;
    if (grid_reduce(red, 0.0, 0.0, bm, s_red, &s_flag, &ts, &ts2, &tm) && threadIdx.x == 0)
        *maxw_out = tm;
}

// Pass 2 (reference src/kkt_solver_diag.cc:42-56): infinite weights become
// 1/regval with regval = min(mu, smallest nonzero g), i.e.
// 1/regval = max(1/mu, *maxw); resscale = 1/sqrt(W[n+i]).
__global__ void __launch_bounds__(kBlock)
kkt_weights_fix_kernel(long long nm, long long n, double* __restrict__ W,
                       double* __restrict__ resscale, double mu, const double* maxw) {
    const double inv_regval = fmax(1.0 / mu, *maxw);
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < nm;
         j += (long long)gridDim.x * kBlock) {
        double w = W[j];
        if (isinf(w)) {
            w = inv_regval;
            W[j] = w;
        }
        if (j >= n) resscale[j - n] = 1.0 / sqrt(w);
    }
}

__global__ void __launch_bounds__(kBlock)
fill_kernel(long long n, double* x, double v) {
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < n;
         j += (long long)gridDim.x * kBlock)
        x[j] = v;
}

// out = a .* b
__global__ void __launch_bounds__(kBlock)
mul_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b,
           double* __restrict__ out) {
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < n;
         j += (long long)gridDim.x * kBlock)
        out[j] = __dmul_rn(a[j], b[j]);
}

// out = a .* b - c
__global__ void __launch_bounds__(kBlock)
mul_sub_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b,
               const double* __restrict__ c, double* __restrict__ out) {
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < n;
         j += (long long)gridDim.x * kBlock)
        out[j] = __dmul_rn(a[j], b[j]) - c[j];
}

// The normal matrix of a model without structural columns is diag(W_slack) (reference
// src/normal_matrix.cc:65-66 with an empty column loop): y = Ws .* x, y[m] = x'y, then the scalar
// step of the CR loop. One CTA.
__global__ void __launch_bounds__(kBlock)
slack_apply_kernel(int m, const double* __restrict__ Ws, const double* __restrict__ x,
                   double* __restrict__ y, int mode, int slot, CrState* st) {
    __shared__ double s_red[kWarps];
    if (st != nullptr && st->done) return;
    double dot = 0.0;
    for (int i = threadIdx.x; i < m; i += kBlock) {
        const double yi = Ws ? __dmul_rn(x[i], Ws[i]) : 0.0;
        y[i] = yi;
        dot += __dmul_rn(x[i], yi);
    }
    const double tot = block_sum(dot, s_red);
    if (threadIdx.x == 0) {
        y[m] = tot;
        if (st && !(mode == kApplyPlain && slot == kSlotNone)) after_apply(st, mode, tot, slot);
    }
}

// x = a - x
__global__ void __launch_bounds__(kBlock)
sub_from_kernel(long long n, const double* __restrict__ a, double* x) {
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < n;
         j += (long long)gridDim.x * kBlock)
        x[j] = a[j] - x[j];
}

}  // namespace ipxgpu
