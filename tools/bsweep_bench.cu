// Standalone tuning harness for the banded sweeps (ipx_b200/csrc/band_sweep.cuh).
// Generates a random sparse matrix of the benchmark shape, runs sweep 1
// (t = W .* A'x) and sweep 2 (y = A t) in several configurations, checks them
// against a host computation and prints device times.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -lineinfo \
//        tools/bsweep_bench.cu -o ipx_b200/_build/bsweep_bench
//   bsweep_bench [m n nnz_per_col] [config ...]
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../ipx_b200/csrc/band_sweep.cuh"

using namespace ipxgpu;

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call,              \
                    cudaGetErrorString(e_));                                          \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rng() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double urand() { return (rng() >> 11) * (1.0 / 9007199254740992.0); }

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
}


// Host emulation of band_sweep_kernel's walk over the row streams (format check).
static void emulate(const BandPlan& P, const BandHost& H, const double* v, std::vector<double>* out) {
    out->assign(P.S, 0.0);
    std::vector<double> acc(P.SB + 1);
    for (int sb = 0; sb < P.NSB; sb++)
        for (int part = 0; part < P.nparts; part++) {
            const int item = sb * P.nparts + part;
            const int vb0 = part * P.K;
            const int nk = std::min(P.NVB, vb0 + P.K) - vb0;
            const int rot = sb % nk;
            std::fill(acc.begin(), acc.end(), 0.0);
            for (int w = 0; w < P.NW; w++) {
                const int* rp = H.row_ptr.data() + ((size_t)item * P.NW + w) * (P.K + 1);
                for (int l = 0; l < 32; l++) {
                    double sum = 0;
                    for (int k = 0; k < nk; k++) {
                        int r = k + rot;
                        if (r >= nk) r -= nk;
                        const int vbase = (vb0 + r) * P.VB;
                        for (int row = rp[k]; row < rp[k + 1]; row++) {
                            const uint32_t* R = H.stream.data() + (size_t)row * 96;
                            const uint32_t key = R[l];
                            const double a = reinterpret_cast<const double*>(R + 32)[l];
                            if ((key >> 16) == (uint32_t)P.SB) continue;
                            sum += v[vbase + (key & 0x7fff)] * a;
                            if (key & kBandLast) {
                                acc[key >> 16] += sum;
                                sum = 0;
                            }
                        }
                    }
                }
            }
            const int nseg = std::min(P.SB, P.S - sb * P.SB);
            for (int q = 0; q < nseg; q++) (*out)[sb * P.SB + q] += acc[q];
        }
}

typedef void (*KernelFn)(BandDev, BandArgs, int, Reduce, CrState*);

struct Cfg {
    int tma;  // 0: register batches (NW, D, LD); 1: bulk-copy rings (NW, NS=D, C=LD)
    int NW, D, NBUF, VB1, VB2, SB2, K2, LD;
};

template <int NW, int D, int LD>
static KernelFn pick_reg(int dbg) {
    switch (dbg) {
        case 0: return band_sweep_kernel<NW, D, 0, LD>;
        case 1: return band_sweep_kernel<NW, D, 1, LD>;
        case 2: return band_sweep_kernel<NW, D, 2, LD>;
        case 4: return band_sweep_kernel<NW, D, 4, LD>;
        default: return band_sweep_kernel<NW, D, 3, LD>;
    }
}
static KernelFn pick(const Cfg& c, int dbg) {
#define REG(NW_, D_, LD_) \
    if (!c.tma && c.NW == NW_ && c.D == D_ && c.LD == LD_) return pick_reg<NW_, D_, LD_>(dbg);
    REG(31, 4, 0) REG(31, 3, 0) REG(31, 2, 0) REG(31, 4, 1) REG(31, 4, 2) REG(31, 4, 3)
    REG(16, 8, 0) REG(16, 8, 1) REG(24, 4, 0) REG(24, 4, 1)
    fprintf(stderr, "no instantiation tma=%d NW=%d D/NS=%d LD/C=%d\n", c.tma, c.NW, c.D, c.LD);
    exit(2);
}

static size_t cfg_smem(const Cfg& c, const BandPlan& P) {
    static const size_t extra = getenv("BSWEEP_EXTRA_SMEM") ? atoi(getenv("BSWEEP_EXTRA_SMEM")) : 0;
    return P.smem + extra;
}

static void launch_any(const Cfg& c, const BandDev& T, const BandArgs& A, int mode, Reduce red,
                       cudaStream_t s) {
    KernelFn fn = pick(c, T.debug);
    const size_t smem = cfg_smem(c, T.plan);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fn<<<T.plan.nitems, (c.NW + 1) * 32, smem, s>>>(T, A, mode, red, nullptr);
}

int main(int argc, char** argv) {
    int m = 100000, n = 1000000, k = 10;
    int argi = 1;
    if (argc >= 4 && atoi(argv[1]) > 0 && !strchr(argv[1], ',')) {
        m = atoi(argv[1]);
        n = atoi(argv[2]);
        k = atoi(argv[3]);
        argi = 4;
    }
    int iters = 15;
    printf("matrix %d x %d, %d per column\n", m, n, k);
    // ---- matrix ----
    const long long nnz = (long long)n * k;
    std::vector<int> cp(n + 1), ci(nnz);
    std::vector<double> cx(nnz);
    {
        std::vector<int> rows(k);
        for (int j = 0; j < n; j++) {
            cp[j] = j * k;
            int got = 0;
            while (got < k) {
                const int r = (int)(rng() % (uint64_t)m);
                bool dup = false;
                for (int q = 0; q < got; q++) dup |= rows[q] == r;
                if (!dup) rows[got++] = r;
            }
            std::sort(rows.begin(), rows.end());
            for (int q = 0; q < k; q++) {
                ci[(size_t)j * k + q] = rows[q];
                const double a = 0.5 + 3.5 * urand();
                cx[(size_t)j * k + q] = (rng() & 1) ? a : -a;
            }
        }
        cp[n] = (int)nnz;
    }
    std::vector<int> rp(m + 1, 0), rj(nnz);
    std::vector<double> rx(nnz);
    for (long long p = 0; p < nnz; p++) rp[ci[p] + 1]++;
    for (int i = 0; i < m; i++) rp[i + 1] += rp[i];
    {
        std::vector<int> next(rp.begin(), rp.end() - 1);
        for (int j = 0; j < n; j++)
            for (int p = cp[j]; p < cp[j + 1]; p++) {
                const int put = next[ci[p]]++;
                rj[put] = j;
                rx[put] = cx[p];
            }
    }
    std::vector<double> x(m), W(n), t_ref(n), y_ref(m);
    for (auto& v : x) v = urand() * 2 - 1;
    for (auto& v : W) v = std::exp(4 * urand() - 2);
    for (int j = 0; j < n; j++) {
        double d = 0;
        for (int p = cp[j]; p < cp[j + 1]; p++) d += x[ci[p]] * cx[p];
        t_ref[j] = d * W[j];
    }
    for (int i = 0; i < m; i++) {
        double d = 0;
        for (int p = rp[i]; p < rp[i + 1]; p++) d += t_ref[rj[p]] * rx[p];
        y_ref[i] = d;
    }
    double tmax = 0, ymax = 0;
    for (double v : t_ref) tmax = std::max(tmax, std::fabs(v));
    for (double v : y_ref) ymax = std::max(ymax, std::fabs(v));

    const bool emu = getenv("BSWEEP_EMULATE") != nullptr;
    cudaStream_t s = nullptr;
    double *d_x = nullptr, *d_W = nullptr, *d_t = nullptr, *d_y = nullptr;
    Reduce red{nullptr, nullptr};
    const size_t flush_bytes = 512ull << 20;
    char* d_flush = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (!emu) {
        CK(cudaStreamCreate(&s));
        CK(cudaMalloc(&d_x, (m + 2) * 8));
        CK(cudaMalloc(&d_W, (size_t)n * 8));
        CK(cudaMalloc(&d_t, ((size_t)n + 2) * 8));
        CK(cudaMalloc(&d_y, (m + 2) * 8));
        CK(cudaMemcpy(d_x, x.data(), m * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_W, W.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&red.partials, 3 * 4096 * 8));
        CK(cudaMalloc(&red.ticket, 4));
        CK(cudaMemset(red.ticket, 0, 4));
        CK(cudaMalloc(&d_flush, flush_bytes));
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
    }
    std::vector<Cfg> cfgs;
    for (; argi < argc; argi++) {
        Cfg c;
        c.LD = 0;
        c.tma = 0;
        const char* str = argv[argi];
        if (str[0] == 't') {
            c.tma = 1;
            str++;
        }
        const int got = sscanf(str, "%d,%d,%d,%d,%d,%d,%d,%d", &c.NW, &c.D, &c.NBUF, &c.VB1, &c.VB2,
                               &c.SB2, &c.K2, &c.LD);
        if (got >= 7)
            cfgs.push_back(c);
        else if (sscanf(argv[argi], "iters=%d", &iters) == 1) {
        } else {
            fprintf(stderr, "bad config %s (NW,D,NBUF,VB1,VB2,SB2,K2)\n", argv[argi]);
            return 2;
        }
    }
    if (cfgs.empty()) cfgs.push_back(Cfg{0, 16, 8, 2, 8192, 8192, 6250, 14, 0});

    auto time_kernel = [&](auto&& fn, double* best, double* med) {
        std::vector<float> ts;
        for (int it = 0; it < iters + 3; it++) {
            CK(cudaMemsetAsync(d_flush, it, flush_bytes, s));
            CK(cudaEventRecord(e0, s));
            fn();
            CK(cudaEventRecord(e1, s));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 3) ts.push_back(ms);
        }
        std::sort(ts.begin(), ts.end());
        *best = ts.front() * 1e3;
        *med = ts[ts.size() / 2] * 1e3;
    };

    const bool do_dbg = getenv("BSWEEP_DBG") != nullptr;
    for (const Cfg& c : cfgs) {
        printf("== cfg %s NW=%d D/NS=%d LD/C=%d NBUF=%d VB1=%d VB2=%d SB2=%d K2=%d\n",
               c.tma ? "tma" : "reg", c.NW, c.D, c.LD, c.NBUF, c.VB1, c.VB2, c.SB2, c.K2);
        BandDev Ts[3];
        BandArgs As[3];
        long long rows_of[3] = {0, 0, 0};
        bool okall = true;
        for (int sweep = 1; sweep <= 2; sweep++) {
            BandPlan P;
            P.NW = c.NW;
            P.NBUF = c.NBUF;
            P.nnz = nnz;
            if (sweep == 1) {
                P.V = m;
                P.S = n;
                P.VB = std::min(c.VB1, (m + 1) & ~1);
                P.SB = (n + 147) / 148;
                P.K = 0;
            } else {
                P.V = n;
                P.S = m;
                P.VB = std::min(c.VB2, (n + 1) & ~1);
                P.SB = std::min(c.SB2, m);
                P.K = c.K2;
            }
            if (!band_finish_plan(&P) || cfg_smem(c, P) > 227 * 1024) {
                printf("  sweep %d: plan rejected (smem %zu)\n", sweep, cfg_smem(c, P));
                okall = false;
                break;
            }
            BandHost H;
            const double tb0 = now_s();
            const bool ok = sweep == 1 ? band_build(P, cp.data(), ci.data(), cx.data(), &H)
                                       : band_build(P, rp.data(), rj.data(), rx.data(), &H);
            const double tb1 = now_s();
            if (!ok) {
                printf("  sweep %d: build rejected\n", sweep);
                okall = false;
                break;
            }
            if (emu) {
                std::vector<double> out;
                emulate(P, H, sweep == 1 ? x.data() : t_ref.data(), &out);
                double e = 0;
                for (int i = 0; i < P.S; i++) {
                    const double ref = sweep == 1 ? t_ref[i] / W[i] : y_ref[i];
                    e = std::max(e, std::fabs(out[i] - ref));
                }
                printf("  sweep %d: emulation abs err %.3e rows=%lld pad=%.2f%% build=%.2fs\n", sweep, e,
                       H.rows, 100.0 * H.pad_entries / (double)(H.rows * 32), tb1 - tb0);
                continue;
            }
            BandDev& T = Ts[sweep];
            T.plan = P;
            T.rows = H.rows;
            rows_of[sweep] = H.rows;
            CK(cudaMalloc(&T.row_ptr, H.row_ptr.size() * 4));
            CK(cudaMemcpy(T.row_ptr, H.row_ptr.data(), H.row_ptr.size() * 4,
                          cudaMemcpyHostToDevice));
            CK(cudaMalloc(&T.stream, H.stream.size() * 4));
            CK(cudaMemcpy(T.stream, H.stream.data(), H.stream.size() * 4, cudaMemcpyHostToDevice));
            if (P.nparts > 1) CK(cudaMalloc(&T.partials, (size_t)P.nparts * P.S * 8));
            printf("  sweep %d: VB=%d SB=%d NVB=%d NSB=%d K=%d nparts=%d nitems=%d smem=%zu rows=%lld "
                   "pad=%.2f%% stream=%.1f MB build=%.2fs\n",
                   sweep, P.VB, P.SB, P.NVB, P.NSB, P.K, P.nparts, P.nitems, P.smem, H.rows,
                   100.0 * H.pad_entries / (double)(H.rows * 32), H.rows * 384 / 1e6, tb1 - tb0);
            if (sweep == 1) As[1] = BandArgs{d_x, d_W, nullptr, nullptr, d_t, 0, 0};
            else As[2] = BandArgs{d_t, nullptr, nullptr, d_x, d_y, 0, 0};
        }
        if (emu || !okall) continue;
        const int grid_comb = 148 * 4;
        auto run_sweep = [&](int sweep) {
            const BandDev& T = Ts[sweep];
            if (sweep == 1) {
                launch_any(c, T, As[1], kBandColScale, red, s);
            } else if (T.plan.nparts == 1) {
                launch_any(c, T, As[2], kBandRowFinal, red, s);
            } else {
                launch_any(c, T, As[2], kBandPartial, red, s);
                band_combine_kernel<<<grid_comb, kBlock, 0, s>>>(T, As[2], kBandRowFinal, red,
                                                                 nullptr);
            }
        };
        // ---- each sweep alone, L2 flushed by a 512 MB memset before every launch ----
        for (int sweep = 1; sweep <= 2; sweep++) {
            if (sweep == 2)
                CK(cudaMemcpy(d_t, t_ref.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
            for (int dbg : {0, 2, 3, 1}) {
                if (dbg != 0 && !do_dbg) continue;
                Ts[sweep].debug = dbg;
                double best, med;
                time_kernel([&]() { run_sweep(sweep); }, &best, &med);
                double err = -1;
                if (dbg == 0) {
                    const int len = sweep == 1 ? n : m;
                    std::vector<double> out(len);
                    CK(cudaMemcpy(out.data(), sweep == 1 ? d_t : d_y, (size_t)len * 8,
                                  cudaMemcpyDeviceToHost));
                    const std::vector<double>& ref = sweep == 1 ? t_ref : y_ref;
                    double e = 0;
                    for (int i = 0; i < len; i++) e = std::max(e, std::fabs(out[i] - ref[i]));
                    err = e / (sweep == 1 ? tmax : ymax);
                }
                printf("    sweep %d flushed debug=%d  best %.1f us  median %.1f us  stream %.0f GB/s  "
                       "relerr %.2e\n",
                       sweep, dbg, best, med, rows_of[sweep] * 384 / (best * 1e-6) / 1e9, err);
            }
            Ts[sweep].debug = 0;
        }
        // ---- the apply as CR runs it: sweep 1, sweep 2 back to back, no flush ----
        for (int adbg : {0, 1, 2, 3}) {
            if (adbg != 0 && !do_dbg) continue;
            Ts[1].debug = Ts[2].debug = adbg;
            if (adbg) printf("    [debug=%d: %s%s] ", adbg, (adbg & 1) ? "no staging " : "",
                             (adbg & 2) ? "no shared-memory work" : "");
            const int reps = 30;
            std::vector<cudaEvent_t> ev(2 * reps + 1);
            for (auto& e : ev) CK(cudaEventCreate(&e));
            for (int w = 0; w < 3; w++) {
                run_sweep(1);
                run_sweep(2);
            }
            CK(cudaEventRecord(ev[0], s));
            for (int r = 0; r < reps; r++) {
                run_sweep(1);
                CK(cudaEventRecord(ev[2 * r + 1], s));
                run_sweep(2);
                CK(cudaEventRecord(ev[2 * r + 2], s));
            }
            CK(cudaStreamSynchronize(s));
            CK(cudaGetLastError());
            double t1 = 0, t2 = 0;
            float ms;
            for (int r = 0; r < reps; r++) {
                CK(cudaEventElapsedTime(&ms, ev[2 * r], ev[2 * r + 1]));
                t1 += ms;
                CK(cudaEventElapsedTime(&ms, ev[2 * r + 1], ev[2 * r + 2]));
                t2 += ms;
            }
            CK(cudaEventElapsedTime(&ms, ev[0], ev[2 * reps]));
            // same loop without the inner events
            CK(cudaEventRecord(ev[0], s));
            for (int r = 0; r < reps; r++) {
                run_sweep(1);
                run_sweep(2);
            }
            CK(cudaEventRecord(ev[1], s));
            CK(cudaStreamSynchronize(s));
            float ms2;
            CK(cudaEventElapsedTime(&ms2, ev[0], ev[1]));
            const double alg = 2.0 * nnz * 12 + 4.0 * (n + 1) + 4.0 * (m + 1) + 8.0 * (n + m) + 16.0 * m;
            printf("    APPLY loop: sweep1 %.1f us  sweep2 %.1f us  apply %.1f us (no inner events %.1f us)"
                   "  algorithmic %.0f GB/s\n",
                   t1 / reps * 1e3, t2 / reps * 1e3, ms / reps * 1e3, ms2 / reps * 1e3,
                   alg / (ms2 / reps * 1e-3) / 1e9);
            if (adbg == 0) {
                std::vector<double> out(m);
                CK(cudaMemcpy(out.data(), d_y, (size_t)m * 8, cudaMemcpyDeviceToHost));
                double e = 0;
                for (int i = 0; i < m; i++) e = std::max(e, std::fabs(out[i] - y_ref[i]));
                printf("    apply relerr %.2e\n", e / ymax);
            }
            for (auto& e2 : ev) CK(cudaEventDestroy(e2));
        }
        Ts[1].debug = Ts[2].debug = 0;
        if (getenv("BSWEEP_TRACE")) {
            for (int sweep = 1; sweep <= 2; sweep++) {
                BandDev& T = Ts[sweep];
                const int NWp = 2 * c.NW + 4, ni = T.plan.nitems;
                CK(cudaMalloc(&T.trace, (size_t)ni * NWp * 8));
                CK(cudaMemset(T.trace, 0, (size_t)ni * NWp * 8));
                T.debug = 4;
                for (int w = 0; w < 3; w++) {
                    run_sweep(sweep == 1 ? 2 : 1);  // evict this sweep's stream from L2
                    CK(cudaStreamSynchronize(s));
                    Ts[3 - sweep].debug = 0;
                    run_sweep(sweep);
                }
                CK(cudaStreamSynchronize(s));
                std::vector<unsigned long long> tr((size_t)ni * NWp);
                CK(cudaMemcpy(tr.data(), T.trace, tr.size() * 8, cudaMemcpyDeviceToHost));
                unsigned long long t0 = ~0ull, t3 = 0;
                for (int i = 0; i < ni; i++) {
                    t0 = std::min(t0, tr[(size_t)i * NWp]);
                    t3 = std::max(t3, tr[(size_t)i * NWp + 3]);
                }
                auto stats = [&](auto&& get, const char* name) {
                    double mn = 1e30, mx = 0, sum = 0;
                    for (int i = 0; i < ni; i++) {
                        const double v = get(i);
                        mn = std::min(mn, v);
                        mx = std::max(mx, v);
                        sum += v;
                    }
                    printf("      %-28s min %7.2f  mean %7.2f  max %7.2f us\n", name, mn / 1e3,
                           sum / ni / 1e3, mx / 1e3);
                };
                printf("    TRACE sweep %d: first start -> last end %.2f us\n", sweep,
                       (t3 - t0) / 1e3);
                stats([&](int i) { return (double)(tr[(size_t)i * NWp] - t0); }, "item start (rel)");
                stats([&](int i) { return (double)(tr[(size_t)i * NWp + 1] - tr[(size_t)i * NWp]); },
                      "prologue");
                stats([&](int i) {
                    unsigned long long e = ~0ull;
                    for (int w = 0; w < c.NW; w++) e = std::min(e, tr[(size_t)i * NWp + 4 + w]);
                    return (double)(e - tr[(size_t)i * NWp + 1]);
                }, "first warp done");
                stats([&](int i) { return (double)(tr[(size_t)i * NWp + 2] - tr[(size_t)i * NWp + 1]); },
                      "all warps done");
                stats([&](int i) { return (double)(tr[(size_t)i * NWp + 3] - tr[(size_t)i * NWp + 2]); },
                      "epilogue");
                stats([&](int i) {
                    double sum = 0;
                    for (int w = 0; w < c.NW; w++) sum += (double)tr[(size_t)i * NWp + 4 + c.NW + w];
                    return sum / c.NW;
                }, "band wait per warp (mean)");
                stats([&](int i) {
                    double mx = 0;
                    for (int w = 0; w < c.NW; w++) mx = std::max(mx, (double)tr[(size_t)i * NWp + 4 + c.NW + w]);
                    return mx;
                }, "band wait per warp (max)");
                stats([&](int i) { return (double)(tr[(size_t)i * NWp + 3] - t0); }, "item end (rel)");
                T.debug = 0;
                CK(cudaFree(T.trace));
                T.trace = nullptr;
            }
        }
        for (int sweep = 1; sweep <= 2; sweep++) {
            CK(cudaFree(Ts[sweep].row_ptr));
            CK(cudaFree(Ts[sweep].stream));
            if (Ts[sweep].partials) CK(cudaFree(Ts[sweep].partials));
        }
        fflush(stdout);
    }
    return 0;
}
