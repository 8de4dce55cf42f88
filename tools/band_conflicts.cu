// Host-only: shared-memory wavefront statistics of the banded row streams
// (ipx_b200/csrc/band_sweep.cuh) for a random sparse matrix of the benchmark shape.
// Per stream row (32 lanes) the kernel issues one 8-byte gather (LDS.64) and, for the
// lanes whose entry ends a run, a read-modify-write of the segment accumulator
// (LDS.64 + STS.64). A 64-bit access of a warp is served as two half-warps; a half-warp
// needs as many wavefronts as its most loaded bank pair (distinct addresses only).
//
//   nvcc -O2 -std=c++17 tools/band_conflicts.cu -o ipx_b200/_build/band_conflicts
//   band_conflicts [m n nnz_per_col]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>

#include "../ipx_b200/csrc/band_plan.cuh"

namespace ipxgpu { thread_local std::string g_last_error; }
using namespace ipxgpu;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rng() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Stats {
    long long rows = 0, entries = 0, wf_gather = 0, wf_acc = 0, acc_rows = 0;
};

static Stats conflicts(const BandPlan& P, const BandHost& H) {
    Stats S;
    S.rows = H.rows;
    for (long long r = 0; r < H.rows; r++) {
        const uint32_t* R = H.stream.data() + (size_t)r * 96;
        for (int half = 0; half < 2; half++) {
            int gcnt[16] = {0}, acnt[16] = {0};
            int gaddr[16][16], aaddr[16][16];
            bool any_acc = false;
            for (int l = 16 * half; l < 16 * half + 16; l++) {
                const uint32_t key = R[l];
                const int seg = key >> 16, idx = key & 0x7fff;
                if (seg != P.SB) S.entries++;
                {   // gather v_s[idx]
                    const int b = idx & 15;
                    bool dup = false;
                    for (int q = 0; q < gcnt[b]; q++) dup |= gaddr[b][q] == idx;
                    if (!dup) gaddr[b][gcnt[b]++] = idx;
                }
                if (key & kBandLast) {
                    any_acc = true;
                    const int b = seg & 15;
                    bool dup = false;
                    for (int q = 0; q < acnt[b]; q++) dup |= aaddr[b][q] == seg;
                    if (!dup) aaddr[b][acnt[b]++] = seg;
                }
            }
            int gmax = 0, amax = 0;
            for (int b = 0; b < 16; b++) {
                gmax = std::max(gmax, gcnt[b]);
                amax = std::max(amax, acnt[b]);
            }
            S.wf_gather += gmax;
            if (any_acc) S.wf_acc += 2 * amax;  // load + store
        }
    }
    return S;
}

int main(int argc, char** argv) {
    const int m = argc > 1 ? atoi(argv[1]) : 100000;
    const int n = argc > 2 ? atoi(argv[2]) : 1000000;
    const int k = argc > 3 ? atoi(argv[3]) : 10;
    std::vector<int> cp(n + 1), ci((size_t)n * k);
    std::vector<double> cx((size_t)n * k);
    for (int j = 0; j < n; j++) {
        cp[j] = j * k;
        int* rows = ci.data() + (size_t)j * k;
        for (int q = 0; q < k;) {
            const int r = (int)(rng() % (uint64_t)m);
            bool dup = false;
            for (int t = 0; t < q; t++) dup |= rows[t] == r;
            if (!dup) rows[q++] = r;
        }
        std::sort(rows, rows + k);
        for (int q = 0; q < k; q++) cx[(size_t)j * k + q] = 0.5 + (rng() % 1000) / 300.0;
    }
    cp[n] = n * k;
    const long long nnz = (long long)n * k;
    std::vector<int> rp(m + 1, 0), rj((size_t)nnz);
    std::vector<double> rx((size_t)nnz);
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
    for (int sweep = 0; sweep < 2; sweep++) {
        BandPlan P;
        const bool ok = sweep == 0 ? plan_band(&P, m, n, nnz, 148, 1.5) : plan_band(&P, n, m, nnz, 148, 1.5);
        if (!ok) { printf("sweep %d: no plan\n", sweep + 1); continue; }
        BandHost H;
        const auto t0 = std::chrono::steady_clock::now();
        const bool built = sweep == 0 ? band_build(P, cp.data(), ci.data(), cx.data(), &H)
                                      : band_build(P, rp.data(), rj.data(), rx.data(), &H);
        const double tb = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!built) { printf("sweep %d: not built\n", sweep + 1); continue; }
        const Stats S = conflicts(P, H);
        printf("sweep %d: SB %d K %d items %d | rows %lld entries %lld pad %.1f%% | build %.2f s | "
               "wavefronts per row: gather %.2f acc %.2f total %.2f (+3 stream) | per 32 entries %.2f\n",
               sweep + 1, P.SB, P.K, P.nitems, S.rows, S.entries,
               100.0 * (32.0 * S.rows - S.entries) / (32.0 * S.rows), tb,
               (double)S.wf_gather / S.rows, (double)S.wf_acc / S.rows,
               (double)(S.wf_gather + S.wf_acc) / S.rows,
               32.0 * (double)(S.wf_gather + S.wf_acc + 3 * S.rows) / S.entries);
    }
    return 0;
}
