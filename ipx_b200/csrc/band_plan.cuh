// Planning, upload and launch of the banded sweeps (band_sweep.cuh) for a
// context.
#pragma once

#include <cstdlib>

#include "band_sweep.cuh"
#include "context.cuh"

namespace ipxgpu {

constexpr int kBandWarps = 31;   // consumer warps per CTA (+1 producer warp = 1024 threads)
constexpr int kBandDepth = 4;    // rows per register batch
constexpr int kBandVB = 8192;    // doubles per staged band (64 KB)
constexpr size_t kBandSmemBudget = 193 * 1024;  // 196 KB carve-out (incl. static + reserved part): a larger one leaves too little L1 for the loads in flight

// Chooses the tiling for S segments gathering from a vector of length V.
// Cost model, in bytes moved per item (one CTA): its share of the matrix
// stream, half weight for the bands it stages (they come from L2), and the
// partial output it writes and the combine pass reads back when the bands are
// split over several items. The plan is refused when the extra traffic
// exceeds `max_ratio` times the stream (the generic sweep is used instead).
static bool plan_band(BandPlan* out, int V, int S, long long nnz, int num_sms, double max_ratio) {
    if (V <= 0 || S <= 0 || nnz <= 0) return false;
    BandPlan best;
    double best_cost = -1.0;
    const int VB = std::min(kBandVB, (V + 1) & ~1);
    const int NVB = (V + VB - 1) / VB;
    const int NBUF = NVB > 1 ? 2 : 1;
    const double stream = 12.0 * (double)nnz;
    for (int pass = 0; pass < 2; pass++) {
        // pass 0: one item per segment block; pass 1: bands split over parts
        for (int NSB = 1; NSB <= std::max(4 * num_sms, 1); NSB++) {
            BandPlan P;
            P.V = V;
            P.S = S;
            P.VB = VB;
            P.NBUF = NBUF;
            P.NW = kBandWarps;
            P.nnz = nnz;
            P.SB = (S + NSB - 1) / NSB;
            if (P.SB > 65535) continue;
            if ((S + P.SB - 1) / P.SB != NSB) continue;  // same blocks as a smaller NSB
            int nparts = 1;
            if (pass == 1) {
                nparts = std::min(NVB, num_sms / NSB);
                if (nparts <= 1) continue;
            }
            P.K = (NVB + nparts - 1) / nparts;
            if (!band_finish_plan(&P) || P.smem > kBandSmemBudget) continue;
            const int waves = (P.nitems + num_sms - 1) / num_sms;
            const double staged = 8.0 * (double)P.K * VB;
            const double partial = P.nparts > 1 ? 24.0 * P.SB : 0.0;
            // the largest item: K bands of SB segments (the last part / block are smaller)
            const double share = std::min(1.0, (double)P.K * VB / V) * ((double)P.SB / S);
            const double per_item = stream * share + 0.5 * staged + partial;
            const double cost = waves * per_item;
            if (best_cost < 0.0 || cost < best_cost) {
                best_cost = cost;
                best = P;
            }
        }
    }
    if (best_cost < 0.0) return false;
    const double ideal = stream / num_sms;
    if (best_cost > (1.0 + max_ratio) * ideal) return false;
    *out = best;
    return true;
}

static void free_band(BandDev* T) {
    dev_free(T->row_ptr);
    dev_free(T->stream);
    dev_free(T->partials);
    *T = BandDev();
}

// Uploads host-built row streams. IPXGPU_ERR_UNSUPPORTED: more than max_pad padding (the
// structure does not suit the banded sweep); T is left empty.
static int upload_band(ipxgpu_ctx* c, BandDev* T, const BandHost& H, double max_pad) {
    if (H.rows > 0 && (double)H.pad_entries > max_pad * 32.0 * (double)H.rows)
        return IPXGPU_ERR_UNSUPPORTED;
    cudaStream_t s = c->stream;
    T->rows = H.rows;
    IPXGPU_TRY(upload(&T->row_ptr, H.row_ptr, s));
    {
        uint32_t* dev = nullptr;
        IPXGPU_TRY(upload(&dev, H.stream, s));
        T->stream = reinterpret_cast<unsigned char*>(dev);
    }
    if (T->plan.nparts > 1)
        IPXGPU_TRY(dev_alloc(&T->partials, (size_t)T->plan.nparts * T->plan.S));
    IPXGPU_CUDA(cudaStreamSynchronize(s));
    IPXGPU_CUDA(cudaFuncSetAttribute(band_sweep_kernel<kBandWarps, kBandDepth>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kBandSmemBudget));
    return IPXGPU_OK;
}

static bool band_usable(const BandDev* T, const double* v) {
    return T != nullptr && (reinterpret_cast<uintptr_t>(v) & 15u) == 0;
}

static int launch_band(ipxgpu_ctx* c, const BandDev& T, const BandArgs& A, int mode, CrState* st) {
    const BandPlan& P = T.plan;
    const int threads = (kBandWarps + 1) * 32;
    if (P.nparts == 1) {
        band_sweep_kernel<kBandWarps, kBandDepth>
            <<<P.nitems, threads, P.smem, c->stream>>>(T, A, mode, c->red, st);
        c->launches++;
    } else {
        band_sweep_kernel<kBandWarps, kBandDepth>
            <<<P.nitems, threads, P.smem, c->stream>>>(T, A, kBandPartial, c->red, st);
        if (P.nparts >= kBandWideParts) {
            const int grid = std::max(1, std::min((P.S + 31) / 32, c->num_sms * 8));
            band_combine_wide_kernel<<<grid, kBlock, 0, c->stream>>>(T, A, mode, c->red, st);
        } else {
            band_combine_kernel<<<grid_for(c, P.S), kBlock, 0, c->stream>>>(T, A, mode, c->red, st);
        }
        c->launches += 2;
    }
    IPXGPU_CUDA(cudaGetLastError());
    return IPXGPU_OK;
}

}  // namespace ipxgpu
