// fix.cu -- stand-alone hooks of the device-resident channel fixes.
//
// Reference path: MesoFixWall::{post_force,pre_exchange,end_of_step} UM/fix_wall_meso.cu:119-238, MesoFixSolidBound
// UM/fix_solid_bound_meso.cu:115-234, MesoFixAddForce::post_force UM/fix_addforce_meso.cu:92-108, MesoFixPoiseuille::post_force
// UM/fix_poiseuille_meso.cu:94-113.  The reference launches one kernel per fix and hook; here one streaming pass applies the
// whole list (a handful of predicated fp64 operations per atom: HBM-bound, 28 B read + 24 B written per atom for the force
// hook), and inside meso_run the bounce rides the step-boundary pass of integrate.cu, so a channel deck keeps the
// kernel count of the plain fluid plus one.
#include "internal.h"
#include "fix_device.cuh"
#include <vector>

namespace meso {

struct SoA3 { double *c[3]; };
struct SoA3c { const double *c[3]; };

// ACC = 0: f += (fp64, the reference's statement order);  ACC = 1: facc += (fp32 accumulator of the pair-once loop)
template <int ACC>
__global__ void __launch_bounds__(256) k_fix_post_force(SoA3c x, SoA3 f, float4 *__restrict__ facc, const int *__restrict__ mask,
                                                        const Counts *__restrict__ cnt, Box box, FixList fl, int only)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int mk = mask[i];
        const double xx[3] = {x.c[0][i], x.c[1][i], x.c[2][i]};
        double ff[3];
        if (ACC) { ff[0] = ff[1] = ff[2] = 0.; }
        else { ff[0] = f.c[0][i]; ff[1] = f.c[1][i]; ff[2] = f.c[2][i]; }
        for (int k = 0; k < fl.n; k++)
            if (only < 0 || only == k) fix_force_one(fl.op[k], box, mk, xx, ff);
        if (ACC) {
            if (ff[0] != 0. || ff[1] != 0. || ff[2] != 0.) {
                float4 a = facc[i];
                a.x += (float)ff[0]; a.y += (float)ff[1]; a.z += (float)ff[2];
                facc[i] = a;
            }
        } else { f.c[0][i] = ff[0]; f.c[1][i] = ff[1]; f.c[2][i] = ff[2]; }
    }
}

__global__ void __launch_bounds__(256) k_fix_bounce(SoA3 x, SoA3 v, const int *__restrict__ mask, const Counts *__restrict__ cnt, Box box,
                                                    FixList fl, int only)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int mk = mask[i];
        double xx[3] = {x.c[0][i], x.c[1][i], x.c[2][i]};
        double vv[3] = {v.c[0][i], v.c[1][i], v.c[2][i]};
        const double x0 = xx[0], x1 = xx[1], x2 = xx[2];
        for (int k = 0; k < fl.n; k++)
            if (only < 0 || only == k) fix_bounce_one(fl.op[k], box, mk, xx, vv);
        if (xx[0] != x0 || xx[1] != x1 || xx[2] != x2) {               // a reflected atom: the only ones written back
#pragma unroll
            for (int d = 0; d < 3; d++) { x.c[d][i] = xx[d]; v.c[d][i] = vv[d]; }
        }
    }
}

// ------------------------------------------------------------------ rdf/fast/meso
// gpu_calc_rdf, UM/fix_rdf_fast_meso.cu:102-150: histogram of the pair distances r < rc over the stored neighbor table
// (full list: every i-j pair is counted from both sides), fp32 on the packed coordinates, bin = floor(r nbin / rc).
// One lane per row of the tile-transposed table (coalesced index loads), block-private histogram in shared memory,
// flushed once per CTA into 64-bit global counters (the reference's 32-bit counters wrap after ~250 samples of 1M atoms).
__global__ void __launch_bounds__(256) k_rdf(const float4 *__restrict__ coord4, const int *__restrict__ mask, const int *__restrict__ pair_count,
                                             const int *__restrict__ pair_table, unsigned long long *__restrict__ hist,
                                             const Counts *__restrict__ cnt, int n_col, float rc, float bin_sz_inv, int nbin, int groupi,
                                             int groupj)
{
    extern __shared__ unsigned hist_local[];
    for (int b = threadIdx.x; b < nbin; b += blockDim.x) hist_local[b] = 0;
    __syncthreads();
    const int n = cnt->nlocal;
    const float rcsq = rc * rc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!(mask[i] & groupi)) continue;
        const float4 c1 = coord4[i];
        const int n_pair = pair_count[i];
        const int *row = pair_table + (size_t)(i & ~31) * (size_t)n_col + (i & 31);
        for (int k = 0; k < n_pair; k++) {
            const int j = row[(size_t)(k & 31) * n_col + (k & ~31)];
            if (groupj != 1 && !(mask[j] & groupj)) continue;          // group bit 1 is LAMMPS' `all`
            const float4 c2 = coord4[j];
            const float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z;
            const float rsq = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            if (rsq < rcsq) {
                const float r = rsq * rsqrtf(rsq);
                const int bid = (int)floorf(r * bin_sz_inv);
                if (bid >= 0 && bid < nbin) atomicAdd(hist_local + bid, 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbin; b += blockDim.x)
        if (hist_local[b]) atomicAdd(hist + b, (unsigned long long)hist_local[b]);
}

__global__ void __launch_bounds__(256) k_group_count(const int *__restrict__ mask, const Counts *__restrict__ cnt, int gi, int gj,
                                                     unsigned long long *__restrict__ out2)
{
    const int n = cnt->nlocal;
    unsigned a = 0, b = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int m = mask[i];
        a += (m & gi) != 0; b += (m & gj) != 0;
    }
    a = __reduce_add_sync(0xffffffffu, a); b = __reduce_add_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0) { atomicAdd(out2, (unsigned long long)a); atomicAdd(out2 + 1, (unsigned long long)b); }
}

int launch_fix_rdf(meso_ctx *ctx, int handle)
{
    if (ctx->fixes.nrdf == 0) return MESO_OK;
    for (int k = 0; k < ctx->fixes.n; k++) {
        const FixOp &o = ctx->fixes.op[k];
        if (o.kind != FIX_RDF || (handle >= 0 && handle != k)) continue;
        const int every = (int)o.p[0], nbin = o.dims;
        if (every > 1 && ctx->ntimestep % every != 0) continue;       // UM/fix_rdf_fast_meso.cu:160
        const float rc = (float)o.p[1];
        k_rdf<<<grid_for(ctx, 4), 256, sizeof(unsigned) * nbin, LS(ctx->stream)>>>(ctx->coord4.p, ctx->mask.p, ctx->pair_count.p, ctx->pair_table.p,
                                                                            ctx->rdf_hist[k].p, ctx->d_counts, ctx->n_col, rc, (float)nbin / rc,
                                                                            nbin, o.groupbit, o.aux);
        ctx->rdf_samples[k]++;
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// histogram (as doubles) and the sizes of the two groups among this rank's local atoms
int fix_rdf_read(meso_ctx *ctx, int handle, int nbin, double *hist, double *ni, double *nj)
{
    const FixOp &o = ctx->fixes.op[handle];
    std::vector<unsigned long long> h((size_t)nbin + 2, 0ull);
    unsigned long long *cnt2 = ctx->rdf_hist[handle].p + nbin;       // two spare counters behind the bins
    MESO_CUDA(cudaMemsetAsync(cnt2, 0, 2 * sizeof(unsigned long long), ctx->stream));
    k_group_count<<<grid_for(ctx, 2), 256, 0, LS(ctx->stream)>>>(ctx->mask.p, ctx->d_counts, o.groupbit, o.aux, cnt2);
    MESO_CUDA(cudaMemcpyAsync(h.data(), ctx->rdf_hist[handle].p, sizeof(unsigned long long) * (nbin + 2), cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < nbin; b++) hist[b] = (double)h[b];
    *ni = (double)h[nbin]; *nj = (double)h[nbin + 1];
    return MESO_OK;
}

int launch_fix_post_force(meso_ctx *ctx, int handle, bool into_facc)
{
    if (ctx->fixes.nforce == 0) return MESO_OK;
    SoA3c x; SoA3 f;
    for (int d = 0; d < 3; d++) { x.c[d] = ctx->x[d].p; f.c[d] = ctx->f[d].p; }
    if (into_facc)
        k_fix_post_force<1><<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(x, f, ctx->facc.p, ctx->mask.p, ctx->d_counts, ctx->box, ctx->fixes, handle);
    else
        k_fix_post_force<0><<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(x, f, ctx->facc.p, ctx->mask.p, ctx->d_counts, ctx->box, ctx->fixes, handle);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_fix_bounce(meso_ctx *ctx, int handle)
{
    if (ctx->fixes.nbounce == 0) return MESO_OK;
    SoA3 x, v;
    for (int d = 0; d < 3; d++) { x.c[d] = ctx->x[d].p; v.c[d] = ctx->v[d].p; }
    k_fix_bounce<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(x, v, ctx->mask.p, ctx->d_counts, ctx->box, ctx->fixes, handle);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
