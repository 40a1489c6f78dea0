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

int launch_fix_post_force(meso_ctx *ctx, int handle, bool into_facc)
{
    if (ctx->fixes.nforce == 0) return MESO_OK;
    SoA3c x; SoA3 f;
    for (int d = 0; d < 3; d++) { x.c[d] = ctx->x[d].p; f.c[d] = ctx->f[d].p; }
    if (into_facc)
        k_fix_post_force<1><<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(x, f, ctx->facc.p, ctx->mask.p, ctx->d_counts, ctx->box, ctx->fixes, handle);
    else
        k_fix_post_force<0><<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(x, f, ctx->facc.p, ctx->mask.p, ctx->d_counts, ctx->box, ctx->fixes, handle);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_fix_bounce(meso_ctx *ctx, int handle)
{
    if (ctx->fixes.nbounce == 0) return MESO_OK;
    SoA3 x, v;
    for (int d = 0; d < 3; d++) { x.c[d] = ctx->x[d].p; v.c[d] = ctx->v[d].p; }
    k_fix_bounce<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(x, v, ctx->mask.p, ctx->d_counts, ctx->box, ctx->fixes, handle);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
