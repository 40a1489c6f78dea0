// fix_device.cuh -- per-atom bodies of the device-resident channel fixes, shared by fix.cu
// (stand-alone hooks of the phase API) and integrate.cu (bounce fused into the step-boundary pass).
//
// Reference path:
//   gpu_fix_wall_force / gpu_fix_wall_bounce                 UM/fix_wall_meso.cu:74-117,147-193
//   gpu_fix_solid_wall_force<Rho5rc1s1> / ..._bounce         UM/fix_solid_bound_meso.cu:72-113,146-192, UM/fix_solid_bound_meso.h:42-63
//   gpu_fix_add_force                                        UM/fix_addforce_meso.cu:72-90
//   gpu_fix_pois_post_force                                  UM/fix_poiseuille_meso.cu:71-92
#pragma once
#include "internal.h"

namespace meso {

// polynomial wall force for rho = 5, rc = 1, s = 1 (UM/fix_solid_bound_meso.h:42-56), Horner form as written there
__device__ __forceinline__ double rho5rc1s1(double h)
{
    double s = +0.282625;
    s = s * h + -1.39021;
    s = s * h + +2.70259;
    s = s * h + -2.47678;
    s = s * h + +0.863184;
    s = s * h + +0.0664266;
    s = s * h + +0.0247250;
    s = s * h + +0.00856667;
    s = s * h + -0.116714;
    s = s * h * h + 0.0355959;
    return 75.0 * 6.2831853071796 * s;
}

// post_force of fix k on one atom: ff += contribution, one fp64 addition per reference statement (same order)
__device__ __forceinline__ void fix_force_one(const FixOp &o, const Box &box, int mk, const double (&xx)[3], double (&ff)[3])
{
    if (!(mk & o.groupbit)) return;
    if (o.kind == FIX_WALL) {
        const double d = o.p[0], dinv = o.p[1], f = o.p[2];
        if (f == 0.) return;                                          // MesoFixWall::boundary_force: `if( !f ) return`
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!((o.dims >> a) & 1)) continue;
            double h = xx[a] - box.boxlo[a];
            if (h <= d) ff[a] += f * erfcf((float)((h - 0.5 * d) * dinv * 1.732050808));
            h = box.boxhi[a] - xx[a];
            if (h <= d) ff[a] -= f * erfcf((float)((h - 0.5 * d) * dinv * 1.732050808));
        }
    } else if (o.kind == FIX_SOLID_BOUND) {
        const double d = 1.0;                                         // Rho5rc1s1::d
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!((o.dims >> a) & 1)) continue;
            double h = xx[a] - box.boxlo[a];
            if (h <= d) ff[a] += rho5rc1s1(h);
            h = box.boxhi[a] - xx[a];
            if (h <= d) ff[a] -= rho5rc1s1(h);
        }
    } else if (o.kind == FIX_ADDFORCE) {
        ff[0] += o.p[0]; ff[1] += o.p[1]; ff[2] += o.p[2];
    } else if (o.kind == FIX_POIS) {
        const int dim_ortho = o.dims & 3, dim_force = (o.dims >> 2) & 3;
        const double lower = box.boxlo[dim_ortho], upper = box.boxhi[dim_ortho];
        const double bisect = o.p[1] * upper + (1.0 - o.p[1]) * lower;
        const double r = dim_ortho == 0 ? xx[0] : (dim_ortho == 1 ? xx[1] : xx[2]);
        const double s = ((r < bisect && r >= lower) || r >= upper) ? o.p[0] : -o.p[0];
        if (dim_force == 0) ff[0] += s; else if (dim_force == 1) ff[1] += s; else ff[2] += s;
    }
}

// bounce-forward of fix k on one atom (pre_exchange and end_of_step hooks of wall/meso and solid_bound/meso)
__device__ __forceinline__ void fix_bounce_one(const FixOp &o, const Box &box, int mk, double (&xx)[3], double (&vv)[3])
{
    if ((o.kind != FIX_WALL && o.kind != FIX_SOLID_BOUND) || !(mk & o.groupbit)) return;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (!((o.dims >> a) & 1)) continue;
        if (xx[a] <= box.boxlo[a]) {
            vv[a] = fabs(vv[a]);
            xx[a] = 2. * box.boxlo[a] - xx[a];
        } else if (xx[a] >= box.boxhi[a]) {
            vv[a] = -fabs(vv[a]);
            xx[a] = 2. * box.boxhi[a] - xx[a];
        }
    }
}

__device__ __forceinline__ void fix_bounce_all(const FixList &fl, const Box &box, int mk, double (&xx)[3], double (&vv)[3])
{
    for (int k = 0; k < fl.n; k++) fix_bounce_one(fl.op[k], box, mk, xx, vv);
}

}  // namespace meso
