// pair.cu -- packing (+ per-step signatures) and the DPD pair-force kernels.
//
// Reference path:
//   gpu_merge_xvt / dp2sp_merged      UM/atom_vec_meso.cu:142-192       (A7)
//   gpu_dpd_fast<ev>                  UM/pair_dpd_fast_meso.cu:91-205   (A10, fp32)
//   gpu_dpd<ev>                       UM/pair_dpd_meso.cu:91-205        (A11, fp64 on the fp32-packed inputs)
//   compute / compute_bulk / compute_border  UM/pair_dpd_meso.cu:241-266
// One thread owns one local particle and walks its row of the tile-transposed table
// (coalesced 128-byte index loads across the warp); neighbor float4s are gathered
// through L1/L2, which the Morton reorder keeps hot.  Forces are assigned (fused clear)
// or accumulated in fp64, and velocity-Verlet's second half-kick can be fused into the
// epilogue (saves one streaming pass over v and f per step).
#include "internal.h"
#include "device_math.cuh"

namespace meso {

struct SoA3 { double *c[3]; };
struct SoA3c { const double *c[3]; };
struct Virial { double *c[6]; };

__device__ __forceinline__ size_t slot(int i, int k, int n_col)
{
    return (size_t)((i & ~31) + (k & 31)) * (size_t)n_col + (size_t)((k >> 5) * 32 + (i & 31));
}

// ------------------------------------------------------------------ pack
__global__ void __launch_bounds__(256) k_pack(SoA3c x, SoA3c v, const int *__restrict__ type, const int *__restrict__ tag,
                                              float4 *__restrict__ coord4, float4 *__restrict__ veloc4,
                                              const Counts *__restrict__ cnt, Box box, uint32_t seed_now, int range)
{
    int beg = 0, end = 0;
    if (range & MESO_LOCAL) end = cnt->nlocal;
    if ((range & MESO_LOCAL) == 0) beg = cnt->nlocal;
    if (range & MESO_GHOST) end = cnt->nlocal + cnt->nghost;
    for (int i = beg + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
        float4 c, w;
        c.x = (float)(x.c[0][i] - box.centre[0]); c.y = (float)(x.c[1][i] - box.centre[1]); c.z = (float)(x.c[2][i] - box.centre[2]);
        c.w = __int_as_float(type[i] - 1);
        w.x = (float)v.c[0][i]; w.y = (float)v.c[1][i]; w.z = (float)v.c[2][i];
        w.w = __uint_as_float(signature(seed_now, tag[i], w.x, w.y, w.z));
        coord4[i] = c; veloc4[i] = w;
    }
}

// ------------------------------------------------------------------ force, fp32 (dpd/fast/meso)
template <int EV>
__global__ void __launch_bounds__(128) k_dpd_sp(const float4 *__restrict__ coord4, const float4 *__restrict__ veloc4,
                                                const int *__restrict__ pair_count, const int *__restrict__ pair_table, SoA3 f,
                                                SoA3 v, Virial vir, double *__restrict__ e_pair, const int *__restrict__ mask,
                                                const int *__restrict__ type, const double *__restrict__ mass,
                                                const float *__restrict__ coeff, const Counts *__restrict__ cnt, int n_col,
                                                int n_type, float dt_inv_sqrt, int range, int accumulate, int fuse_final,
                                                double dtf, int groupbit)
{
    extern __shared__ float cf[];
    for (int p = threadIdx.x; p < n_type * n_type * NCOEFF; p += blockDim.x) cf[p] = coeff[p];
    __syncthreads();
    const int p_beg = (range & MESO_BULK) ? 0 : cnt->n_bulk;
    const int p_end = (range & MESO_BORDER) ? cnt->nlocal : cnt->n_bulk;
    // warp-aligned start so that lanes keep matching the 32-wide tiles of the table
    for (int i = (p_beg & ~31) + blockIdx.x * blockDim.x + threadIdx.x; i < p_end; i += gridDim.x * blockDim.x) {
        if (i < p_beg) continue;
        const float4 c1 = coord4[i], v1 = veloc4[i];
        const uint32_t t1 = __float_as_uint(c1.w), s1 = __float_as_uint(v1.w);
        const int n_pair = pair_count[i];
        float fx = 0.f, fy = 0.f, fz = 0.f;
        float vr0 = 0.f, vr1 = 0.f, vr2 = 0.f, vr3 = 0.f, vr4 = 0.f, vr5 = 0.f, energy = 0.f;
        for (int k = 0; k < n_pair; k++) {
            const int j = __ldcs(pair_table + slot(i, k, n_col));
            const float4 c2 = coord4[j];
            const float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z;
            const float rsq = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            const float *kk = cf + (t1 * n_type + __float_as_uint(c2.w)) * NCOEFF;
            if (rsq < kk[P_CUTSQ] && rsq >= 1.0E-20f) {
                const float4 v2 = veloc4[j];
                const float rn = gaussian_sp(s1, __float_as_uint(v2.w));
                const float rinv = rsqrtf(rsq);
                const float r = rsq * rinv;
                const float dvx = v1.x - v2.x, dvy = v1.y - v2.y, dvz = v1.z - v2.z;
                const float dot = dx * dvx + dy * dvy + dz * dvz;
                const float wc = 1.0f - r * kk[P_CUTINV];
                const float ew = kk[P_EXPW];
                const float wr = (ew == 1.0f) ? wc : powf(wc, ew);      // powf(x,1) == x exactly
                float fpair = kk[P_A0] * wc - (kk[P_GAMMA] * wr * wr * dot * rinv) + (kk[P_SIGMA] * wr * rn * dt_inv_sqrt);
                fpair *= rinv;
                fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
                if (EV) {
                    vr0 += dx * dx * fpair; vr1 += dy * dy * fpair; vr2 += dz * dz * fpair;
                    vr3 += dx * dy * fpair; vr4 += dx * dz * fpair; vr5 += dy * dz * fpair;
                    energy += 0.5f * kk[P_A0] * kk[P_CUT] * wc * wc;
                }
            }
        }
        double Fx = fx, Fy = fy, Fz = fz;
        if (accumulate) { Fx += f.c[0][i]; Fy += f.c[1][i]; Fz += f.c[2][i]; }
        f.c[0][i] = Fx; f.c[1][i] = Fy; f.c[2][i] = Fz;
        if (EV) {
            const float h = 0.5f;
            double w0 = vr0 * h, w1 = vr1 * h, w2 = vr2 * h, w3 = vr3 * h, w4 = vr4 * h, w5 = vr5 * h;
            if (accumulate) { w0 += vir.c[0][i]; w1 += vir.c[1][i]; w2 += vir.c[2][i]; w3 += vir.c[3][i]; w4 += vir.c[4][i]; w5 += vir.c[5][i]; }
            vir.c[0][i] = w0; vir.c[1][i] = w1; vir.c[2][i] = w2; vir.c[3][i] = w3; vir.c[4][i] = w4; vir.c[5][i] = w5;
            e_pair[i] = energy * h;                                   // assigned, UM/pair_dpd_fast_meso.cu:187
        }
        if (fuse_final && (mask[i] & groupbit)) {                     // gpu_fix_NVE_final_integrate, UM/fix_nve_meso.cu:157-178
            const double dtfm = __dmul_rn(dtf, rcp_nr(mass[type[i]]));
            v.c[0][i] = __fma_rn(dtfm, Fx, v.c[0][i]);
            v.c[1][i] = __fma_rn(dtfm, Fy, v.c[1][i]);
            v.c[2][i] = __fma_rn(dtfm, Fz, v.c[2][i]);
        }
    }
}

// ------------------------------------------------------------------ force, fp64 (dpd/meso)
template <int EV>
__global__ void __launch_bounds__(128) k_dpd_dp(const float4 *__restrict__ coord4, const float4 *__restrict__ veloc4,
                                                const int *__restrict__ pair_count, const int *__restrict__ pair_table, SoA3 f,
                                                SoA3 v, Virial vir, double *__restrict__ e_pair, const int *__restrict__ mask,
                                                const int *__restrict__ type, const double *__restrict__ mass,
                                                const double *__restrict__ coeff, const Counts *__restrict__ cnt, int n_col,
                                                int n_type, double dt_inv_sqrt, int range, int accumulate, int fuse_final,
                                                double dtf, int groupbit)
{
    extern __shared__ double cd[];
    for (int p = threadIdx.x; p < n_type * n_type * NCOEFF; p += blockDim.x) cd[p] = coeff[p];
    __syncthreads();
    const int p_beg = (range & MESO_BULK) ? 0 : cnt->n_bulk;
    const int p_end = (range & MESO_BORDER) ? cnt->nlocal : cnt->n_bulk;
    for (int i = (p_beg & ~31) + blockIdx.x * blockDim.x + threadIdx.x; i < p_end; i += gridDim.x * blockDim.x) {
        if (i < p_beg) continue;
        const float4 c1 = coord4[i], v1 = veloc4[i];
        const uint32_t t1 = __float_as_uint(c1.w), s1 = __float_as_uint(v1.w);
        const int n_pair = pair_count[i];
        double fx = 0., fy = 0., fz = 0.;
        double vr0 = 0., vr1 = 0., vr2 = 0., vr3 = 0., vr4 = 0., vr5 = 0., energy = 0.;
        for (int k = 0; k < n_pair; k++) {
            const int j = __ldcs(pair_table + slot(i, k, n_col));
            const float4 c2 = coord4[j];
            // fp32 differences widened to fp64: f3u members are r32 (UM/type_meso.h:17-31, UM/pair_dpd_meso.cu:136-139)
            const double dx = (double)__fsub_rn(c1.x, c2.x), dy = (double)__fsub_rn(c1.y, c2.y), dz = (double)__fsub_rn(c1.z, c2.z);
            const double rsq = __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
            const double *kk = cd + (t1 * n_type + __float_as_uint(c2.w)) * NCOEFF;
            if (rsq < kk[P_CUTSQ] && rsq >= 1.0E-20) {
                const float4 v2 = veloc4[j];
                const double rn = gaussian_dp(s1, __float_as_uint(v2.w));
                const double rinv = rsqrt(rsq);
                const double r = rsq * rinv;
                const double dvx = (double)__fsub_rn(v1.x, v2.x), dvy = (double)__fsub_rn(v1.y, v2.y), dvz = (double)__fsub_rn(v1.z, v2.z);
                const double dot = __fma_rn(dz, dvz, __fma_rn(dy, dvy, __dmul_rn(dx, dvx)));
                const double wc = 1.0 - r * kk[P_CUTINV];
                const double wr = pow_poly(wc, kk[P_EXPW]);
                double fpair = kk[P_A0] * wc - (kk[P_GAMMA] * wr * wr * dot * rinv) + (kk[P_SIGMA] * wr * rn * dt_inv_sqrt);
                fpair *= rinv;
                fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
                if (EV) {
                    vr0 += dx * dx * fpair; vr1 += dy * dy * fpair; vr2 += dz * dz * fpair;
                    vr3 += dx * dy * fpair; vr4 += dx * dz * fpair; vr5 += dy * dz * fpair;
                    energy += 0.5 * kk[P_A0] * kk[P_CUT] * wc * wc;
                }
            }
        }
        double Fx = fx, Fy = fy, Fz = fz;
        if (accumulate) { Fx += f.c[0][i]; Fy += f.c[1][i]; Fz += f.c[2][i]; }
        f.c[0][i] = Fx; f.c[1][i] = Fy; f.c[2][i] = Fz;
        if (EV) {
            double w0 = vr0 * 0.5, w1 = vr1 * 0.5, w2 = vr2 * 0.5, w3 = vr3 * 0.5, w4 = vr4 * 0.5, w5 = vr5 * 0.5;
            if (accumulate) { w0 += vir.c[0][i]; w1 += vir.c[1][i]; w2 += vir.c[2][i]; w3 += vir.c[3][i]; w4 += vir.c[4][i]; w5 += vir.c[5][i]; }
            vir.c[0][i] = w0; vir.c[1][i] = w1; vir.c[2][i] = w2; vir.c[3][i] = w3; vir.c[4][i] = w4; vir.c[5][i] = w5;
            e_pair[i] = energy * 0.5;
        }
        if (fuse_final && (mask[i] & groupbit)) {
            const double dtfm = __dmul_rn(dtf, rcp_nr(mass[type[i]]));
            v.c[0][i] = __fma_rn(dtfm, Fx, v.c[0][i]);
            v.c[1][i] = __fma_rn(dtfm, Fy, v.c[1][i]);
            v.c[2][i] = __fma_rn(dtfm, Fz, v.c[2][i]);
        }
    }
}

// ------------------------------------------------------------------ parity helpers (A8, A9)
__global__ void k_eval_gaussian(int n, const uint32_t *si, const uint32_t *sj, float *osp, double *odp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (osp) osp[i] = gaussian_sp(si[i], sj[i]);
    if (odp) odp[i] = gaussian_dp(si[i], sj[i]);
}
__global__ void k_eval_math(int fn, int n, const double *a, const double *b, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = a[i], r = 0;
    switch (fn) {
    case 0: r = rsqrt_nr(x); break;
    case 1: r = rcp_nr(x); break;
    case 2: r = log2_frac(x); break;
    case 3: r = exp2_frac(x); break;
    case 4: r = sinpi_poly(x); break;
    case 5: r = cospi_poly(x); break;
    case 6: r = sqrt_nr(x); break;
    case 7: r = pow_poly(x, b[i]); break;
    }
    out[i] = r;
}
__global__ void k_eval_log2u(int n, const uint32_t *a, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = log2u(a[i]);
}

// ------------------------------------------------------------------ host drivers
uint32_t seed_now(const meso_ctx *ctx) { return host_premix_tea64((uint32_t)ctx->seed, (uint32_t)ctx->ntimestep); }

int launch_pack(meso_ctx *ctx, int range)
{
    SoA3c x, v;
    for (int d = 0; d < 3; d++) { x.c[d] = ctx->x[d].p; v.c[d] = ctx->v[d].p; }
    k_pack<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(x, v, ctx->type.p, ctx->tag.p, ctx->coord4.p, ctx->veloc4.p, ctx->d_counts, ctx->box,
                                                   seed_now(ctx), range);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_pair(meso_ctx *ctx, int range, int evflag, bool accumulate, bool fuse_final, int groupbit)
{
    SoA3 f, v;
    Virial vir;
    for (int d = 0; d < 3; d++) { f.c[d] = ctx->f[d].p; v.c[d] = ctx->v[d].p; }
    for (int q = 0; q < 6; q++) vir.c[q] = ctx->virial.p + (size_t)q * ctx->cap;
    const int nt = ctx->ntypes;
    const double dtf = 0.5 * ctx->dt;   // ftm2v = 1 in lj units (FixNVEMeso::init, UM/fix_nve_meso.cu:42-46)
    const int grid = grid_for(ctx, 16);
    if (ctx->precision == MESO_SP) {
        size_t sh = sizeof(float) * nt * nt * NCOEFF;
        float dtis = (float)(1.0 / sqrt(ctx->dt));
        if (evflag)
            k_dpd_sp<1><<<grid, 128, sh, ctx->stream>>>(ctx->coord4.p, ctx->veloc4.p, ctx->pair_count.p, ctx->pair_table.p, f, v, vir,
                                                      ctx->e_pair.p, ctx->mask.p, ctx->type.p, ctx->mass_dev.p, ctx->coeff_sp.p,
                                                      ctx->d_counts, ctx->n_col, nt, dtis, range, accumulate, fuse_final, dtf, groupbit);
        else
            k_dpd_sp<0><<<grid, 128, sh, ctx->stream>>>(ctx->coord4.p, ctx->veloc4.p, ctx->pair_count.p, ctx->pair_table.p, f, v, vir,
                                                      ctx->e_pair.p, ctx->mask.p, ctx->type.p, ctx->mass_dev.p, ctx->coeff_sp.p,
                                                      ctx->d_counts, ctx->n_col, nt, dtis, range, accumulate, fuse_final, dtf, groupbit);
    } else {
        size_t sh = sizeof(double) * nt * nt * NCOEFF;
        double dtis = 1.0 / sqrt(ctx->dt);
        if (evflag)
            k_dpd_dp<1><<<grid, 128, sh, ctx->stream>>>(ctx->coord4.p, ctx->veloc4.p, ctx->pair_count.p, ctx->pair_table.p, f, v, vir,
                                                      ctx->e_pair.p, ctx->mask.p, ctx->type.p, ctx->mass_dev.p, ctx->coeff_dp.p,
                                                      ctx->d_counts, ctx->n_col, nt, dtis, range, accumulate, fuse_final, dtf, groupbit);
        else
            k_dpd_dp<0><<<grid, 128, sh, ctx->stream>>>(ctx->coord4.p, ctx->veloc4.p, ctx->pair_count.p, ctx->pair_table.p, f, v, vir,
                                                      ctx->e_pair.p, ctx->mask.p, ctx->type.p, ctx->mass_dev.p, ctx->coeff_dp.p,
                                                      ctx->d_counts, ctx->n_col, nt, dtis, range, accumulate, fuse_final, dtf, groupbit);
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int eval_gaussian(meso_ctx *ctx, int n, const uint32_t *si, const uint32_t *sj, float *osp, double *odp)
{
    k_eval_gaussian<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, si, sj, osp, odp);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}
int eval_math(meso_ctx *ctx, int fn, int n, const double *a, const double *b, double *out)
{
    k_eval_math<<<(n + 255) / 256, 256, 0, ctx->stream>>>(fn, n, a, b, out);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}
int eval_log2u(meso_ctx *ctx, int n, const uint32_t *a, double *out)
{
    k_eval_log2u<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, a, out);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
