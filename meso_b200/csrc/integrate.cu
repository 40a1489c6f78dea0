// integrate.cu -- velocity-Verlet halves, force clear, AoS<->SoA transfer kernels and reductions.
//
// Reference path:
//   gpu_fix_NVE_init_intgrate<0> / gpu_fix_NVE_final_integrate   UM/fix_nve_meso.cu:62-95,157-178   (A12)
//   MesoAtomVec::force_clear                                     UM/atom_vec_meso.cu:325-336        (A16)
//   gpu_deinterleave / gpu_interleave (host AoS <-> device SoA)  UM/atom_vec_meso.h:44-88           (A2)
//   gpu_eK_scalar + gpu_reduce_sum_host                          UM/compute_temp_meso.cu:58-101, UM/math_meso.h:677-692 (A17)
// The first half-kick + drift also emits this step's packed float4 views and signatures
// (the reference re-reads x, v, type and tag in a separate gpu_merge_xvt pass).
#include "internal.h"
#include "device_math.cuh"
#include "fix_device.cuh"

namespace meso {

struct SoA3 { double *c[3]; };
struct SoA3c { const double *c[3]; };

template <int PACK>
__global__ void __launch_bounds__(256) k_initial_integrate(SoA3 x, SoA3 v, SoA3c f, const int *__restrict__ mask,
                                                           const int *__restrict__ type, const int *__restrict__ tag,
                                                           const double *__restrict__ mass, float4 *__restrict__ coord4,
                                                           float4 *__restrict__ veloc4, const Counts *__restrict__ cnt, Box box,
                                                           double dtf, double dtv, int groupbit, uint32_t seed_now)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int ty = type[i];
        double xx[3], vv[3];
#pragma unroll
        for (int d = 0; d < 3; d++) { xx[d] = x.c[d][i]; vv[d] = v.c[d][i]; }
        if (mask[i] & groupbit) {
            const double dtfm = __dmul_rn(dtf, rcp_nr(mass[ty]));
#pragma unroll
            for (int d = 0; d < 3; d++) {
                vv[d] = __fma_rn(dtfm, f.c[d][i], vv[d]);
                xx[d] = __fma_rn(dtv, vv[d], xx[d]);
                v.c[d][i] = vv[d];
                x.c[d][i] = xx[d];
            }
        }
        if (PACK) {
            float4 c, w;
            c.x = (float)(xx[0] - box.centre[0]); c.y = (float)(xx[1] - box.centre[1]); c.z = (float)(xx[2] - box.centre[2]);
            c.w = __int_as_float(ty - 1);
            w.x = (float)vv[0]; w.y = (float)vv[1]; w.z = (float)vv[2];
            w.w = __uint_as_float(signature(seed_now, tag[i], w.x, w.y, w.z));
            coord4[i] = c; veloc4[i] = w;
        }
    }
}

// Step boundary of the fused run loop (api.cu:meso_run): the second half-kick of step k-1 and the first half-kick + drift
// (+ pack + signature) of step k read the same force, so one streaming pass does both.  The force comes from the fp32
// accumulator of the pair-once kernel (facc, src_acc) or from the fp64 arrays; the source is cleared for the next
// reduction and/or mirrored into f when the run returns to the caller.  Each half is the reference's own fma
// (UM/fix_nve_meso.cu:83-92,171-176), so results equal the unfused sequence bit for bit.
// FIX: wall fixes are registered (fix.cu).  Their end_of_step bounce-forward of step k-1 sits between the two kicks and their
// pre_exchange bounce (re-neighbouring steps) follows the drift -- per atom the same operation sequence as the reference's
// separate kernels (UM/mvv_meso.cu:262-275,398-402), so the fusion is exact.
template <int PACK, int FIX>
__global__ void __launch_bounds__(256) k_step_integrate(SoA3 x, SoA3 v, SoA3 f, float4 *__restrict__ facc, const int *__restrict__ mask,
                                                        const int *__restrict__ type, const int *__restrict__ tag,
                                                        const double *__restrict__ mass, float4 *__restrict__ coord4,
                                                        float4 *__restrict__ veloc4, const Counts *__restrict__ cnt, Box box, double dtf,
                                                        double dtv, int groupbit, uint32_t seed_now, int do_final, int do_initial,
                                                        int src_acc, int zero_src, int write_f, FixList fl, int bounce_after)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // all loads first: the arrays are not __restrict__ (x, v, f are read and written), so a store in the middle
        // would serialise the remaining loads behind it
        const int ty = type[i], mk = mask[i];
        const int tg = PACK ? tag[i] : 0;
        double ff[3], xx[3], vv[3];
        if (src_acc) {
            const float4 a = facc[i];
            ff[0] = (double)a.x; ff[1] = (double)a.y; ff[2] = (double)a.z;
        } else {
#pragma unroll
            for (int d = 0; d < 3; d++) ff[d] = f.c[d][i];
        }
#pragma unroll
        for (int d = 0; d < 3; d++) { vv[d] = v.c[d][i]; xx[d] = (do_initial || FIX) ? x.c[d][i] : 0.; }
        const double ms = mass[ty];
        if (FIX) {
            const bool grp = (mk & groupbit) != 0;
            const double dtfm = __dmul_rn(dtf, rcp_nr(ms));
            if (do_final) {
                if (grp) {
#pragma unroll
                    for (int d = 0; d < 3; d++) vv[d] = __fma_rn(dtfm, ff[d], vv[d]);
                }
                fix_bounce_all(fl, box, mk, xx, vv);                  // end_of_step of the previous step
            }
            if (do_initial) {
                if (grp) {
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        vv[d] = __fma_rn(dtfm, ff[d], vv[d]);
                        xx[d] = __fma_rn(dtv, vv[d], xx[d]);
                    }
                }
                if (bounce_after) fix_bounce_all(fl, box, mk, xx, vv);   // pre_exchange of this (re-neighbouring) step
            }
        } else if (mk & groupbit) {
            const double dtfm = __dmul_rn(dtf, rcp_nr(ms));
#pragma unroll
            for (int d = 0; d < 3; d++) {
                if (do_final) vv[d] = __fma_rn(dtfm, ff[d], vv[d]);
                if (do_initial) {
                    vv[d] = __fma_rn(dtfm, ff[d], vv[d]);
                    xx[d] = __fma_rn(dtv, vv[d], xx[d]);
                }
            }
        }
        // stores
        if (src_acc && zero_src) facc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (write_f || (!src_acc && zero_src)) {
#pragma unroll
            for (int d = 0; d < 3; d++) f.c[d][i] = write_f ? ff[d] : 0.;
        }
        if (FIX || (mk & groupbit)) {
#pragma unroll
            for (int d = 0; d < 3; d++) { v.c[d][i] = vv[d]; if (do_initial || FIX) x.c[d][i] = xx[d]; }
        }
        if (PACK) {
            float4 c, w;
            c.x = (float)(xx[0] - box.centre[0]); c.y = (float)(xx[1] - box.centre[1]); c.z = (float)(xx[2] - box.centre[2]);
            c.w = __int_as_float(ty - 1);
            w.x = (float)vv[0]; w.y = (float)vv[1]; w.z = (float)vv[2];
            w.w = __uint_as_float(signature(seed_now, tg, w.x, w.y, w.z));
            coord4[i] = c; veloc4[i] = w;
        }
    }
}

__global__ void __launch_bounds__(256) k_final_integrate(SoA3 v, SoA3c f, const int *__restrict__ mask, const int *__restrict__ type,
                                                         const double *__restrict__ mass, const Counts *__restrict__ cnt, double dtf,
                                                         int groupbit)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (mask[i] & groupbit) {
            const double dtfm = __dmul_rn(dtf, rcp_nr(mass[type[i]]));
#pragma unroll
            for (int d = 0; d < 3; d++) v.c[d][i] = __fma_rn(dtfm, f.c[d][i], v.c[d][i]);
        }
    }
}

__global__ void __launch_bounds__(256) k_clear(SoA3 f, double *__restrict__ virial, size_t cap, const Counts *__restrict__ cnt,
                                               int range, int vflag)
{
    const int beg = (range & MESO_BULK) ? 0 : cnt->n_bulk;
    const int end = (range & MESO_BORDER) ? cnt->nlocal : cnt->n_bulk;
    for (int i = beg + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
        f.c[0][i] = 0.; f.c[1][i] = 0.; f.c[2][i] = 0.;
        if (vflag)
            for (int q = 0; q < 6; q++) virial[q * cap + i] = 0.;
    }
}

// ------------------------------------------------------------------ AoS <-> SoA
__global__ void __launch_bounds__(256) k_deinterleave3(const double *__restrict__ aos, SoA3 soa, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        soa.c[0][i] = aos[3 * (size_t)i]; soa.c[1][i] = aos[3 * (size_t)i + 1]; soa.c[2][i] = aos[3 * (size_t)i + 2];
    }
}
__global__ void __launch_bounds__(256) k_interleave3(SoA3c soa, double *__restrict__ aos, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        aos[3 * (size_t)i] = soa.c[0][i]; aos[3 * (size_t)i + 1] = soa.c[1][i]; aos[3 * (size_t)i + 2] = soa.c[2][i];
    }
}
__global__ void __launch_bounds__(256) k_fill_defaults(int *tag, int *type, int *mask, int *image, int n, int set_tag, int set_type,
                                                       int set_mask, int set_image)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (set_tag) tag[i] = i + 1;
        if (set_type) type[i] = 1;
        if (set_mask) mask[i] = 1;
        if (set_image) image[i] = (512 << 20) | (512 << 10) | 512;
    }
}
__global__ void __launch_bounds__(256) k_zero3(SoA3 a, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { a.c[0][i] = 0.; a.c[1][i] = 0.; a.c[2][i] = 0.; }
}

// ------------------------------------------------------------------ deterministic two-stage reductions (no atomics)
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NV], double *__restrict__ partial)
{
    __shared__ double sm[NV][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) sm[q][w] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.;
        for (int ww = 0; ww < (int)(blockDim.x >> 5); ww++) s += sm[threadIdx.x][ww];
        partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) k_ke_partial(SoA3c v, const int *__restrict__ mask, const int *__restrict__ type,
                                                    const double *__restrict__ mass, const Counts *__restrict__ cnt, int groupbit,
                                                    double *__restrict__ partial)
{
    const int n = cnt->nlocal;
    double acc[2] = {0., 0.};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (mask[i] & groupbit) {
            const double a = v.c[0][i], b = v.c[1][i], c = v.c[2][i];
            acc[0] += mass[type[i]] * (a * a + b * b + c * c);     // gpu_eK_scalar, UM/compute_temp_meso.cu:73
            acc[1] += 1.0;
        }
    }
    block_reduce_store<2>(acc, partial);
}

__global__ void __launch_bounds__(256) k_virial_partial(const double *__restrict__ virial, const double *__restrict__ e_pair, size_t cap,
                                                        const Counts *__restrict__ cnt, double *__restrict__ partial)
{
    const int n = cnt->nlocal;
    double acc[7] = {0., 0., 0., 0., 0., 0., 0.};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int q = 0; q < 6; q++) acc[q] += virial[q * cap + i];
        acc[6] += e_pair[i];
    }
    block_reduce_store<7>(acc, partial);
}

__global__ void __launch_bounds__(256) k_reduce_final(const double *__restrict__ partial, int nblocks, int nv, double *__restrict__ out)
{
    __shared__ double sm[8];
    for (int q = 0; q < nv; q++) {
        double s = 0.;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(size_t)q * nblocks + b];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.;
            for (int w = 0; w < 8; w++) t += sm[w];
            out[q] = t;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host drivers
static inline SoA3 soa(DevBuf<double> *b) { SoA3 s; for (int d = 0; d < 3; d++) s.c[d] = b[d].p; return s; }
static inline SoA3c soac(DevBuf<double> *b) { SoA3c s; for (int d = 0; d < 3; d++) s.c[d] = b[d].p; return s; }

int launch_initial_integrate(meso_ctx *ctx, int groupbit, bool pack)
{
    const double dtv = ctx->dt, dtf = 0.5 * ctx->dt * ctx->ftm2v;   // FixNVEMeso::init, UM/fix_nve_meso.cu:42-46
    if (pack)
        k_initial_integrate<1><<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soa(ctx->x), soa(ctx->v), soac(ctx->f), ctx->mask.p, ctx->type.p,
                                                                       ctx->tag.p, ctx->mass_dev.p, ctx->coord4.p, ctx->veloc4.p,
                                                                       ctx->d_counts, ctx->box, dtf, dtv, groupbit, seed_now(ctx));
    else
        k_initial_integrate<0><<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soa(ctx->x), soa(ctx->v), soac(ctx->f), ctx->mask.p, ctx->type.p,
                                                                       ctx->tag.p, ctx->mass_dev.p, ctx->coord4.p, ctx->veloc4.p,
                                                                       ctx->d_counts, ctx->box, dtf, dtv, groupbit, 0u);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_step_integrate(meso_ctx *ctx, int groupbit, bool do_final, bool do_initial, bool pack, bool src_acc, bool zero_src, bool write_f,
                          bool bounce_after)
{
    const double dtv = ctx->dt, dtf = 0.5 * ctx->dt * ctx->ftm2v;
    pack = pack && do_initial;
    const bool fix = ctx->fixes.nbounce > 0;
#define MESO_STEP_ARGS soa(ctx->x), soa(ctx->v), soa(ctx->f), ctx->facc.p, ctx->mask.p, ctx->type.p, ctx->tag.p, ctx->mass_dev.p, ctx->coord4.p, \
                       ctx->veloc4.p, ctx->d_counts, ctx->box, dtf, dtv, groupbit, pack ? seed_now(ctx) : 0u, do_final, do_initial, src_acc,        \
                       zero_src, write_f, ctx->fixes, bounce_after
    const int g = grid_for(ctx, 8);
    if (pack && fix) k_step_integrate<1, 1><<<g, 256, 0, LS(ctx->stream)>>>(MESO_STEP_ARGS);
    else if (pack) k_step_integrate<1, 0><<<g, 256, 0, LS(ctx->stream)>>>(MESO_STEP_ARGS);
    else if (fix) k_step_integrate<0, 1><<<g, 256, 0, LS(ctx->stream)>>>(MESO_STEP_ARGS);
    else k_step_integrate<0, 0><<<g, 256, 0, LS(ctx->stream)>>>(MESO_STEP_ARGS);
#undef MESO_STEP_ARGS
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_final_integrate(meso_ctx *ctx, int groupbit)
{
    k_final_integrate<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soa(ctx->v), soac(ctx->f), ctx->mask.p, ctx->type.p, ctx->mass_dev.p,
                                                              ctx->d_counts, 0.5 * ctx->dt * ctx->ftm2v, groupbit);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_clear(meso_ctx *ctx, int range, int vflag)
{
    k_clear<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soa(ctx->f), ctx->virial.p, ctx->cap, ctx->d_counts, range, vflag);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_ke(meso_ctx *ctx, int groupbit, double *mv2, double *count)
{
    const int nb = grid_for(ctx, 4);
    if (!ctx->partial.reserve((size_t)nb * 8)) { ctx->err = "reduce: out of device memory"; return MESO_ECUDA; }
    k_ke_partial<<<nb, 256, 0, LS(ctx->stream)>>>(soac(ctx->v), ctx->mask.p, ctx->type.p, ctx->mass_dev.p, ctx->d_counts, groupbit, ctx->partial.p);
    k_reduce_final<<<1, 256, 0, LS(ctx->stream)>>>(ctx->partial.p, nb, 2, ctx->partial.p + (size_t)nb * 7);
    MESO_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->partial.p + (size_t)nb * 7, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    *mv2 = ctx->h_result[0];
    *count = ctx->h_result[1];
    return MESO_OK;
}

int launch_virial_sum(meso_ctx *ctx, double out7[7])
{
    const int nb = grid_for(ctx, 4);
    if (!ctx->partial.reserve((size_t)nb * 8)) { ctx->err = "reduce: out of device memory"; return MESO_ECUDA; }
    k_virial_partial<<<nb, 256, 0, LS(ctx->stream)>>>(ctx->virial.p, ctx->e_pair.p, ctx->cap, ctx->d_counts, ctx->partial.p);
    k_reduce_final<<<1, 256, 0, LS(ctx->stream)>>>(ctx->partial.p, nb, 7, ctx->partial.p + (size_t)nb * 7);
    MESO_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->partial.p + (size_t)nb * 7, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < 7; q++) out7[q] = ctx->h_result[q];
    return MESO_OK;
}

// used by api.cu
int launch_deinterleave3(meso_ctx *ctx, const double *aos, DevBuf<double> *dst, int n)
{
    k_deinterleave3<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(aos, soa(dst), n);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}
int launch_interleave3(meso_ctx *ctx, DevBuf<double> *src, double *aos, int n)
{
    k_interleave3<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soac(src), aos, n);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}
int launch_fill_defaults(meso_ctx *ctx, int n, int st, int sy, int sm, int si)
{
    k_fill_defaults<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p, n, st, sy, sm, si);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}
int launch_zero3(meso_ctx *ctx, DevBuf<double> *a, int n)
{
    k_zero3<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soa(a), n);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
