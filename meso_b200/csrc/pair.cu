// pair.cu -- packing (+ per-step signatures) and the DPD pair-force kernels.
//
// Reference path:
//   gpu_merge_xvt / dp2sp_merged      UM/atom_vec_meso.cu:142-192       (A7)
//   gpu_dpd_fast<ev>                  UM/pair_dpd_fast_meso.cu:91-205   (A10, fp32)
//   gpu_dpd<ev>                       UM/pair_dpd_meso.cu:91-205        (A11, fp64 on the fp32-packed inputs)
//   compute / compute_bulk / compute_border  UM/pair_dpd_meso.cu:241-266
//
// Blackwell design.  One thread owns one local particle and a warp owns one 32-wide tile of the
// tile-transposed table, so index loads are 128-byte coalesced.  The reference evaluates the
// expensive part (4 TEA rounds + Box-Muller + force) inside the divergent `if (rsq < cutsq)` of its
// neighbor loop: at rho = 4 only 45 % of the stored neighbors are in range, so most issue slots of
// the heavy code are predicated off.  Here the loop is split in two warp-uniform phases:
//   scan : walk the row, test the distance, push in-range j's into a per-lane FIFO in shared
//          memory (column layout [depth][32 lanes]: bank == lane, conflict-free, no atomics);
//   drain: every lane pops its own FIFO, so the heavy code runs max_lane(hits) times instead of
//          once per stored neighbor, in list order (same summation order as the reference).
// Forces are assigned (fused clear) or accumulated in fp64, and velocity-Verlet's second
// half-kick can be fused into the epilogue (one streaming pass over v and f less per step).
// The kernel is issue-bound, not HBM-bound: the fp32 path uses MUFU-based
// lg2/sin/rsqrt/sqrt (non-branching, ~1e-6 abs on the Gaussian, inside the 1e-5 force tolerance).
#include "internal.h"
#include "device_math.cuh"
#include <algorithm>
#include <cstdlib>

namespace meso {

struct SoA3 { double *c[3]; };
struct SoA3c { const double *c[3]; };
struct Virial { double *c[6]; };

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

// ------------------------------------------------------------------ force
constexpr int PAIR_THREADS = 128;
constexpr int QDEPTH = 20;       // FIFO slots per lane (mean in-range count at rho = 4 is 16.8); overflow -> early drain
constexpr int SCAN_CHUNK = 4;

template <typename REAL> struct Acc {
    REAL fx = 0, fy = 0, fz = 0, v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, e = 0;
};

// one in-range pair, fp32 (UM/pair_dpd_fast_meso.cu:143-173)
template <int EV>
__device__ __forceinline__ void pair_sp(const float4 c1, const float4 v1, const float4 c2, const float4 v2, const float *__restrict__ kk,
                                        float dtis, Acc<float> &a)
{
    const float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z;
    const float rsq = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    const float rn = gaussian_sp_fast(__float_as_uint(v1.w), __float_as_uint(v2.w));
    const float rinv = rsqrt_approx(rsq);
    const float r = rsq * rinv;
    const float dvx = v1.x - v2.x, dvy = v1.y - v2.y, dvz = v1.z - v2.z;
    const float dot = dx * dvx + dy * dvy + dz * dvz;
    const float wc = 1.0f - r * kk[P_CUTINV];
    const float ew = kk[P_EXPW];
    const float wr = (ew == 1.0f) ? wc : __powf(wc, ew);           // pow(x,1) == x; other exponents via ex2(ew*lg2(wc))
    float fpair = kk[P_A0] * wc - (kk[P_GAMMA] * wr * wr * dot * rinv) + (kk[P_SIGMA] * wr * rn * dtis);
    fpair *= rinv;
    a.fx += dx * fpair; a.fy += dy * fpair; a.fz += dz * fpair;
    if (EV) {
        a.v0 += dx * dx * fpair; a.v1 += dy * dy * fpair; a.v2 += dz * dz * fpair;
        a.v3 += dx * dy * fpair; a.v4 += dx * dz * fpair; a.v5 += dy * dz * fpair;
        a.e += 0.5f * kk[P_A0] * kk[P_CUT] * wc * wc;
    }
}

// one in-range pair, fp64 arithmetic on the fp32-packed inputs (UM/pair_dpd_meso.cu:136-173)
template <int EV>
__device__ __forceinline__ void pair_dp(const float4 c1, const float4 v1, const float4 c2, const float4 v2, const double *__restrict__ kk,
                                        double dtis, Acc<double> &a)
{
    // fp32 differences widened to fp64: f3u members are r32 (UM/type_meso.h:17-31)
    const double dx = (double)__fsub_rn(c1.x, c2.x), dy = (double)__fsub_rn(c1.y, c2.y), dz = (double)__fsub_rn(c1.z, c2.z);
    const double rsq = __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
    const double rn = gaussian_dp(__float_as_uint(v1.w), __float_as_uint(v2.w));
    const double rinv = rsqrt(rsq);
    const double r = rsq * rinv;
    const double dvx = (double)__fsub_rn(v1.x, v2.x), dvy = (double)__fsub_rn(v1.y, v2.y), dvz = (double)__fsub_rn(v1.z, v2.z);
    const double dot = __fma_rn(dz, dvz, __fma_rn(dy, dvy, __dmul_rn(dx, dvx)));
    const double wc = 1.0 - r * kk[P_CUTINV];
    const double ew = kk[P_EXPW];
    const double wr = (ew == 1.0) ? wc : pow_poly(wc, ew);         // pow(x,1) == x (the polynomial pow gives x(1 +- ~1e-15))
    double fpair = kk[P_A0] * wc - (kk[P_GAMMA] * wr * wr * dot * rinv) + (kk[P_SIGMA] * wr * rn * dtis);
    fpair *= rinv;
    a.fx += dx * fpair; a.fy += dy * fpair; a.fz += dz * fpair;
    if (EV) {
        a.v0 += dx * dx * fpair; a.v1 += dy * dy * fpair; a.v2 += dz * dz * fpair;
        a.v3 += dx * dy * fpair; a.v4 += dx * dz * fpair; a.v5 += dy * dz * fpair;
        a.e += 0.5 * kk[P_A0] * kk[P_CUT] * wc * wc;
    }
}

template <typename REAL, int EV>
__global__ void __launch_bounds__(PAIR_THREADS) k_dpd(const float4 *__restrict__ coord4, const float4 *__restrict__ veloc4,
                                                      const int *__restrict__ pair_count, const int *__restrict__ pair_table, SoA3 f,
                                                      SoA3 v, Virial vir, double *__restrict__ e_pair, const int *__restrict__ mask,
                                                      const int *__restrict__ type, const double *__restrict__ mass,
                                                      const REAL *__restrict__ coeff, const Counts *__restrict__ cnt, int n_col,
                                                      int n_type, REAL dt_inv_sqrt, int range, int accumulate, int fuse_final,
                                                      double dtf, int groupbit)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    REAL *cf = reinterpret_cast<REAL *>(smem_raw);
    const int ncf = n_type * n_type * NCOEFF;
    int *queue = reinterpret_cast<int *>(smem_raw + (((size_t)ncf * sizeof(REAL) + 15) & ~(size_t)15)) + (threadIdx.x >> 5) * (QDEPTH * 32) +
                 (threadIdx.x & 31);
    for (int p = threadIdx.x; p < ncf; p += blockDim.x) cf[p] = coeff[p];
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int p_beg = (range & MESO_BULK) ? 0 : cnt->n_bulk;
    const int p_end = (range & MESO_BORDER) ? cnt->nlocal : cnt->n_bulk;
    // warp-aligned start: lanes keep matching the 32-wide tiles of the table; whole warps iterate together
    for (int i = (p_beg & ~31) + blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < p_end; i += gridDim.x * blockDim.x) {
        const bool active = i >= p_beg && i < p_end;
        float4 c1 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = c1;
        int n_pair = 0;
        if (active) { c1 = coord4[i]; v1 = veloc4[i]; n_pair = pair_count[i]; }
        const uint32_t t1 = __float_as_uint(c1.w);
        const int nmax = __reduce_max_sync(full, n_pair);
        const int *tp = pair_table + (size_t)(i & ~31) * (size_t)n_col + (i & 31);   // slot(i, k) = tp[(k&31)*n_col + (k>>5)*32]
        const REAL *cf1 = cf + t1 * n_type * NCOEFF;       // coefficient rows of my type
        const int jsafe = active ? i : 0;                   // gather target of masked-off slots
        Acc<REAL> acc;
        int qlen = 0;

        // ---- drain: pop my FIFO; the next entry's gathers are issued before the current pair's math
        auto drain = [&]() {
            const int m = __reduce_max_sync(full, qlen);
            int jn = qlen > 0 ? queue[0] : jsafe;
            float4 c2n = coord4[jn], v2n = veloc4[jn];
            for (int q = 0; q < m; q++) {
                const float4 c2 = c2n, v2 = v2n;
                jn = (q + 1 < qlen) ? queue[(q + 1) * 32] : jsafe;
                c2n = coord4[jn]; v2n = veloc4[jn];
                if (q < qlen) {
                    const REAL *kk = cf1 + __float_as_uint(c2.w) * NCOEFF;
                    if constexpr (sizeof(REAL) == 4) pair_sp<EV>(c1, v1, c2, v2, kk, dt_inv_sqrt, acc);
                    else pair_dp<EV>(c1, v1, c2, v2, kk, dt_inv_sqrt, acc);
                }
            }
            qlen = 0;
        };

        // ---- scan: branch-free chunks of SCAN_CHUNK slots; the next chunk's index loads are in flight while
        //      this chunk's float4 gathers and distance tests run.  k0 is a multiple of SCAN_CHUNK (a divisor of 32),
        //      so a chunk never straddles a 32-slot tile: slot(k0+u) = tp[off(k0) + u*n_col].  Slots beyond a lane's
        //      n_pair are read (in-bounds, stale data) but their j is replaced by jsafe before the gather.
        auto load_chunk = [&](int k0, int (&jj)[SCAN_CHUNK]) {
            const int *p = tp + (size_t)((k0 & 31) * n_col + (k0 >> 5) * 32);
#pragma unroll
            for (int u = 0; u < SCAN_CHUNK; u++) {
                const int raw = __ldcs(p + (size_t)u * n_col);
                jj[u] = (k0 + u < n_pair) ? raw : -1;
            }
        };
        int jnext[SCAN_CHUNK];
        if (nmax > 0) load_chunk(0, jnext);
        for (int k0 = 0; k0 < nmax; k0 += SCAN_CHUNK) {
            int jc[SCAN_CHUNK];
#pragma unroll
            for (int u = 0; u < SCAN_CHUNK; u++) jc[u] = jnext[u];
            if (k0 + SCAN_CHUNK < nmax) load_chunk(k0 + SCAN_CHUNK, jnext);
            if (__any_sync(full, qlen > QDEPTH - SCAN_CHUNK)) drain();
            float4 c2[SCAN_CHUNK];
#pragma unroll
            for (int u = 0; u < SCAN_CHUNK; u++) c2[u] = coord4[jc[u] >= 0 ? jc[u] : jsafe];
#pragma unroll
            for (int u = 0; u < SCAN_CHUNK; u++) {
                const REAL cutsq = cf1[__float_as_uint(c2[u].w) * NCOEFF + P_CUTSQ];
                bool hit;
                if constexpr (sizeof(REAL) == 4) {
                    const float dx = c1.x - c2[u].x, dy = c1.y - c2[u].y, dz = c1.z - c2[u].z;
                    const float rsq = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                    hit = rsq < cutsq && rsq >= 1.0E-20f;            // UM/pair_dpd_fast_meso.cu:143
                } else {
                    const double dx = (double)__fsub_rn(c1.x, c2[u].x), dy = (double)__fsub_rn(c1.y, c2[u].y), dz = (double)__fsub_rn(c1.z, c2[u].z);
                    const double rsq = __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
                    hit = rsq < cutsq && rsq >= 1.0E-20;             // UM/pair_dpd_meso.cu:143
                }
                hit = hit && jc[u] >= 0;
                if (hit) queue[qlen * 32] = jc[u];
                qlen += hit ? 1 : 0;
            }
        }
        drain();

        if (active) {
            double Fx = acc.fx, Fy = acc.fy, Fz = acc.fz;
            if (accumulate) { Fx += f.c[0][i]; Fy += f.c[1][i]; Fz += f.c[2][i]; }
            f.c[0][i] = Fx; f.c[1][i] = Fy; f.c[2][i] = Fz;
            if (EV) {
                const REAL h = (REAL)0.5;
                double w0 = acc.v0 * h, w1 = acc.v1 * h, w2 = acc.v2 * h, w3 = acc.v3 * h, w4 = acc.v4 * h, w5 = acc.v5 * h;
                if (accumulate) { w0 += vir.c[0][i]; w1 += vir.c[1][i]; w2 += vir.c[2][i]; w3 += vir.c[3][i]; w4 += vir.c[4][i]; w5 += vir.c[5][i]; }
                vir.c[0][i] = w0; vir.c[1][i] = w1; vir.c[2][i] = w2; vir.c[3][i] = w3; vir.c[4][i] = w4; vir.c[5][i] = w5;
                e_pair[i] = acc.e * h;                                    // assigned, UM/pair_dpd_fast_meso.cu:187
            }
            if (fuse_final && (mask[i] & groupbit)) {                     // gpu_fix_NVE_final_integrate, UM/fix_nve_meso.cu:157-178
                const double dtfm = __dmul_rn(dtf, rcp_nr(mass[type[i]]));
                v.c[0][i] = __fma_rn(dtfm, Fx, v.c[0][i]);
                v.c[1][i] = __fma_rn(dtfm, Fy, v.c[1][i]);
                v.c[2][i] = __fma_rn(dtfm, Fz, v.c[2][i]);
            }
        }
    }
}

// ------------------------------------------------------------------ force, each local pair evaluated ONCE
// The reference evaluates every pair from both sides (newton off, full list): F_ij and F_ji come from two
// separate runs of TEA + Box-Muller + force.  Every term of the pair force is antisymmetric bit for bit in
// (i,j): d = r_i - r_j negates exactly, rsq, rinv, d.dv and the Gaussian (keyed on (max,min) of the two
// signatures) are identical from both sides -- so F_ji == -F_ij to the last bit and one evaluation serves both
// atoms.  The full table stays the only list; row i owns the pair (i,j), j local, iff
//      (i+j) odd ? i < j : i > j          (balanced: every atom owns ~half of its in-range neighbors)
// and pairs with a ghost j are always evaluated by i (the ghost's owner evaluates the mirror pair itself,
// exactly as before: no reverse communication).  The build stores the owned entries as a prefix of the row.  The j side is updated with ONE 16-byte vector reduction
// (red.global.add.v4.f32 -> REDG.E.ADD.F32x4, resolved in L2) into a float4 accumulator per atom that the
// integrator consumes and clears; the fp64 style reduces into the fp64 force arrays (REDG.E.ADD.F64).
// Only the order of the fp32 additions differs from the two-sided kernel (each addend is bit-identical).
constexpr int QD1 = 24;          // per-lane FIFO depth (mean owned in-range count at rho = 4 is 8.4, max over a warp ~14)

template <typename REAL> struct Coef1 { REAL cut, cutsq, cutinv, expw, a0, gamma, sigma; };

// facc[j] += {a, b, c, 0} if pred (predicated REDG.E.ADD.F32x4: no branch around the reduction)
__device__ __forceinline__ void red_add_v4_if(bool pred, float4 *p, float a, float b, float c)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n\t}"
                 ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(0.f), "r"((int)pred) : "memory");
}

__device__ __forceinline__ void sts32(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int lds32(unsigned addr) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// ONE_TYPE: a single atom type -> the 7 coefficients arrive as kernel parameters (constant bank operands, no shared-memory
// lookups per candidate / per pair); POW1: every exponent is 1 -> wr = wc (pow(x,1) == x, result-identical).
// GMODE (experiment knob, MESO_PAIR_TEX): bit 0 = scan-phase coordinate gathers, bit 1 = drain-phase gathers go through the
// texture data pipe (tex.1d.v4.f32.s32 on a linear texture object over the same buffers) instead of the LSU data pipe
template <typename REAL, bool ONE_TYPE, bool POW1, int GMODE>
__global__ void __launch_bounds__(PAIR_THREADS) k_dpd_once(cudaTextureObject_t tex_c, cudaTextureObject_t tex_v,
                                                           const float4 *__restrict__ coord4, const float4 *__restrict__ veloc4,
                                                           const int *__restrict__ owned_count, const int *__restrict__ pair_table,
                                                           float4 *__restrict__ facc, SoA3 f, const REAL *__restrict__ coeff,
                                                           const Coef1<REAL> k1, const Counts *__restrict__ cnt, int n_col, int n_type,
                                                           REAL dt_inv_sqrt, int range, int far)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int queue_s[PAIR_THREADS / 32][QD1][32];
    REAL *cf = reinterpret_cast<REAL *>(smem_raw);
    if (!ONE_TYPE) {
        const int ncf = n_type * n_type * NCOEFF;
        for (int p = threadIdx.x; p < ncf; p += blockDim.x) cf[p] = coeff[p];
        __syncthreads();
    }
    // 32-bit shared-window address of this lane's FIFO column (entry q lives at qbase + q*128: bank == lane)
    const unsigned qbase = (unsigned)__cvta_generic_to_shared(&queue_s[threadIdx.x >> 5][0][threadIdx.x & 31]);
    const unsigned full = 0xffffffffu;
    const int nlocal = cnt->nlocal;
    const int p_beg = (range & MESO_BULK) ? 0 : cnt->n_bulk;
    const int p_end = (range & MESO_BORDER) ? nlocal : cnt->n_bulk;
    const ptrdiff_t step4 = (ptrdiff_t)SCAN_CHUNK * n_col, wrap = 32 - (ptrdiff_t)32 * n_col;
    for (int i = (p_beg & ~31) + blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < p_end; i += gridDim.x * blockDim.x) {
        const bool active = i >= p_beg && i < p_end;
        // masked-off table slots (beyond the row, not owned by this lane) gather the `far` slot: one coordinate record at
        // ~3.4e38 kept behind the atoms, so rsq = +inf fails the cutoff test without a validity flag, and all masked
        // lanes of a warp read the same 16 bytes (one L1 tag lookup instead of a scattered line each)
        const int self = active ? i : far;
        const float4 c1 = coord4[self], v1 = veloc4[self];
        const int n_pair = active ? owned_count[i] : 0;      // length of the owned prefix of the row
        const int nmax = __reduce_max_sync(full, n_pair);
        const REAL *cf1 = cf + (ONE_TYPE ? 0 : __float_as_uint(c1.w) * n_type * NCOEFF);
        REAL fx = 0, fy = 0, fz = 0;
        unsigned qp = qbase;                                // FIFO write address

        auto one_pair = [&](const float4 c2, const float4 v2, int j) {
            if constexpr (sizeof(REAL) == 4) {
                const float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z;
                const float rsq = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                const float rn = gaussian_sp_fast(__float_as_uint(v1.w), __float_as_uint(v2.w));
                const float rinv = rsqrt_approx(rsq);
                const float r = rsq * rinv;
                const float dvx = v1.x - v2.x, dvy = v1.y - v2.y, dvz = v1.z - v2.z;
                const float dot = dx * dvx + dy * dvy + dz * dvz;
                float cutinv, ew, a0, gamma, sigma;
                if (ONE_TYPE) { cutinv = k1.cutinv; ew = k1.expw; a0 = k1.a0; gamma = k1.gamma; sigma = k1.sigma; }
                else {
                    const float *kk = cf1 + __float_as_uint(c2.w) * NCOEFF;
                    cutinv = kk[P_CUTINV]; ew = kk[P_EXPW]; a0 = kk[P_A0]; gamma = kk[P_GAMMA]; sigma = kk[P_SIGMA];
                }
                const float wc = 1.0f - r * cutinv;
                const float wr = (POW1 || ew == 1.0f) ? wc : __powf(wc, ew);
                float fpair = a0 * wc - (gamma * wr * wr * dot * rinv) + (sigma * wr * rn * dt_inv_sqrt);
                const float nf = -(fpair * rinv);             // -fpair: the j side's addend is computed directly, the i side subtracts it
                const float qx = dx * nf, qy = dy * nf, qz = dz * nf;
                red_add_v4_if(j < nlocal, facc + j, qx, qy, qz);
                fx -= qx; fy -= qy; fz -= qz;
            } else {
                const double dx = (double)__fsub_rn(c1.x, c2.x), dy = (double)__fsub_rn(c1.y, c2.y), dz = (double)__fsub_rn(c1.z, c2.z);
                const double rsq = __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
                const double rn = gaussian_dp(__float_as_uint(v1.w), __float_as_uint(v2.w));
                const double rinv = rsqrt(rsq);
                const double r = rsq * rinv;
                const double dvx = (double)__fsub_rn(v1.x, v2.x), dvy = (double)__fsub_rn(v1.y, v2.y), dvz = (double)__fsub_rn(v1.z, v2.z);
                const double dot = __fma_rn(dz, dvz, __fma_rn(dy, dvy, __dmul_rn(dx, dvx)));
                double cutinv, ew, a0, gamma, sigma;
                if (ONE_TYPE) { cutinv = k1.cutinv; ew = k1.expw; a0 = k1.a0; gamma = k1.gamma; sigma = k1.sigma; }
                else {
                    const double *kk = cf1 + __float_as_uint(c2.w) * NCOEFF;
                    cutinv = kk[P_CUTINV]; ew = kk[P_EXPW]; a0 = kk[P_A0]; gamma = kk[P_GAMMA]; sigma = kk[P_SIGMA];
                }
                const double wc = 1.0 - r * cutinv;
                // pow(x, 1) == x: the reference's polynomial pow returns x(1 +- ~1e-15) there, far inside the 1e-12 bar,
                // and costs ~40 of the ~150 fp64 operations of a pair
                const double wr = (POW1 || ew == 1.0) ? wc : pow_poly(wc, ew);
                double fpair = a0 * wc - (gamma * wr * wr * dot * rinv) + (sigma * wr * rn * dt_inv_sqrt);
                fpair *= rinv;
                const double px = dx * fpair, py = dy * fpair, pz = dz * fpair;
                if (j < nlocal) { atomicAdd(f.c[0] + j, -px); atomicAdd(f.c[1] + j, -py); atomicAdd(f.c[2] + j, -pz); }
                fx += px; fy += py; fz += pz;
            }
        };

        // ---- drain: pop my FIFO; the next entry's two gathers are in flight during the current pair's math
        auto drain = [&]() {
            const int qlen = (int)(qp - qbase) >> 7;
            const int m = __reduce_max_sync(full, qlen);
            int jn = qlen > 0 ? lds32(qbase) : far;
            float4 c2n, v2n;
            if (GMODE & 2) { c2n = tex1Dfetch<float4>(tex_c, jn); v2n = tex1Dfetch<float4>(tex_v, jn); }
            else { c2n = coord4[jn]; v2n = veloc4[jn]; }
#pragma unroll 2
            for (int q = 0; q < m; q++) {
                const float4 c2 = c2n, v2 = v2n;
                const int j = jn;
                jn = (q + 1 < qlen) ? lds32(qbase + (q + 1) * 128) : far;
                if (GMODE & 2) { c2n = tex1Dfetch<float4>(tex_c, jn); v2n = tex1Dfetch<float4>(tex_v, jn); }
                else { c2n = coord4[jn]; v2n = veloc4[jn]; }
                if (q < qlen) one_pair(c2, v2, j);
            }
            qp = qbase;
        };

        // ---- scan: SCAN_CHUNK slots per step; slot(k) = tp[(k&31)*n_col + (k>>5)*32], walked with one running pointer.
        // The build (neighbor.cu) puts the entries this row owns -- ghost j, or (i+j) odd ? i<j : i>j -- at the front of
        // the row, so the scan covers the owned prefix only and needs no ownership test.
        const int *pk = pair_table + (size_t)(i & ~31) * (size_t)n_col + (i & 31);
        auto load_chunk = [&](int k0, int (&jj)[SCAN_CHUNK]) {
            if (k0 < nmax) {
                const int kn = k0 - n_pair;
#pragma unroll
                for (int u = 0; u < SCAN_CHUNK; u++) {
                    const int j = __ldcs(pk + (ptrdiff_t)u * n_col);
                    jj[u] = (kn + u) < 0 ? j : far;
                }
                pk += step4;
                if (((k0 + SCAN_CHUNK) & 31) == 0) pk += wrap;
            } else {
#pragma unroll
                for (int u = 0; u < SCAN_CHUNK; u++) jj[u] = far;
            }
        };
        auto test_chunk = [&](const int (&jc)[SCAN_CHUNK]) {
            float4 c2[SCAN_CHUNK];
#pragma unroll
            for (int u = 0; u < SCAN_CHUNK; u++) c2[u] = (GMODE & 1) ? tex1Dfetch<float4>(tex_c, jc[u]) : coord4[jc[u]];
#pragma unroll
            for (int u = 0; u < SCAN_CHUNK; u++) {
                REAL cutsq;
                if (ONE_TYPE) cutsq = k1.cutsq; else cutsq = cf1[__float_as_uint(c2[u].w) * NCOEFF + P_CUTSQ];
                bool hit;
                if constexpr (sizeof(REAL) == 4) {
                    const float dx = c1.x - c2[u].x, dy = c1.y - c2[u].y, dz = c1.z - c2[u].z;
                    const float rsq = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                    hit = rsq < cutsq && rsq >= 1.0E-20f;            // UM/pair_dpd_fast_meso.cu:143
                } else {
                    const double dx = (double)__fsub_rn(c1.x, c2[u].x), dy = (double)__fsub_rn(c1.y, c2[u].y), dz = (double)__fsub_rn(c1.z, c2[u].z);
                    const double rsq = __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
                    hit = rsq < cutsq && rsq >= 1.0E-20;             // UM/pair_dpd_meso.cu:143
                }
                if (hit) { sts32(qp, jc[u]); qp += 128; }
            }
        };
        // two chunks per iteration with explicit double buffering: the index loads of the chunk after next are in flight
        // while the current chunk's gathers and distance tests run (no register shuffling between iterations)
        int ja[SCAN_CHUNK], jb[SCAN_CHUNK];
        load_chunk(0, ja);
        int k0 = 0;
        do {
            // scan until the row ends or some lane's FIFO could overflow within the next two chunks (rare)
            while (k0 < nmax && !__any_sync(full, qp > qbase + (QD1 - 2 * SCAN_CHUNK) * 128)) {
                load_chunk(k0 + SCAN_CHUNK, jb);
                test_chunk(ja);
                load_chunk(k0 + 2 * SCAN_CHUNK, ja);
                test_chunk(jb);
                k0 += 2 * SCAN_CHUNK;
            }
            drain();
        } while (k0 < nmax);

        if (active) {
            if constexpr (sizeof(REAL) == 4) red_add_v4_if(true, facc + i, fx, fy, fz);
            else { atomicAdd(f.c[0] + i, fx); atomicAdd(f.c[1] + i, fy); atomicAdd(f.c[2] + i, fz); }
        }
    }
}

// ------------------------------------------------------------------ parity helpers (A8, A9)
__global__ void k_eval_gaussian(int n, const uint32_t *si, const uint32_t *sj, float *osp, double *odp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (osp) osp[i] = gaussian_sp_fast(si[i], sj[i]);      // the variant the fp32 force kernel uses
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
    k_pack<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(x, v, ctx->type.p, ctx->tag.p, ctx->coord4.p, ctx->veloc4.p, ctx->d_counts, ctx->box,
                                                   seed_now(ctx), range);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

template <typename REAL, int EV>
static int launch_pair_t(meso_ctx *ctx, const REAL *coeff, REAL dtis, int range, bool accumulate, bool fuse_final, int groupbit)
{
    SoA3 f, v;
    Virial vir;
    for (int d = 0; d < 3; d++) { f.c[d] = ctx->f[d].p; v.c[d] = ctx->v[d].p; }
    for (int q = 0; q < 6; q++) vir.c[q] = ctx->virial.p + (size_t)q * ctx->cap;
    const int nt = ctx->ntypes;
    const double dtf = 0.5 * ctx->dt * ctx->ftm2v;   // FixNVEMeso::init, UM/fix_nve_meso.cu:42-46
    const size_t sh = (((size_t)nt * nt * NCOEFF * sizeof(REAL) + 15) & ~(size_t)15) + (size_t)(PAIR_THREADS / 32) * QDEPTH * 32 * sizeof(int);
    // one CTA per 128 particles (host-side upper bound of the range); the grid-stride loop covers any excess
    int grid = (int)((nlocal_bound(ctx) + PAIR_THREADS - 1) / PAIR_THREADS) + 1;
    grid = std::max(1, std::min(grid, ctx->sm_count * 4096));
    cudaFuncSetAttribute(k_dpd<REAL, EV>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);   // per device; a hint, cheap to repeat
    k_dpd<REAL, EV><<<grid, PAIR_THREADS, sh, LS(ctx->stream)>>>(ctx->coord4.p, ctx->veloc4.p, ctx->pair_count.p, ctx->pair_table.p, f, v, vir,
                                                           ctx->e_pair.p, ctx->mask.p, ctx->type.p, ctx->mass_dev.p, coeff, ctx->d_counts,
                                                           ctx->n_col, nt, dtis, range, accumulate, fuse_final, dtf, groupbit);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_pair(meso_ctx *ctx, int range, int evflag, bool accumulate, bool fuse_final, int groupbit)
{
    if (ctx->precision == MESO_SP) {
        const float dtis = (float)(1.0 / sqrt(ctx->dt));
        return evflag ? launch_pair_t<float, 1>(ctx, ctx->coeff_sp.p, dtis, range, accumulate, fuse_final, groupbit)
                      : launch_pair_t<float, 0>(ctx, ctx->coeff_sp.p, dtis, range, accumulate, fuse_final, groupbit);
    }
    const double dtis = 1.0 / sqrt(ctx->dt);
    return evflag ? launch_pair_t<double, 1>(ctx, ctx->coeff_dp.p, dtis, range, accumulate, fuse_final, groupbit)
                  : launch_pair_t<double, 0>(ctx, ctx->coeff_dp.p, dtis, range, accumulate, fuse_final, groupbit);
}

template <typename REAL>
static int launch_pair_once_t(meso_ctx *ctx, const REAL *coeff, REAL dtis, int range)
{
    SoA3 f;
    for (int d = 0; d < 3; d++) f.c[d] = ctx->f[d].p;
    const int nt = ctx->ntypes;
    Coef1<REAL> k1;
    k1.cut = (REAL)ctx->coeff[P_CUT]; k1.cutsq = (REAL)ctx->coeff[P_CUTSQ]; k1.cutinv = (REAL)ctx->coeff[P_CUTINV];
    k1.expw = (REAL)ctx->coeff[P_EXPW]; k1.a0 = (REAL)ctx->coeff[P_A0]; k1.gamma = (REAL)ctx->coeff[P_GAMMA]; k1.sigma = (REAL)ctx->coeff[P_SIGMA];
    int grid = (int)((nlocal_bound(ctx) + PAIR_THREADS - 1) / PAIR_THREADS) + 1;
    grid = std::max(1, std::min(grid, ctx->sm_count * 4096));
    bool pow1 = true;
    for (int t = 0; t < nt * nt; t++) pow1 = pow1 && ctx->coeff[(size_t)t * NCOEFF + P_EXPW] == 1.0;
#define MESO_ONCE_ARGS ctx->tex_coord, ctx->tex_veloc, ctx->coord4.p, ctx->veloc4.p, ctx->owned_count.p, ctx->pair_table.p, ctx->facc.p, f, coeff, k1, ctx->d_counts, ctx->n_col, nt, dtis, range, (int)ctx->cap
    const int gmode = ctx->pair_tex;
    if (nt == 1 && pow1) {
        switch (gmode) {
        case 1: k_dpd_once<REAL, true, true, 1><<<grid, PAIR_THREADS, 0, LS(ctx->stream)>>>(MESO_ONCE_ARGS); break;
        case 2: k_dpd_once<REAL, true, true, 2><<<grid, PAIR_THREADS, 0, LS(ctx->stream)>>>(MESO_ONCE_ARGS); break;
        case 3: k_dpd_once<REAL, true, true, 3><<<grid, PAIR_THREADS, 0, LS(ctx->stream)>>>(MESO_ONCE_ARGS); break;
        default: k_dpd_once<REAL, true, true, 0><<<grid, PAIR_THREADS, 0, LS(ctx->stream)>>>(MESO_ONCE_ARGS);
        }
    } else if (nt == 1) k_dpd_once<REAL, true, false, 0><<<grid, PAIR_THREADS, 0, LS(ctx->stream)>>>(MESO_ONCE_ARGS);
    else k_dpd_once<REAL, false, false, 0><<<grid, PAIR_THREADS, (size_t)nt * nt * NCOEFF * sizeof(REAL), LS(ctx->stream)>>>(MESO_ONCE_ARGS);
#undef MESO_ONCE_ARGS
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// each owned pair once, no energy/virial tally; the accumulator (facc for fp32, f for fp64) must be zero on entry
int launch_pair_once(meso_ctx *ctx, int range)
{
    if (ctx->precision == MESO_SP) return launch_pair_once_t<float>(ctx, ctx->coeff_sp.p, (float)(1.0 / sqrt(ctx->dt)), range);
    return launch_pair_once_t<double>(ctx, ctx->coeff_dp.p, 1.0 / sqrt(ctx->dt), range);
}

int eval_gaussian(meso_ctx *ctx, int n, const uint32_t *si, const uint32_t *sj, float *osp, double *odp)
{
    k_eval_gaussian<<<(n + 255) / 256, 256, 0, LS(ctx->stream)>>>(n, si, sj, osp, odp);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}
int eval_math(meso_ctx *ctx, int fn, int n, const double *a, const double *b, double *out)
{
    k_eval_math<<<(n + 255) / 256, 256, 0, LS(ctx->stream)>>>(fn, n, a, b, out);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}
int eval_log2u(meso_ctx *ctx, int n, const uint32_t *a, double *out)
{
    k_eval_log2u<<<(n + 255) / 256, 256, 0, LS(ctx->stream)>>>(n, a, out);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
