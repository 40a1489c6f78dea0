// device_math.cuh -- integer RNG core, Morton codes and branch-free transcendentals
// for the sm_100a DPD kernels.  Algorithms follow the reference (cited per function,
// paths relative to /root/reference/src, UM/ = USER-MESO/); the code is written for
// Blackwell: everything is __forceinline__, FMA chains are spelled with __fma_rn so
// that ptxas cannot re-associate them, integer work maps to LOP3/IADD3/SHF.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace meso {

// ---------------------------------------------------------------- integer core
// Tiny Encryption Algorithm rounds with the reference's fixed key (UM/math_meso.h:444-456).
template <int N>
__device__ __forceinline__ void tea(uint32_t &v0, uint32_t &v1)
{
    uint32_t sum = 0;
#pragma unroll
    for (int r = 0; r < N; r++) {
        sum += 0x9E3779B9u;
        v0 += ((v1 << 4) + 0xA341316Cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xC8013EA4u);
        v1 += ((v0 << 4) + 0xAD90777Du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7E95761Eu);
    }
}
template <int N>
__device__ __forceinline__ uint32_t premix_tea(uint32_t v0, uint32_t v1)   // UM/math_meso.h:460-464
{
    tea<N>(v0, v1);
    return v0 ^ v1;
}
// host copy for seed_now = premix_TEA<64>(seed, ntimestep), UM/pair_dpd_meso.cu:268-270
static inline uint32_t host_premix_tea64(uint32_t v0, uint32_t v1)
{
    uint32_t sum = 0;
    for (int r = 0; r < 64; r++) {
        sum += 0x9E3779B9u;
        v0 += ((v1 << 4) + 0xA341316Cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xC8013EA4u);
        v1 += ((v0 << 4) + 0xAD90777Du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7E95761Eu);
    }
    return v0 ^ v1;
}

// 11-bit -> every-third-bit spread (UM/math_meso.h:166-173)
__host__ __device__ __forceinline__ uint32_t bit_space3(uint32_t x)
{
    x = (x | (x << 12)) & 0x00FC003Fu;
    x = (x | (x << 6)) & 0x381C0E07u;
    x = (x | (x << 4)) & 0x190C8643u;
    x = (x | (x << 2)) & 0x49249249u;
    return x;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t i, uint32_t j, uint32_t k)  // UM/math_meso.h:175-183
{
    return bit_space3(i) | (bit_space3(j) << 1) | (bit_space3(k) << 2);
}
// 11 mantissa bits (12..22) of each fp32 velocity component, interleaved (UM/math_meso.h:436-442)
__device__ __forceinline__ uint32_t mantissa3(float u, float v, float w)
{
    return morton3((__float_as_uint(u) & 0x7FF000u) >> 12, (__float_as_uint(v) & 0x7FF000u) >> 12,
                   (__float_as_uint(w) & 0x7FF000u) >> 12);
}
// per-particle, per-step signature (UM/atom_vec_meso.cu:164)
__device__ __forceinline__ uint32_t signature(uint32_t seed_now, int tag, float vx, float vy, float vz)
{
    return seed_now ^ premix_tea<16>(__brev((uint32_t)tag), mantissa3(vx, vy, vz));
}

// ---------------------------------------------------------------- fp64 transcendentals (A9)
// Magic-constant seed + 4 Newton steps, Chebyshev polynomials in FMA form (UM/math_meso.h:204-424).
__device__ __forceinline__ double two_to_n(int n) { return __hiloint2double((1023 + n) << 20, 0); }

__device__ __forceinline__ double rsqrt_nr(double x)     // UM/math_meso.h:210-221
{
    double r = __longlong_as_double(0x5FE660FCB5422422LL - (__double_as_longlong(x) >> 1));
    double x2m = __dmul_rn(x, -0.5);
#pragma unroll
    for (int i = 0; i < 4; i++) r = __dmul_rn(r, __fma_rn(__dmul_rn(r, r), x2m, 1.5));
    return r;
}
__device__ __forceinline__ double sqrt_nr(double x) { return __dmul_rn(x, rsqrt_nr(x)); }  // :223-226

__device__ __forceinline__ double rcp_nr(double x)       // UM/math_meso.h:230-238
{
    double xi = __longlong_as_double(0x7FDE62361B1C4042LL - __double_as_longlong(x));
#pragma unroll
    for (int i = 0; i < 4; i++) xi = __dsub_rn(xi, __dmul_rn(__fma_rn(x, xi, -1.), xi));
    return xi;
}

#define MESO_SQRT_2 1.4142135623730950488
#define MESO_1_OVER_SQ2 7.0710678118654757274E-1
#define MESO_LN_2 6.9314718055994528623E-1

__device__ __forceinline__ double log2_frac(double x)    // x in [1,2], UM/math_meso.h:264-280
{
    bool pred = x > MESO_SQRT_2;
    x = __dmul_rn(x, pred ? 0.5 : MESO_1_OVER_SQ2);
    double z = __dmul_rn(__dsub_rn(x, 1.), rcp_nr(__dadd_rn(x, 1.)));
    double y = __dmul_rn(__dmul_rn(z, z), 33.9705627484771406);
    double s = 4.0928048937567843469E-12;
    s = __fma_rn(s, y, 1.4374842194796670219E-10);
    s = __fma_rn(s, y, 5.7988453014506741861E-9);
    s = __fma_rn(s, y, 2.4074128088151586443E-7);
    s = __fma_rn(s, y, 1.0514733588011180538E-5);
    s = __fma_rn(s, y, 5.0006798065881969549E-4);
    s = __fma_rn(s, y, 2.8312651192953993354E-2);
    s = __fma_rn(s, y, 2.8853900817779268114E+0);
    return __fma_rn(z, s, pred ? 1.0 : 0.5);
}
__device__ __forceinline__ double exp2_frac(double x)    // x in [0,1], UM/math_meso.h:308-324
{
    double s = 6.3026908837748924689E-10;
    s = __fma_rn(s, x, 6.5379419072372670333E-9);
    s = __fma_rn(s, x, 1.0258347084283025531E-7);
    s = __fma_rn(s, x, 1.3207676270599404858E-6);
    s = __fma_rn(s, x, 1.5253232908458899497E-5);
    s = __fma_rn(s, x, 1.5403509189194102748E-4);
    s = __fma_rn(s, x, 1.3333558738165095559E-3);
    s = __fma_rn(s, x, 9.6181290971755593396E-3);
    s = __fma_rn(s, x, 5.5504108665909870679E-2);
    s = __fma_rn(s, x, 2.4022650695904222220E-1);
    s = __fma_rn(s, x, 6.9314718055994653980E-1);
    s = __fma_rn(s, x, 9.9999999999999999572E-1);
    return s;
}
__device__ __forceinline__ double pow_poly(double a, double b)   // UM/math_meso.h:332-345
{
    int hi = __double2hiint(a), lo = __double2loint(a);
    double I = (double)((hi >> 20) - 1023);
    double F = log2_frac(__hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo));
    double II = floor(__dmul_rn(b, __dadd_rn(I, F)));
    return __dmul_rn(two_to_n((int)II), exp2_frac(__fma_rn(b, F, __fma_rn(b, I, -II))));
}
__device__ __forceinline__ double sinpi_poly(double x)   // x in [-1,1] -> see UM/math_meso.h:357-370
{
    x = __dsub_rn(__dmul_rn(2.0, x), 1.0);
    x = __dmul_rn(x, x);
    double s = 4.49220128554338954E-7;
    s = __fma_rn(s, x, -2.51721958850906157E-5);
    s = __fma_rn(s, x, 9.19240000031795430E-4);
    s = __fma_rn(s, x, -2.08634736828917670E-2);
    s = __fma_rn(s, x, 2.53669506722714547E-1);
    s = __fma_rn(s, x, -1.23370055006260170E+0);
    s = __fma_rn(s, x, 9.99999999999249900E-1);
    return s;
}
__device__ __forceinline__ double cospi_poly(double x)   // UM/math_meso.h:380-392
{
    x = __dsub_rn(__dmul_rn(2.0, x), 1.0);
    double x2 = __dmul_rn(x, x);
    double s = 3.41817283473266926E-6;
    s = __fma_rn(s, x2, -1.60217135750921262E-4);
    s = __fma_rn(s, x2, 4.68162024021793872E-3);
    s = __fma_rn(s, x2, -7.96925872866600517E-2);
    s = __fma_rn(s, x2, 6.45964092644060746E-1);
    s = __fma_rn(s, x2, -1.57079632662144460E+0);
    return __dmul_rn(s, x);
}
__device__ __forceinline__ double log2u(uint32_t x)      // log2(x / 2^32), UM/math_meso.h:401-424
{
    int I = 31 - __clz(x);
    double xd = __dmul_rn((double)x, two_to_n(-I));
    bool pred = xd > MESO_SQRT_2;
    xd = __dmul_rn(xd, pred ? 0.5 : MESO_1_OVER_SQ2);
    double z = __dmul_rn(__dsub_rn(xd, 1.), rcp_nr(__dadd_rn(xd, 1.)));
    double y = __dmul_rn(__dmul_rn(z, z), 33.9705627484771406);
    double s = 2.55854634203511155E-7;
    s = __fma_rn(s, y, 1.05013262724846015E-5);
    s = __fma_rn(s, y, 5.00072802051539862E-4);
    s = __fma_rn(s, y, 2.83126505877817866E-2);
    s = __fma_rn(s, y, 2.88539008179006374E+0);
    return __fma_rn(z, s, __dadd_rn(pred ? 1.0 : 0.5, (double)(I - 32)));
}

// ---------------------------------------------------------------- per-pair Gaussians (A8)
// Symmetric in (i,j): the larger signature goes first, 4 TEA rounds, Box-Muller, clamp +-4.
__device__ __forceinline__ double gaussian_dp(uint32_t si, uint32_t sj)   // UM/math_meso.h:466-474
{
    bool pred = si > sj;
    uint32_t v0 = pred ? si : sj, v1 = pred ? sj : si;
    tea<4>(v0, v1);
    double f = __dmul_rn(cospi_poly(__dmul_rn((double)(v0 & 0x7FFFFFFFu), 4.6566128730773925781E-10)),
                         (v0 & 0x80000000u) ? 1.0 : -1.0);
    double r = sqrt_nr(__dmul_rn(-2.0 * MESO_LN_2, log2u(max(v1, 1u))));
    return fmax(-4.0, fmin(__dmul_rn(r, f), 4.0));
}
__device__ __forceinline__ float gaussian_sp(uint32_t si, uint32_t sj)    // UM/math_meso.h:476-484
{
    bool pred = si > sj;
    uint32_t v0 = pred ? si : sj, v1 = pred ? sj : si;
    tea<4>(v0, v1);
    float f = sinpif(__fmul_rn((float)(int)v0, 4.6566128730773925781E-10f));
    float r = sqrtf(__fmul_rn(-2.0f * (float)MESO_LN_2, log2f(__fmul_rn((float)v1, 2.3283064365386962891E-10f))));
    return fmaxf(-4.0f, fminf(__fmul_rn(r, f), 4.0f));
}

// MUFU-based, non-branching fp32 variant used by the fp32 force kernel: lg2/sin/sqrt.approx replace
// log2f/sinpif/sqrtf (sin.approx: 2^-21.4 abs on [-pi,pi]; lg2.approx: 2^-22 abs) -> |error| < ~5e-6 on a
// value of O(1), inside the 1e-5 relative force tolerance.  Same clamp / NaN semantics as the reference.
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sin_approx(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float gaussian_sp_fast(uint32_t si, uint32_t sj)
{
    bool pred = si > sj;
    uint32_t v0 = pred ? si : sj, v1 = pred ? sj : si;
    tea<4>(v0, v1);
    const float f = sin_approx(__fmul_rn((float)(int)v0, 3.14159265358979323846f * 4.6566128730773925781E-10f));
    const float r = sqrt_approx(__fmul_rn(-2.0f * (float)MESO_LN_2, lg2_approx(__fmul_rn((float)v1, 2.3283064365386962891E-10f))));
    return fmaxf(-4.0f, fminf(__fmul_rn(r, f), 4.0f));
}

// saturating double -> int truncation, then clamp to [lo, hi)  (UM/math_meso.h:155-158)
__device__ __forceinline__ int clamp_rz(double v, int lo, int hi) { return max(lo, min(__double2int_rz(v), hi - 1)); }

}  // namespace meso
