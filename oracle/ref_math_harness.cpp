// ref_math_harness.cpp -- host-side shims so that the reference's own
// __host__ __device__ integer/fp64 math (extracted at build time, by line range,
// from /root/reference/src/USER-MESO/math_meso.h into oracle/_ref/) compiles with
// g++ and can be called from the oracle tests.  TEST INFRASTRUCTURE ONLY.
// No reference source is stored in this repository: the Makefile extracts
// math_meso.h:12-24 (constants) and :143-505 (functions) into _ref/ on the fly.
#include <cmath>
#include <cstring>
#include <cstdint>
typedef unsigned int uint;
#define __host__
#define __device__
#define __inline__ inline
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __hiloint2double(int hi, int lo)
{ unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &u, 8); return d; }
static inline void __double2hiloint(double x, int &hi, int &lo)
{ unsigned long long u; std::memcpy(&u, &x, 8); hi = (int)(u >> 32); lo = (int)(u & 0xFFFFFFFFu); }
static inline uint __float_as_uint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
static inline int __clz(uint x) { return x ? __builtin_clz(x) : 32; }
static inline float sinpif(float x) { return (float)std::sin(M_PI * (double)x); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline uint max(uint a, int b) { return a > (uint)b ? a : (uint)b; }
static inline double max(double a, double b) { return std::fmax(a, b); }
static inline double min(double a, double b) { return std::fmin(a, b); }
static inline float max(float a, float b) { return std::fmaxf(a, b); }
static inline float min(float a, float b) { return std::fminf(a, b); }
using std::floor;

static inline void pred_swap(bool pred, int &u, int &v) { int I = pred ? u : v, J = pred ? v : u; u = I; v = J; }

#include "_ref/ref_math_extract.h"

extern "C" {
void ref_tea(uint *v0, uint *v1, int rounds)
{
    switch (rounds) {
    case 1: __TEA_core<1>(*v0, *v1); break;
    case 4: __TEA_core<4>(*v0, *v1); break;
    case 8: __TEA_core<8>(*v0, *v1); break;
    case 16: __TEA_core<16>(*v0, *v1); break;
    case 64: __TEA_core<64>(*v0, *v1); break;
    default: *v0 = *v1 = 0;
    }
}
uint ref_premix16(uint a, uint b) { return premix_TEA<16>(a, b); }
uint ref_premix64(uint a, uint b) { return premix_TEA<64>(a, b); }
uint ref_interleave3(uint i, uint j, uint k) { return interleave3(i, j, k); }
uint ref_morton(uint i, uint j, uint k) { return morton_encode(i, j, k); }
uint ref_mantissa(float u, float v, float w) { return __mantissa(u, v, w); }
int ref_clamp(int i, int lo, int hi) { return clamp(i, lo, hi); }
double ref_gaussian_dp(uint si, uint sj) { return gaussian_TEA<4>(si > sj, si, sj); }
float ref_gaussian_sp(uint si, uint sj) { return gaussian_TEA_fast<4>(si > sj, si, sj); }
double ref_rsqrt(double x) { return __rsqrt(x); }
double ref_sqrtd(double x) { return __sqrtd(x); }
double ref_rcp(double x) { return __rcp(x); }
double ref_log2d_frac(double x) { return __log2d_frac(x); }
double ref_exp2d_frac(double x) { return __exp2d_frac(x); }
double ref_powd(double a, double b) { return __powd(a, b); }
double ref_sinpi(double x) { return __sinpi(x); }
double ref_cospi(double x) { return __cospi(x); }
double ref_log2u(uint x) { return __log2u(x); }
}
