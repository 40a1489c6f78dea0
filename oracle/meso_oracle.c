/* meso_oracle.c -- CPU restatement of the USER-MESO DPD time-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see meso_oracle.h).  Plain C99, scalar, one thread.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (contraction is spelled out
 * with fma()/fmaf() wherever the reference's nvcc build would contract and the
 * result feeds an integer decision).
 *
 * Citations: paths relative to /root/reference/src, UM/ = USER-MESO/.
 */
#include "meso_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

static char g_err[512];
const char *orc_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return -1; } while (0)

/* ====================================================================== */
/* integer core                                                           */
/* ====================================================================== */

/* UM/math_meso.h:444-456  __TEA_core<N>, fixed key, delta 9E3779B9, sum starts at 0 */
void orc_tea(uint32_t *pv0, uint32_t *pv1, int rounds)
{
    const uint32_t K0 = 0xA341316Cu, K1 = 0xC8013EA4u, K2 = 0xAD90777Du, K3 = 0x7E95761Eu;
    uint32_t v0 = *pv0, v1 = *pv1, sum = 0;
    for (int r = 0; r < rounds; r++) {
        sum += 0x9E3779B9u;
        v0 += ((v1 << 4) + K0) ^ (v1 + sum) ^ ((v1 >> 5) + K1);
        v1 += ((v0 << 4) + K2) ^ (v0 + sum) ^ ((v0 >> 5) + K3);
    }
    *pv0 = v0; *pv1 = v1;
}

/* UM/math_meso.h:460-464 */
uint32_t orc_premix_tea(uint32_t v0, uint32_t v1, int rounds)
{
    orc_tea(&v0, &v1, rounds);
    return v0 ^ v1;
}

/* UM/math_meso.h:166-173 */
uint32_t orc_bit_space3(uint32_t x)
{
    x = (x | (x << 12)) & 0x00FC003Fu;
    x = (x | (x << 6)) & 0x381C0E07u;
    x = (x | (x << 4)) & 0x190C8643u;
    x = (x | (x << 2)) & 0x49249249u;
    return x;
}

/* UM/math_meso.h:175-183 (morton_encode == interleave3) */
uint32_t orc_interleave3(uint32_t i, uint32_t j, uint32_t k)
{
    return orc_bit_space3(i) | (orc_bit_space3(j) << 1) | (orc_bit_space3(k) << 2);
}

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* UM/math_meso.h:436-442 : 11 mantissa bits (bits 12..22) of each component */
uint32_t orc_mantissa(float u, float v, float w)
{
    uint32_t i = f2u(u) & 0x7FF000u, j = f2u(v) & 0x7FF000u, k = f2u(w) & 0x7FF000u;
    return orc_interleave3(i >> 12, j >> 12, k >> 12);
}

uint32_t orc_brev(uint32_t x)
{
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}

/* UM/pair_dpd_meso.cu:268-270 : premix_TEA<64>(seed, ntimestep) */
uint32_t orc_seed_now(uint32_t seed, uint32_t ntimestep)
{
    return orc_premix_tea(seed, ntimestep, 64);
}

/* UM/atom_vec_meso.cu:164 */
uint32_t orc_signature(uint32_t seed_now, int tag, float vx, float vy, float vz)
{
    return seed_now ^ orc_premix_tea(orc_brev((uint32_t)tag), orc_mantissa(vx, vy, vz), 16);
}

/* ====================================================================== */
/* fp64 transcendentals, UM/math_meso.h:204-424 (FMA chains spelled out)  */
/* ====================================================================== */

static double ll2d(int64_t i) { double d; memcpy(&d, &i, 8); return d; }
static int64_t d2ll(double d) { int64_t i; memcpy(&i, &d, 8); return i; }
static double hilo2d(int32_t hi, int32_t lo)
{
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d; memcpy(&d, &u, 8); return d;
}
/* UM/math_meso.h:204-207 */
static double two_to_n(int n) { return hilo2d((1023 + n) << 20, 0); }

/* UM/math_meso.h:210-221 */
double orc_rsqrt(double x)
{
    double r = ll2d(0x5FE660FCB5422422LL - (d2ll(x) >> 1));
    double x2m = x * -0.5;
    for (int i = 0; i < 4; i++) r *= fma(r * r, x2m, 1.5);
    return r;
}
/* UM/math_meso.h:223-226 */
double orc_sqrtd(double x) { return x * orc_rsqrt(x); }

/* UM/math_meso.h:230-238 */
double orc_rcp(double x)
{
    double xi = ll2d(0x7FDE62361B1C4042LL - d2ll(x));
    for (int i = 0; i < 4; i++) xi -= fma(x, xi, -1.) * xi;
    return xi;
}

#define SQRT_2 1.4142135623730950488
#define ONE_OVER_SQ2 7.0710678118654757274E-1
#define LN_2 6.9314718055994528623E-1

/* UM/math_meso.h:264-280 */
double orc_log2d_frac(double x)
{
    static const double c[8] = {2.8853900817779268114E+0, 2.8312651192953993354E-2,
        5.0006798065881969549E-4, 1.0514733588011180538E-5, 2.4074128088151586443E-7,
        5.7988453014506741861E-9, 1.4374842194796670219E-10, 4.0928048937567843469E-12};
    int pred = x > SQRT_2;
    x *= pred ? 0.5 : ONE_OVER_SQ2;
    double z = (x - 1.) * orc_rcp(x + 1.);
    double y = z * z * 33.9705627484771406;
    double s = c[7];
    for (int i = 6; i >= 0; i--) s = fma(s, y, c[i]);
    return fma(z, s, (pred ? 1.0 : 0.5));
}

/* UM/math_meso.h:308-324 */
double orc_exp2d_frac(double x)
{
    static const double c[12] = {9.9999999999999999572E-1, 6.9314718055994653980E-1,
        2.4022650695904222220E-1, 5.5504108665909870679E-2, 9.6181290971755593396E-3,
        1.3333558738165095559E-3, 1.5403509189194102748E-4, 1.5253232908458899497E-5,
        1.3207676270599404858E-6, 1.0258347084283025531E-7, 6.5379419072372670333E-9,
        6.3026908837748924689E-10};
    double s = c[11];
    for (int i = 10; i >= 0; i--) s = fma(s, x, c[i]);
    return s;
}

/* UM/math_meso.h:332-345 */
double orc_powd(double a, double b)
{
    int64_t bits = d2ll(a);
    int32_t hi = (int32_t)(bits >> 32), lo = (int32_t)(bits & 0xFFFFFFFF);
    double I = (hi >> 20) - 1023;
    double F = orc_log2d_frac(hilo2d((hi & 0x000FFFFF) | 0x3FF00000, lo));
    double II = floor(b * (I + F));
    return two_to_n((int)II) * orc_exp2d_frac(fma(b, F, fma(b, I, -II)));
}

/* UM/math_meso.h:357-370 */
double orc_sinpi(double x)
{
    static const double c[7] = {9.99999999999249900E-1, -1.23370055006260170E+0,
        2.53669506722714547E-1, -2.08634736828917670E-2, 9.19240000031795430E-4,
        -2.51721958850906157E-5, 4.49220128554338954E-7};
    x = 2.0 * x - 1.0;
    x *= x;
    double s = c[6];
    for (int i = 5; i >= 0; i--) s = fma(s, x, c[i]);
    return s;
}

/* UM/math_meso.h:380-392 */
double orc_cospi(double x)
{
    static const double c[6] = {-1.57079632662144460E+0, 6.45964092644060746E-1,
        -7.96925872866600517E-2, 4.68162024021793872E-3, -1.60217135750921262E-4,
        3.41817283473266926E-6};
    x = 2.0 * x - 1.0;
    double x2 = x * x;
    double s = c[5];
    for (int i = 4; i >= 0; i--) s = fma(s, x2, c[i]);
    return s * x;
}

/* UM/math_meso.h:401-424 */
static double log2u_frac(double x, double e)
{
    static const double c[5] = {2.88539008179006374E+0, 2.83126505877817866E-2,
        5.00072802051539862E-4, 1.05013262724846015E-5, 2.55854634203511155E-7};
    int pred = x > SQRT_2;
    x *= pred ? 0.5 : ONE_OVER_SQ2;
    double z = (x - 1.) * orc_rcp(x + 1.);
    double y = z * z * 33.9705627484771406;
    double s = c[4];
    for (int i = 3; i >= 0; i--) s = fma(s, y, c[i]);
    return fma(z, s, (pred ? 1.0 : 0.5) + e);
}
double orc_log2u(uint32_t x)
{
    int I = 31 - __builtin_clz(x);
    return log2u_frac((double)x * two_to_n(-I), (double)(I - 32));
}

/* ====================================================================== */
/* per-pair Gaussian, UM/math_meso.h:466-484; call sites                  */
/* UM/pair_dpd_meso.cu:145, UM/pair_dpd_fast_meso.cu:145                  */
/* ====================================================================== */

double orc_gaussian_dp(uint32_t si, uint32_t sj)
{
    int pred = si > sj;                       /* unsigned compare (f3u::i is uint) */
    uint32_t v0 = pred ? si : sj, v1 = !pred ? si : sj;
    orc_tea(&v0, &v1, 4);
    double f = orc_cospi((v0 & 0x7FFFFFFFu) * 4.6566128730773925781E-10) *
               ((v0 & 0x80000000u) ? 1.0 : -1.0);
    uint32_t m = v1 > 1u ? v1 : 1u;
    double r = orc_sqrtd(-2.0 * LN_2 * orc_log2u(m));
    double x = r * f;
    x = x < 4.0 ? x : 4.0;                    /* bound(): max(lower, min(x, upper)) */
    return x > -4.0 ? x : -4.0;
}

float orc_gaussian_sp(uint32_t si, uint32_t sj)
{
    int pred = si > sj;
    uint32_t v0 = pred ? si : sj, v1 = !pred ? si : sj;
    orc_tea(&v0, &v1, 4);
    float a = (float)(int32_t)v0 * (float)4.6566128730773925781E-10; /* int(v0)*2^-31 */
    float f = (float)sin(M_PI * (double)a);                          /* sinpif */
    float u = (float)v1 * (float)2.3283064365386962891E-10;          /* v1*2^-32 */
    float r = sqrtf(-2.0f * (float)LN_2 * log2f(u));
    float x = r * f;
    /* CUDA fminf/fmaxf return the non-NaN operand: NaN -> 4 -> 4 */
    x = fminf(x, 4.0f);
    x = fmaxf(-4.0f, x);
    return x;
}

/* ====================================================================== */
/* world                                                                   */
/* ====================================================================== */

#define NCOEFF 7
enum { P_CUT, P_CUTSQ, P_CUTINV, P_EXPW, P_A0, P_GAMMA, P_SIGMA }; /* UM/pair_dpd_meso.h:15-24 */

typedef struct {
    int dim, dir;            /* dir 0: send to lower neighbor, 1: to upper */
    double lo, hi;           /* slab */
    int pbc_flag, pbc[3];
    int sendproc, recvproc, sendflag;
    int *sendlist, nsend, maxsend;
    int nrecv, firstrecv;
} orc_swap;

typedef struct {
    int id, loc[3];
    double sublo[3], subhi[3];
    int procneigh[3][2];
    int sendneed[3][2], recvneed[3][2];
    int nswap; orc_swap swap[6];

    int nlocal, nghost, nmax, n_bulk, n_border;
    double *x, *v, *f; int *tag, *type, *mask, *image;
    double *virial, *e_pair, *e_bond;
    float *coord4, *veloc4;

    int m[3]; double binsize[3], bininv[3];
    int bins_ready;
    double expected_neigh_count;
    int ncell, *cell_start, *cell_atoms, *cell_id;
    int n_col, *pair_count, *pair_rows, list_max;
    uint64_t *key; int *perm_from; int key_max;

    /* exchange scratch */
    double *buf_l, *buf_r; int nbuf_l, nbuf_r, maxbuf_l, maxbuf_r;
} orc_rank;

struct orc_world {
    int nranks, procgrid[3];
    double boxlo[3], boxhi[3], prd[3]; int periodic[3];
    double cut_max, skin, cutneighmax, cutghost;
    int maxneed[3];
    int ntypes; double *mass, *coeff;
    uint32_t seed; double dt; int every, ago; long ntimestep;
    int precision;
    orc_rank *rk;
    /* bead-spring topology, keyed by TAG (a bond is a property of the atom, wherever the rank arrays keep it) */
    int bond_per_atom, tag_max, nbondtypes; double special_lj12;
    int *num_bond, *bond_type, *bond_atom;      /* [tag_max+1], [tag_max+1][bond_per_atom] */
    double *bond_k, *bond_r0;                   /* [nbondtypes+1] */
    /* channel fixes, in registration order */
    int nfix; struct { int kind, groupbit, dims; double p[4]; unsigned long long *hist; long samples; } fix[8];
    int integrate_groupbit;                     /* group of the deck's fix nve/meso (0 = unset -> 1) */
};

static int rank_of(const orc_world *w, int ix, int iy, int iz)
{
    return (ix * w->procgrid[1] + iy) * w->procgrid[2] + iz;
}

static void rank_grow(orc_rank *r, int n)
{
    if (n <= r->nmax) return;
    int nm = n + n / 2 + 1024;
    r->x = realloc(r->x, sizeof(double) * 3 * nm);
    r->v = realloc(r->v, sizeof(double) * 3 * nm);
    r->f = realloc(r->f, sizeof(double) * 3 * nm);
    r->tag = realloc(r->tag, sizeof(int) * nm);
    r->type = realloc(r->type, sizeof(int) * nm);
    r->mask = realloc(r->mask, sizeof(int) * nm);
    r->image = realloc(r->image, sizeof(int) * nm);
    r->virial = realloc(r->virial, sizeof(double) * 6 * nm);
    r->e_pair = realloc(r->e_pair, sizeof(double) * nm);
    r->e_bond = realloc(r->e_bond, sizeof(double) * nm);
    r->coord4 = realloc(r->coord4, sizeof(float) * 4 * nm);
    r->veloc4 = realloc(r->veloc4, sizeof(float) * 4 * nm);
    r->cell_id = realloc(r->cell_id, sizeof(int) * nm);
    r->cell_atoms = realloc(r->cell_atoms, sizeof(int) * nm);
    memset(r->f + 3 * r->nmax, 0, sizeof(double) * 3 * (nm - r->nmax));
    r->nmax = nm;
}

/* Comm::setup, comm.cpp:393-640 (orthogonal, style SINGLE, uniform) */
static int comm_setup(orc_world *w)
{
    w->cutghost = w->cutneighmax;   /* comm.cpp:408,414 */
    for (int d = 0; d < 3; d++) {
        w->maxneed[d] = (int)(w->cutghost * w->procgrid[d] / w->prd[d]) + 1; /* comm.cpp:470-472 */
        if (!w->periodic[d] && w->maxneed[d] > w->procgrid[d] - 1) w->maxneed[d] = w->procgrid[d] - 1;
        if (w->maxneed[d] > 1)
            FAIL("sub-domain thinner than ghost cutoff in dim %d (maxneed %d): unsupported", d, w->maxneed[d]);
    }
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        for (int d = 0; d < 3; d++) {
            int mn = w->maxneed[d], p = w->procgrid[d], me = r->loc[d];
            if (!w->periodic[d]) {                       /* comm.cpp:478-488 */
                r->recvneed[d][0] = mn < me ? mn : me;
                r->recvneed[d][1] = mn < p - me - 1 ? mn : p - me - 1;
                int left = me - 1; if (left < 0) left = p - 1;
                r->sendneed[d][0] = mn < p - left - 1 ? mn : p - left - 1;
                int right = me + 1; if (right == p) right = 0;
                r->sendneed[d][1] = mn < right ? mn : right;
            } else {
                r->recvneed[d][0] = r->recvneed[d][1] = r->sendneed[d][0] = r->sendneed[d][1] = mn;
            }
        }
        r->nswap = 0;
        for (int d = 0; d < 3; d++)
            for (int ineed = 0; ineed < 2 * w->maxneed[d]; ineed++) {   /* comm.cpp:583-637 */
                orc_swap *s = &r->swap[r->nswap++];
                int *keep_list = s->sendlist; int keep_max = s->maxsend;
                memset(s, 0, sizeof *s);
                s->sendlist = keep_list; s->maxsend = keep_max;
                s->dim = d; s->dir = ineed % 2;
                if (ineed % 2 == 0) {
                    s->sendproc = r->procneigh[d][0]; s->recvproc = r->procneigh[d][1];
                    s->lo = -1.0e20; s->hi = r->sublo[d] + w->cutghost;
                    if (r->loc[d] == 0) { s->pbc_flag = 1; s->pbc[d] = 1; }
                } else {
                    s->sendproc = r->procneigh[d][1]; s->recvproc = r->procneigh[d][0];
                    s->lo = r->subhi[d] - w->cutghost; s->hi = 1.0e20;
                    if (r->loc[d] == w->procgrid[d] - 1) { s->pbc_flag = 1; s->pbc[d] = -1; }
                }
                /* UM/comm_meso.cu:85-86 */
                s->sendflag = (ineed / 2 >= r->sendneed[d][ineed % 2]) ? 0 : 1;
            }
    }
    return 0;
}

orc_world *orc_world_create(const double boxlo[3], const double boxhi[3],
                            const int periodic[3], const int procgrid[3],
                            int ntypes, const double *mass, const double *coeff7,
                            double cut_max, double skin, int every,
                            uint32_t seed, double dt, int precision)
{
    orc_world *w = calloc(1, sizeof *w);
    w->special_lj12 = 1.0;
    for (int d = 0; d < 3; d++) {
        w->boxlo[d] = boxlo[d]; w->boxhi[d] = boxhi[d]; w->prd[d] = boxhi[d] - boxlo[d];
        w->periodic[d] = periodic[d]; w->procgrid[d] = procgrid[d];
    }
    w->nranks = procgrid[0] * procgrid[1] * procgrid[2];
    w->ntypes = ntypes;
    w->mass = malloc(sizeof(double) * (ntypes + 1));
    memcpy(w->mass, mass, sizeof(double) * (ntypes + 1));
    w->coeff = malloc(sizeof(double) * ntypes * ntypes * NCOEFF);
    memcpy(w->coeff, coeff7, sizeof(double) * ntypes * ntypes * NCOEFF);
    w->cut_max = cut_max; w->skin = skin; w->cutneighmax = cut_max + skin; /* neighbor.cpp: cutneighmax = cutforce + skin */
    w->every = every; w->seed = seed; w->dt = dt; w->precision = precision;
    w->rk = calloc(w->nranks, sizeof(orc_rank));
    for (int ix = 0; ix < procgrid[0]; ix++)
        for (int iy = 0; iy < procgrid[1]; iy++)
            for (int iz = 0; iz < procgrid[2]; iz++) {
                orc_rank *r = &w->rk[rank_of(w, ix, iy, iz)];
                r->id = rank_of(w, ix, iy, iz);
                r->loc[0] = ix; r->loc[1] = iy; r->loc[2] = iz;
                for (int d = 0; d < 3; d++) {
                    /* Domain::set_local_box, domain.cpp: sublo = boxlo + prd*split[loc] */
                    int p = procgrid[d], me = r->loc[d];
                    double inv = 1.0 / p;
                    r->sublo[d] = w->boxlo[d] + w->prd[d] * (me * inv);
                    r->subhi[d] = (me < p - 1) ? w->boxlo[d] + w->prd[d] * ((me + 1) * inv) : w->boxhi[d];
                    int l[3] = {ix, iy, iz}, u[3] = {ix, iy, iz};
                    l[d] = (me - 1 + p) % p; u[d] = (me + 1) % p;
                    r->procneigh[d][0] = rank_of(w, l[0], l[1], l[2]);
                    r->procneigh[d][1] = rank_of(w, u[0], u[1], u[2]);
                }
            }
    if (comm_setup(w)) { orc_world_destroy(w); return NULL; }
    return w;
}

void orc_world_destroy(orc_world *w)
{
    if (!w) return;
    for (int i = 0; i < w->nranks; i++) {
        orc_rank *r = &w->rk[i];
        free(r->x); free(r->v); free(r->f); free(r->tag); free(r->type); free(r->mask);
        free(r->image); free(r->virial); free(r->e_pair); free(r->coord4); free(r->veloc4);
        free(r->cell_start); free(r->cell_atoms); free(r->cell_id);
        free(r->pair_count); free(r->pair_rows); free(r->key); free(r->perm_from);
        free(r->buf_l); free(r->buf_r);
        for (int s = 0; s < 6; s++) free(r->swap[s].sendlist);
    }
    free(w->num_bond); free(w->bond_type); free(w->bond_atom); free(w->bond_k); free(w->bond_r0);
    for (int k = 0; k < w->nfix; k++) free(w->fix[k].hist);
    free(w->rk); free(w->mass); free(w->coeff); free(w);
}

int orc_nranks(orc_world *w) { return w->nranks; }
void orc_set_timestep(orc_world *w, long t) { w->ntimestep = t; }
long orc_get_timestep(orc_world *w) { return w->ntimestep; }

#define IMGMASK 1023
#define IMGBITS 10
#define IMG2BITS 20

/* Domain::pbc as restated by MesoDomain::pbc, UM/domain_meso.cu:30-145 (orthogonal, no deform) */
static void pbc_one(const orc_world *w, double *x, int *image)
{
    for (int d = 0; d < 3; d++) {
        if (!w->periodic[d]) continue;
        int sh = d * IMGBITS;
        if (x[d] < w->boxlo[d]) {
            x[d] += w->prd[d];
            int idim = (*image >> sh) & IMGMASK; int other = *image ^ (idim << sh);
            idim = (idim - 1) & IMGMASK; *image = other | (idim << sh);
        }
        if (x[d] >= w->boxhi[d]) {
            x[d] -= w->prd[d];
            if (x[d] < w->boxlo[d]) x[d] = w->boxlo[d];
            int idim = (*image >> sh) & IMGMASK; int other = *image ^ (idim << sh);
            idim = (idim + 1) & IMGMASK; *image = other | (idim << sh);
        }
    }
}

int orc_world_set_atoms(orc_world *w, int n, const double *x, const double *v,
                        const int *tag, const int *type, const int *mask, const int *image)
{
    for (int i = 0; i < w->nranks; i++) { w->rk[i].nlocal = 0; w->rk[i].nghost = 0; }
    for (int i = 0; i < n; i++) {
        double xi[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        int img = image ? image[i] : ((512 << IMG2BITS) | (512 << IMGBITS) | 512);
        pbc_one(w, xi, &img);       /* read_data remaps into the box (Domain::remap) */
        int loc[3];
        for (int d = 0; d < 3; d++) {
            int p = w->procgrid[d]; loc[d] = -1;
            for (int k = 0; k < p; k++) {
                double inv = 1.0 / p;
                double lo = w->boxlo[d] + w->prd[d] * (k * inv);
                double hi = (k < p - 1) ? w->boxlo[d] + w->prd[d] * ((k + 1) * inv) : w->boxhi[d];
                if (xi[d] >= lo && xi[d] < hi) { loc[d] = k; break; }
            }
            if (loc[d] < 0) FAIL("atom %d outside the non-periodic box", i);
        }
        orc_rank *r = &w->rk[rank_of(w, loc[0], loc[1], loc[2])];
        rank_grow(r, r->nlocal + 1);
        int k = r->nlocal++;
        for (int d = 0; d < 3; d++) { r->x[3 * k + d] = xi[d]; r->v[3 * k + d] = v ? v[3 * i + d] : 0.0; r->f[3 * k + d] = 0.0; }
        r->tag[k] = tag ? tag[i] : i + 1;
        r->type[k] = type ? type[i] : 1;
        r->mask[k] = mask ? mask[i] : 1;
        r->image[k] = img;
    }
    return 0;
}

static void copy_atom(orc_rank *r, int from, int to)   /* AtomVecAtomic::copy */
{
    if (from == to) return;
    memcpy(r->x + 3 * to, r->x + 3 * from, 24);
    memcpy(r->v + 3 * to, r->v + 3 * from, 24);
    r->tag[to] = r->tag[from]; r->type[to] = r->type[from];
    r->mask[to] = r->mask[from]; r->image[to] = r->image[from];
}

/* ---------------------------------------------------------------------- */
/* MesoComm::exchange, UM/comm_meso.cu:256-420 (the "#if 1" branch)        */
/* record = {x[3], v[3], tag, type, mask, image} as 10 doubles             */
/* ---------------------------------------------------------------------- */
#define XREC 10
static void buf_push(double **buf, int *n, int *max, const orc_rank *r, int i)
{
    if (*n + XREC > *max) { *max = *max * 2 + 1024 * XREC; *buf = realloc(*buf, sizeof(double) * *max); }
    double *b = *buf + *n;
    for (int d = 0; d < 3; d++) { b[d] = r->x[3 * i + d]; b[3 + d] = r->v[3 * i + d]; }
    b[6] = r->tag[i]; b[7] = r->type[i]; b[8] = r->mask[i]; b[9] = r->image[i];
    *n += XREC;
}
static void unpack_exchange(orc_rank *r, const double *b)
{
    rank_grow(r, r->nlocal + 1);
    int k = r->nlocal++;
    for (int d = 0; d < 3; d++) { r->x[3 * k + d] = b[d]; r->v[3 * k + d] = b[3 + d]; }
    r->tag[k] = (int)b[6]; r->type[k] = (int)b[7]; r->mask[k] = (int)b[8]; r->image[k] = (int)b[9];
}
static void recv_exchange(orc_rank *r, int dim, const double *buf, int n)
{
    double lo = r->sublo[dim], hi = r->subhi[dim];
    for (int m = 0; m < n; m += XREC) {
        double value = buf[m + dim];
        if (value >= lo && value < hi) unpack_exchange(r, buf + m);
        /* else: the reference prints "rejected" and drops the atom */
    }
}

static void exchange(orc_world *w)
{
    for (int i = 0; i < w->nranks; i++) w->rk[i].nghost = 0;
    for (int dim = 0; dim < 3; dim++) {
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *r = &w->rk[ir];
            double lo = r->sublo[dim], hi = r->subhi[dim], mid = 0.5 * (lo + hi);
            r->nbuf_l = r->nbuf_r = 0;
            int i = 0, nlocal = r->nlocal;
            while (i < nlocal) {
                double xd = r->x[3 * i + dim];
                if (xd >= hi || xd < lo) {
                    double dist = xd - mid;
                    if (w->periodic[dim]) {            /* Domain::minimum_image, orthogonal */
                        if (fabs(dist) > 0.5 * w->prd[dim]) { if (dist < 0.0) dist += w->prd[dim]; else dist -= w->prd[dim]; }
                    }
                    if (dist < 0) buf_push(&r->buf_l, &r->nbuf_l, &r->maxbuf_l, r, i);
                    else buf_push(&r->buf_r, &r->nbuf_r, &r->maxbuf_r, r, i);
                    copy_atom(r, nlocal - 1, i);
                    nlocal--;
                } else i++;
            }
            r->nlocal = nlocal;
        }
        int p = w->procgrid[dim];
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *r = &w->rk[ir];
            if (p == 1) {
                /* copies of the buffers: recv may append to r */
                int nl = r->nbuf_l, nr = r->nbuf_r;
                double *tl = malloc(sizeof(double) * (nl + 1)), *tr = malloc(sizeof(double) * (nr + 1));
                memcpy(tl, r->buf_l, sizeof(double) * nl); memcpy(tr, r->buf_r, sizeof(double) * nr);
                recv_exchange(r, dim, tl, nl); recv_exchange(r, dim, tr, nr);
                free(tl); free(tr);
            } else if (p == 2) {
                orc_rank *o = &w->rk[r->procneigh[dim][1]];
                recv_exchange(r, dim, o->buf_l, o->nbuf_l);
                recv_exchange(r, dim, o->buf_r, o->nbuf_r);
            } else {
                orc_rank *right = &w->rk[r->procneigh[dim][1]], *left = &w->rk[r->procneigh[dim][0]];
                recv_exchange(r, dim, right->buf_l, right->nbuf_l);
                recv_exchange(r, dim, left->buf_r, left->nbuf_r);
            }
        }
    }
}

/* ---------------------------------------------------------------------- */
/* cell lattice: MesoNeighbor::setup_bins, UM/neighbor_meso.cu:858-931     */
/* ---------------------------------------------------------------------- */
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }
static int ceiling(int x, int inc) { return ((x + inc - 1) / inc) * inc; } /* UM/math_meso.h:33-36 */

static void setup_bins(orc_world *w, orc_rank *r)
{
    double dim[3]; for (int d = 0; d < 3; d++) dim[d] = r->subhi[d] - r->sublo[d];
    double vol = dim[0] * dim[1] * dim[2];
    double dens = r->nlocal / vol;
    if (dens < 3) dens = 3;
    double enc = dens * (4.0 / 3.0 * 3.142 * pow(w->cutneighmax, 3.0));
    enc *= 4.0;
    if (enc < 32.0) enc = 32.0;
    r->expected_neigh_count = enc;
    double binsize_optimal = 1.0 * w->cutneighmax, inv = 1.0 / binsize_optimal;
    for (int d = 0; d < 3; d++) {
        r->m[d] = imax((int)(dim[d] * inv), 1) + 2;
        r->binsize[d] = dim[d] / (r->m[d] - 2);
        r->bininv[d] = 1.0 / r->binsize[d];
    }
    r->ncell = r->m[0] * r->m[1] * r->m[2];
    r->cell_start = realloc(r->cell_start, sizeof(int) * (r->ncell + 1));
    r->bins_ready = 1;
}

/* CUDA cvt.rzi.s32.f64 saturates; C's cast is UB out of range */
static int d2i(double v)
{
    if (!(v == v)) return 0;
    if (v >= 2147483647.0) return INT_MAX;
    if (v <= -2147483648.0) return INT_MIN;
    return (int)v;
}
/* UM/math_meso.h:155-158 : clamp at [nmin, nmax) */
static int clampi(int i, int nmin, int nmax) { return imax(nmin, imin(i, nmax - 1)); }

/* UM/comm_meso.cu:188-254 : number of send slabs containing local atom i */
static int borderness(const orc_rank *r, int i)
{
    int b = 0;
    for (int s = 0; s < r->nswap; s++) {
        const orc_swap *sw = &r->swap[s];
        if (!sw->sendflag) continue;
        double xd = r->x[3 * i + sw->dim];
        if (xd >= sw->lo && xd <= sw->hi) b++;      /* UM/comm_meso.h:71-74 */
    }
    return b;
}

/* stable LSD radix sort of (key,val), 16 bits per pass */
static void radix_sort_u64(uint64_t *key, int *val, int n, int bits)
{
    uint64_t *k2 = malloc(sizeof(uint64_t) * (n + 1)); int *v2 = malloc(sizeof(int) * (n + 1));
    int *cnt = malloc(sizeof(int) * 65537);
    for (int sh = 0; sh < bits; sh += 16) {
        memset(cnt, 0, sizeof(int) * 65537);
        for (int i = 0; i < n; i++) cnt[((key[i] >> sh) & 0xFFFF) + 1]++;
        for (int i = 0; i < 65536; i++) cnt[i + 1] += cnt[i];
        for (int i = 0; i < n; i++) { int p = cnt[(key[i] >> sh) & 0xFFFF]++; k2[p] = key[i]; v2[p] = val[i]; }
        memcpy(key, k2, sizeof(uint64_t) * n); memcpy(val, v2, sizeof(int) * n);
    }
    free(k2); free(v2); free(cnt);
}

/* MesoAtom::sort_local, UM/atom_meso.cu:343-384; key UM/atom_meso.cu:288-306 */
static void sort_local(orc_world *w, orc_rank *r)
{
    (void)w;
    int n = r->nlocal;
    if (n > r->key_max) {
        r->key_max = n + n / 2 + 16;
        r->key = realloc(r->key, sizeof(uint64_t) * r->key_max);
        r->perm_from = realloc(r->perm_from, sizeof(int) * r->key_max);
    }
    const int l2_resoln = 16;
    int max_bin = imax(imax(r->m[0], r->m[1]), r->m[2]);
    int l1_width = 3 * (int)floor(log2(max_bin * 2.0));
    int l2_width = 3 * (int)log2((double)l2_resoln);
    int l0_shift = l2_width + l1_width, l1_shift = l2_width;
    uint64_t border_mask = 1ULL << l0_shift;
    int nb = 0;
    for (int i = 0; i < n; i++) {
        uint32_t b[3], s[3];
        for (int d = 0; d < 3; d++) {
            double xd = r->x[3 * i + d];
            /* a*b+c and a-b*c are spelled as the FMAs nvcc's default -fmad=true contracts them to */
            b[d] = (uint32_t)clampi(d2i(fma(xd - r->sublo[d], r->bininv[d], 1.0)), 0, r->m[d]);
            double coord_inv = l2_resoln * r->bininv[d];
            /* (bin_id - 1) is unsigned arithmetic in the reference */
            s[d] = (uint32_t)clampi(d2i(fma(-(double)(uint32_t)(b[d] - 1u), r->binsize[d], xd) * coord_inv), 0, l2_resoln);
        }
        uint64_t z1 = orc_interleave3(b[0], b[1], b[2]), z2 = orc_interleave3(s[0], s[1], s[2]);
        uint64_t k = (z1 << l1_shift) | z2;
        if (borderness(r, i)) { k |= border_mask; nb++; }
        r->key[i] = k; r->perm_from[i] = i;
    }
    r->n_border = nb; r->n_bulk = n - nb;          /* UM/atom_meso.cu:316-341 */
    radix_sort_u64(r->key, r->perm_from, n, 1 + l1_width + l2_width);
    /* transfer_post_sort: gather ESSENTIAL attributes into sorted order */
    double *x2 = malloc(sizeof(double) * 3 * (n + 1)), *v2 = malloc(sizeof(double) * 3 * (n + 1));
    int *t2 = malloc(sizeof(int) * 4 * (n + 1));
    for (int p = 0; p < n; p++) {
        int o = r->perm_from[p];
        memcpy(x2 + 3 * p, r->x + 3 * o, 24); memcpy(v2 + 3 * p, r->v + 3 * o, 24);
        t2[4 * p] = r->tag[o]; t2[4 * p + 1] = r->type[o]; t2[4 * p + 2] = r->mask[o]; t2[4 * p + 3] = r->image[o];
    }
    memcpy(r->x, x2, sizeof(double) * 3 * n); memcpy(r->v, v2, sizeof(double) * 3 * n);
    for (int p = 0; p < n; p++) { r->tag[p] = t2[4 * p]; r->type[p] = t2[4 * p + 1]; r->mask[p] = t2[4 * p + 2]; r->image[p] = t2[4 * p + 3]; }
    free(x2); free(v2); free(t2);
}

/* ---------------------------------------------------------------------- */
/* MesoComm::borders, UM/comm_meso.cu:41-186                               */
/* ---------------------------------------------------------------------- */
static void borders(orc_world *w)
{
    int maxswap = 0;
    for (int ir = 0; ir < w->nranks; ir++) maxswap = imax(maxswap, w->rk[ir].nswap);
    /* nlast bookkeeping per rank */
    int *nfirst = calloc(w->nranks, sizeof(int)), *nlast = calloc(w->nranks, sizeof(int));
    for (int s = 0; s < maxswap; s++) {
        /* all ranks select + "pack" (nswap identical on all ranks: maxneed is global) */
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *r = &w->rk[ir]; orc_swap *sw = &r->swap[s];
            if (sw->dir == 0) {            /* ineed % 2 == 0: first swap of this dim */
                nfirst[ir] = r->n_bulk;    /* nlast = n_bulk at dim start, then nfirst = nlast */
                nlast[ir] = r->nlocal + r->nghost;
            }
            sw->nsend = 0;
            if (sw->sendflag) {
                for (int i = nfirst[ir]; i < nlast[ir]; i++) {
                    double xd = r->x[3 * i + sw->dim];
                    if (xd >= sw->lo && xd <= sw->hi) {
                        if (sw->nsend >= sw->maxsend) { sw->maxsend = sw->maxsend * 2 + 1024; sw->sendlist = realloc(sw->sendlist, sizeof(int) * sw->maxsend); }
                        sw->sendlist[sw->nsend++] = i;
                    }
                }
            }
        }
        /* all ranks receive from recvproc: the sender is the rank whose sendproc == me,
           which on a periodic ring is recvproc */
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *r = &w->rk[ir]; orc_swap *sw = &r->swap[s];
            orc_rank *o = &w->rk[sw->recvproc]; orc_swap *so = &o->swap[s];
            int nrecv = so->nsend;
            sw->nrecv = nrecv; sw->firstrecv = r->nlocal + r->nghost;
            rank_grow(r, r->nlocal + r->nghost + nrecv);
            /* rank_grow may move r's arrays, but o's send indices stay valid (o may be r) */
            for (int k = 0; k < nrecv; k++) {
                int j = so->sendlist[k], i = sw->firstrecv + k;
                for (int d = 0; d < 3; d++) {
                    /* pack_border_vel, UM/atom_vec_dpd_atomic_meso.cu:61-100 */
                    double sh = so->pbc_flag ? so->pbc[d] * w->prd[d] : 0.0;
                    r->x[3 * i + d] = so->pbc_flag ? o->x[3 * j + d] + sh : o->x[3 * j + d];
                    r->v[3 * i + d] = o->v[3 * j + d];
                }
                r->type[i] = o->type[j]; r->mask[i] = o->mask[j]; r->tag[i] = o->tag[j];
                r->image[i] = 0;
            }
        }
        for (int ir = 0; ir < w->nranks; ir++) w->rk[ir].nghost += w->rk[ir].swap[s].nrecv;
    }
    free(nfirst); free(nlast);
}

/* Comm::forward_comm, comm.cpp:686-753 with pack/unpack_comm_vel */
void orc_forward_comm(orc_world *w)
{
    int maxswap = 0;
    for (int ir = 0; ir < w->nranks; ir++) maxswap = imax(maxswap, w->rk[ir].nswap);
    for (int s = 0; s < maxswap; s++) {
        /* two-phase so that a rank never reads values written in this same swap */
        double **tmp = malloc(sizeof(double *) * w->nranks);
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *o = &w->rk[ir]; orc_swap *so = &o->swap[s];
            tmp[ir] = malloc(sizeof(double) * 6 * (so->nsend + 1));
            for (int k = 0; k < so->nsend; k++) {
                int j = so->sendlist[k];
                for (int d = 0; d < 3; d++) {
                    double sh = so->pbc_flag ? so->pbc[d] * w->prd[d] : 0.0;
                    tmp[ir][6 * k + d] = so->pbc_flag ? o->x[3 * j + d] + sh : o->x[3 * j + d];
                    tmp[ir][6 * k + 3 + d] = o->v[3 * j + d];
                }
            }
        }
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *r = &w->rk[ir]; orc_swap *sw = &r->swap[s];
            const double *b = tmp[sw->recvproc];
            for (int k = 0; k < sw->nrecv; k++) {
                int i = sw->firstrecv + k;
                for (int d = 0; d < 3; d++) { r->x[3 * i + d] = b[6 * k + d]; r->v[3 * i + d] = b[6 * k + 3 + d]; }
            }
        }
        for (int ir = 0; ir < w->nranks; ir++) free(tmp[ir]);
        free(tmp);
    }
}

/* ---------------------------------------------------------------------- */
/* pack + signature: gpu_merge_xvt, UM/atom_vec_meso.cu:142-192            */
/* ---------------------------------------------------------------------- */
static void pack_rank(orc_rank *r, uint32_t seed_now, int beg, int end)
{
    double c[3]; for (int d = 0; d < 3; d++) c[d] = 0.5 * (r->subhi[d] + r->sublo[d]);
    for (int i = beg; i < end; i++) {
        for (int d = 0; d < 3; d++) r->coord4[4 * i + d] = (float)(r->x[3 * i + d] - c[d]);
        r->coord4[4 * i + 3] = u2f((uint32_t)(r->type[i] - 1));
        float vx = (float)r->v[3 * i], vy = (float)r->v[3 * i + 1], vz = (float)r->v[3 * i + 2];
        r->veloc4[4 * i] = vx; r->veloc4[4 * i + 1] = vy; r->veloc4[4 * i + 2] = vz;
        r->veloc4[4 * i + 3] = u2f(orc_signature(seed_now, r->tag[i], vx, vy, vz));
    }
}
void orc_pack(orc_world *w, uint32_t seed_now)
{
    for (int ir = 0; ir < w->nranks; ir++) pack_rank(&w->rk[ir], seed_now, 0, w->rk[ir].nlocal + w->rk[ir].nghost);
}

/* ---------------------------------------------------------------------- */
/* binning: UM/neighbor_meso.cu:386-475,535-600                            */
/* ---------------------------------------------------------------------- */
static void binning(orc_rank *r)
{
    int nall = r->nlocal + r->nghost;
    int *cnt = r->cell_start;
    memset(cnt, 0, sizeof(int) * (r->ncell + 1));
    for (int i = 0; i < nall; i++) {
        int b[3];
        for (int d = 0; d < 3; d++) {
            double xd = r->x[3 * i + d];
            b[d] = clampi(d2i(fma(xd - r->sublo[d], r->bininv[d], 1.0)), 0, r->m[d]);
            if (i >= r->nlocal)                       /* ghosts clamp to the outer layer */
                b[d] = (xd >= r->sublo[d]) ? (xd <= r->subhi[d] ? b[d] : r->m[d] - 1) : 0;
        }
        int c = b[0] + r->m[0] * (b[1] + b[2] * r->m[1]);
        r->cell_id[i] = c; cnt[c + 1]++;
    }
    for (int c = 0; c < r->ncell; c++) cnt[c + 1] += cnt[c];
    int *fill = malloc(sizeof(int) * (r->ncell + 1));
    memcpy(fill, cnt, sizeof(int) * (r->ncell + 1));
    for (int i = 0; i < nall; i++) r->cell_atoms[fill[r->cell_id[i]]++] = i;   /* stable: ascending i in a cell */
    free(fill);
}

/* stencil order: gpu_stencil_full_bin_3d, UM/neighbor_meso.cu:772-825 */
static int stencil_of(const orc_rank *r, int cell, int *out)
{
    int bx = cell % r->m[0], by = (cell / r->m[0]) % r->m[1], bz = cell / (r->m[0] * r->m[1]);
    uint32_t key[27]; int n = 0;
    for (int k = -1; k <= 1; k++) for (int j = -1; j <= 1; j++) for (int i = -1; i <= 1; i++) {
        int x = bx + i, y = by + j, z = bz + k;
        if (x < 0 || x >= r->m[0] || y < 0 || y >= r->m[1] || z < 0 || z >= r->m[2]) continue;
        out[n] = x + r->m[0] * (y + z * r->m[1]);
        key[n] = orc_interleave3((uint32_t)x, (uint32_t)y, (uint32_t)z);
        if (x == 0 || x == r->m[0] - 1 || y == 0 || y == r->m[1] - 1 || z == 0 || z == r->m[2] - 1) key[n] += 0x80000000u;
        n++;
    }
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++)
        if (key[i] > key[j]) { int t = out[i]; out[i] = out[j]; out[j] = t; uint32_t u = key[i]; key[i] = key[j]; key[j] = u; }
    return n;
}
int orc_get_stencil(orc_world *w, int ir, int cell, int *out27) { return stencil_of(&w->rk[ir], cell, out27); }

/* neighbor list: UM/neigh_build_meso.cu:20-119 (build), 166-200 (join) */
static int neighbor_build(orc_world *w, orc_rank *r)
{
    binning(r);
    pack_rank(r, 0, 0, r->nlocal + r->nghost);          /* UM/neigh_build_meso.cu:266 */
    r->n_col = ceiling((int)r->expected_neigh_count, 32); /* UM/neigh_list_meso.cu:38-40 */
    if (r->nlocal > r->list_max) {
        r->list_max = r->nlocal + r->nlocal / 2 + 32;
        r->pair_count = realloc(r->pair_count, sizeof(int) * r->list_max);
        r->pair_rows = realloc(r->pair_rows, sizeof(int) * (size_t)r->list_max * r->n_col);
    }
    float rc2_core = (float)pow(w->cutneighmax - w->skin, 2.0);
    float rc2_tail = (float)pow(w->cutneighmax, 2.0);
    int *skin = malloc(sizeof(int) * r->n_col * 4);
    for (int c = 0; c < r->ncell; c++) {
        int beg = r->cell_start[c], end = r->cell_start[c + 1];
        if (beg == end) continue;
        if (r->cell_atoms[beg] >= r->nlocal) continue;    /* bin_is_ghost: flag of first atom */
        int st[27], ns = stencil_of(r, c, st);
        for (int p = beg; p < end; p++) {
            int i = r->cell_atoms[p];
            if (i >= r->nlocal) FAIL("ghost atom %d inside a local cell %d", i, c);
            int *row = r->pair_rows + (size_t)i * r->n_col;
            int n_core = 0, n_skin = 0;
            float xi = r->coord4[4 * i], yi = r->coord4[4 * i + 1], zi = r->coord4[4 * i + 2];
            for (int s = 0; s < ns; s++)
                for (int q = r->cell_start[st[s]]; q < r->cell_start[st[s] + 1]; q++) {
                    int j = r->cell_atoms[q];
                    float dx = xi - r->coord4[4 * j], dy = yi - r->coord4[4 * j + 1], dz = zi - r->coord4[4 * j + 2];
                    float dr2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));   /* nvcc -fmad contraction of dx*dx+dy*dy+dz*dz */
                    if (i == j) continue;
                    if (dr2 <= rc2_core) {
                        if (n_core + n_skin >= r->n_col) { free(skin); FAIL("pair table overflow, atom %d", i); }
                        row[n_core++] = j;
                    } else if (dr2 <= rc2_tail) {
                        if (n_core + n_skin >= r->n_col) { free(skin); FAIL("pair table overflow, atom %d", i); }
                        skin[n_skin++] = j;
                    }
                }
            for (int t = 0; t < n_skin; t++) row[n_core + t] = skin[n_skin - 1 - t]; /* join: reverse encounter order */
            r->pair_count[i] = n_core + n_skin;
        }
    }
    free(skin);
    return 0;
}

static void filter_exclusion(const orc_world *w, orc_rank *r);

int orc_rebuild(orc_world *w)
{
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        for (int i = 0; i < r->nlocal; i++) pbc_one(w, r->x + 3 * i, r->image + i);
        if (!r->bins_ready) setup_bins(w, r);      /* setup(): bins before exchange, UM/mvv_meso.cu:155-156 */
    }
    exchange(w);
    for (int ir = 0; ir < w->nranks; ir++) sort_local(w, &w->rk[ir]);
    borders(w);
    for (int ir = 0; ir < w->nranks; ir++) if (neighbor_build(w, &w->rk[ir])) return -1;
    if (w->bond_per_atom > 0 && w->special_lj12 == 0.0)
        for (int ir = 0; ir < w->nranks; ir++) filter_exclusion(w, &w->rk[ir]);
    w->ago = 0;
    return 0;
}

/* ---------------------------------------------------------------------- */
/* forces                                                                  */
/* ---------------------------------------------------------------------- */
void orc_force_clear(orc_world *w)
{
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        memset(r->f, 0, sizeof(double) * 3 * r->nlocal);
        memset(r->virial, 0, sizeof(double) * 6 * r->nlocal);
    }
}

/* gpu_dpd_fast, UM/pair_dpd_fast_meso.cu:91-205 */
static void force_sp(orc_world *w, orc_rank *r, int evflag)
{
    int nt = w->ntypes;
    float *cf = malloc(sizeof(float) * nt * nt * NCOEFF);
    for (int i = 0; i < nt * nt * NCOEFF; i++) cf[i] = (float)w->coeff[i];
    float dtis = (float)(1.0 / sqrt(w->dt));
    for (int i = 0; i < r->nlocal; i++) {
        const float *c1 = r->coord4 + 4 * i, *v1 = r->veloc4 + 4 * i;
        uint32_t ti = f2u(c1[3]), si = f2u(v1[3]);
        float fx = 0.f, fy = 0.f, fz = 0.f, vr[6] = {0, 0, 0, 0, 0, 0}, energy = 0.f;
        const int *row = r->pair_rows + (size_t)i * r->n_col;
        for (int p = 0; p < r->pair_count[i]; p++) {
            int j = row[p];
            const float *c2 = r->coord4 + 4 * j;
            float dx = c1[0] - c2[0], dy = c1[1] - c2[1], dz = c1[2] - c2[2];
            float rsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float *k = cf + (ti * nt + f2u(c2[3])) * NCOEFF;
            if (rsq < k[P_CUTSQ] && rsq >= 1.0E-20f) {
                const float *v2 = r->veloc4 + 4 * j;
                float rn = orc_gaussian_sp(si, f2u(v2[3]));
                float rinv = 1.0f / sqrtf(rsq);
                float rr = rsq * rinv;
                float dvx = v1[0] - v2[0], dvy = v1[1] - v2[1], dvz = v1[2] - v2[2];
                float dot = fmaf(dz, dvz, fmaf(dy, dvy, dx * dvx));
                float wc = 1.0f - rr * k[P_CUTINV];
                float wr = powf(wc, k[P_EXPW]);
                float fpair = k[P_A0] * wc - (k[P_GAMMA] * wr * wr * dot * rinv) + (k[P_SIGMA] * wr * rn * dtis);
                fpair *= rinv;
                fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
                if (evflag) {
                    vr[0] += dx * dx * fpair; vr[1] += dy * dy * fpair; vr[2] += dz * dz * fpair;
                    vr[3] += dx * dy * fpair; vr[4] += dx * dz * fpair; vr[5] += dy * dz * fpair;
                    energy += 0.5f * k[P_A0] * k[P_CUT] * wc * wc;
                }
            }
        }
        r->f[3 * i] += fx; r->f[3 * i + 1] += fy; r->f[3 * i + 2] += fz;
        if (evflag) {
            for (int q = 0; q < 6; q++) r->virial[6 * i + q] += vr[q] * 0.5f;
            r->e_pair[i] = energy * 0.5f;
        }
    }
    free(cf);
}

/* gpu_dpd, UM/pair_dpd_meso.cu:91-205: fp64 arithmetic on the fp32-packed inputs */
static void force_dp(orc_world *w, orc_rank *r, int evflag)
{
    int nt = w->ntypes;
    const double *cf = w->coeff;
    double dtis = 1.0 / sqrt(w->dt);
    for (int i = 0; i < r->nlocal; i++) {
        const float *c1 = r->coord4 + 4 * i, *v1 = r->veloc4 + 4 * i;
        uint32_t ti = f2u(c1[3]), si = f2u(v1[3]);
        double fx = 0., fy = 0., fz = 0., vr[6] = {0, 0, 0, 0, 0, 0}, energy = 0.;
        const int *row = r->pair_rows + (size_t)i * r->n_col;
        for (int p = 0; p < r->pair_count[i]; p++) {
            int j = row[p];
            const float *c2 = r->coord4 + 4 * j;
            /* float differences widened to double (f3u members are r32) */
            double dx = (float)(c1[0] - c2[0]), dy = (float)(c1[1] - c2[1]), dz = (float)(c1[2] - c2[2]);
            double rsq = fma(dz, dz, fma(dy, dy, dx * dx));
            const double *k = cf + (ti * nt + f2u(c2[3])) * NCOEFF;
            if (rsq < k[P_CUTSQ] && rsq >= 1.0E-20) {
                const float *v2 = r->veloc4 + 4 * j;
                double rn = orc_gaussian_dp(si, f2u(v2[3]));
                double rinv = 1.0 / sqrt(rsq);         /* CUDA rsqrt(double) */
                double rr = rsq * rinv;
                double dvx = (float)(v1[0] - v2[0]), dvy = (float)(v1[1] - v2[1]), dvz = (float)(v1[2] - v2[2]);
                double dot = fma(dz, dvz, fma(dy, dvy, dx * dvx));
                double wc = 1.0 - rr * k[P_CUTINV];
                double wr = orc_powd(wc, k[P_EXPW]);
                double fpair = k[P_A0] * wc - (k[P_GAMMA] * wr * wr * dot * rinv) + (k[P_SIGMA] * wr * rn * dtis);
                fpair *= rinv;
                fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
                if (evflag) {
                    vr[0] += dx * dx * fpair; vr[1] += dy * dy * fpair; vr[2] += dz * dz * fpair;
                    vr[3] += dx * dy * fpair; vr[4] += dx * dz * fpair; vr[5] += dy * dz * fpair;
                    energy += 0.5 * k[P_A0] * k[P_CUT] * wc * wc;
                }
            }
        }
        r->f[3 * i] += fx; r->f[3 * i + 1] += fy; r->f[3 * i + 2] += fz;
        if (evflag) {
            for (int q = 0; q < 6; q++) r->virial[6 * i + q] += vr[q] * 0.5;
            r->e_pair[i] = energy * 0.5;
        }
    }
}

/* MesoPairDPD::compute, UM/pair_dpd_meso.cu:241-270: pack with seed_now, then kernel over LOCAL */
void orc_pair_compute(orc_world *w, int eflag, int vflag)
{
    uint32_t sn = orc_seed_now(w->seed, (uint32_t)w->ntimestep);
    orc_pack(w, sn);
    for (int ir = 0; ir < w->nranks; ir++) {
        if (w->precision) force_dp(w, &w->rk[ir], eflag || vflag);
        else force_sp(w, &w->rk[ir], eflag || vflag);
    }
}

/* ---------------------------------------------------------------------- */
/* integration: UM/fix_nve_meso.cu:62-95,157-178                           */
/* ---------------------------------------------------------------------- */
void orc_initial_integrate(orc_world *w, int groupbit)
{
    double dtv = w->dt, dtf = 0.5 * w->dt * 1.0; /* ftm2v = 1 (lj) */
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        for (int i = 0; i < r->nlocal; i++) {
            if (!(r->mask[i] & groupbit)) continue;
            double dtfm = dtf * orc_rcp(w->mass[r->type[i]]);
            for (int d = 0; d < 3; d++) {
                r->v[3 * i + d] = fma(dtfm, r->f[3 * i + d], r->v[3 * i + d]);
                r->x[3 * i + d] = fma(dtv, r->v[3 * i + d], r->x[3 * i + d]);
            }
        }
    }
}
void orc_final_integrate(orc_world *w, int groupbit)
{
    double dtf = 0.5 * w->dt * 1.0;
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        for (int i = 0; i < r->nlocal; i++) {
            if (!(r->mask[i] & groupbit)) continue;
            double dtfm = dtf * orc_rcp(w->mass[r->type[i]]);
            for (int d = 0; d < 3; d++) r->v[3 * i + d] = fma(dtfm, r->f[3 * i + d], r->v[3 * i + d]);
        }
    }
}

/* MesoComputeTemp, UM/compute_temp_meso.cu:49-101; lj units: mvv2e = boltz = 1, extra_dof = 3 */
double orc_temperature(orc_world *w, int groupbit)
{
    double t = 0.0, natoms = 0.0;
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        for (int i = 0; i < r->nlocal; i++) {
            if (!(r->mask[i] & groupbit)) continue;
            const double *v = r->v + 3 * i;
            t += w->mass[r->type[i]] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            natoms += 1.0;
        }
    }
    double dof = 3.0 * natoms - 3.0;
    return dof > 0.0 ? t / dof : 0.0;
}

/* ---------------------------------------------------------------------- */
/* bead-spring topology                                 */
/* ---------------------------------------------------------------------- */
int orc_world_set_bonds(orc_world *w, int n, int bond_per_atom, const int *tag, const int *num_bond,
                        const int *bond_type, const int *bond_atom)
{
    int tmax = 0;
    for (int i = 0; i < n; i++) { int t = tag ? tag[i] : i + 1; if (t > tmax) tmax = t; }
    free(w->num_bond); free(w->bond_type); free(w->bond_atom);
    w->bond_per_atom = bond_per_atom; w->tag_max = tmax;
    w->num_bond = calloc(tmax + 1, sizeof(int));
    w->bond_type = calloc((size_t)(tmax + 1) * (bond_per_atom > 0 ? bond_per_atom : 1), sizeof(int));
    w->bond_atom = calloc((size_t)(tmax + 1) * (bond_per_atom > 0 ? bond_per_atom : 1), sizeof(int));
    for (int i = 0; i < n; i++) {
        int t = tag ? tag[i] : i + 1;
        if (num_bond[i] > bond_per_atom) FAIL("num_bond exceeds bond_per_atom for atom %d", i);
        w->num_bond[t] = num_bond[i];
        for (int p = 0; p < num_bond[i]; p++) {
            w->bond_type[(size_t)t * bond_per_atom + p] = bond_type[(size_t)i * bond_per_atom + p];
            w->bond_atom[(size_t)t * bond_per_atom + p] = bond_atom[(size_t)i * bond_per_atom + p];
        }
    }
    return 0;
}

void orc_set_bond_coeff(orc_world *w, int nbondtypes, const double *k, const double *r0)
{
    free(w->bond_k); free(w->bond_r0);
    w->nbondtypes = nbondtypes;
    w->bond_k = malloc(sizeof(double) * (nbondtypes + 1)); w->bond_r0 = malloc(sizeof(double) * (nbondtypes + 1));
    memcpy(w->bond_k, k, sizeof(double) * (nbondtypes + 1)); memcpy(w->bond_r0, r0, sizeof(double) * (nbondtypes + 1));
}

void orc_set_special_lj12(orc_world *w, double v) { w->special_lj12 = v; }

static int bonds_on(const orc_world *w) { return w->bond_per_atom > 0 && w->nbondtypes > 0; }

static double minimum_image(double dr, double p)          /* UM/math_meso.h:148-152 */
{
    double p_half = p * 0.5;
    return dr + (dr > -p_half ? (dr < p_half ? 0.0 : -p) : p);
}

/* gpu_filter_exclusion, UM/neigh_build_meso.cu:497-544, for special_bonds lj 0: bonded partners (by tag) leave the row */
static void filter_exclusion(const orc_world *w, orc_rank *r)
{
    for (int i = 0; i < r->nlocal; i++) {
        int t = r->tag[i], nb = w->num_bond[t];
        if (!nb) continue;
        int *row = r->pair_rows + (size_t)i * r->n_col, keep = 0;
        for (int k = 0; k < r->pair_count[i]; k++) {
            int tj = r->tag[row[k]], ok = 1;
            for (int p = 0; p < nb; p++) if (w->bond_atom[(size_t)t * w->bond_per_atom + p] == tj) ok = 0;
            if (ok) row[keep++] = row[k];
        }
        r->pair_count[i] = keep;
    }
}

/* gpu_bond_harmonic, UM/bond_harmonic_meso.cu:46-117; partner = lowest index holding the tag (gpu_set_map's atomicMin,
 * UM/atom_meso.cu:74-82), minimum image on the packed fp32 coordinates */
void orc_bond_compute(orc_world *w, int eflag, int vflag)
{
    if (!bonds_on(w)) return;
    int ev = eflag || vflag;
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        int nall = r->nlocal + r->nghost;
        int *map = malloc(sizeof(int) * (w->tag_max + 1));
        for (int t = 0; t <= w->tag_max; t++) map[t] = -1;
        for (int i = nall - 1; i >= 0; i--) if (r->tag[i] >= 0 && r->tag[i] <= w->tag_max) map[r->tag[i]] = i;
        double period[3];
        for (int d = 0; d < 3; d++) period[d] = w->periodic[d] ? w->prd[d] : 0.0;
        for (int i = 0; i < r->nlocal; i++) {
            int t = r->tag[i], nb = w->num_bond[t];
            const float *c1 = r->coord4 + 4 * i;
            double fx = 0, fy = 0, fz = 0, e = 0;
            for (int p = 0; p < nb; p++) {
                int j = map[w->bond_atom[(size_t)t * w->bond_per_atom + p]], ty = w->bond_type[(size_t)t * w->bond_per_atom + p];
                if (j < 0) continue;
                const float *c2 = r->coord4 + 4 * j;
                double dx = minimum_image((double)(float)(c2[0] - c1[0]), period[0]);
                double dy = minimum_image((double)(float)(c2[1] - c1[1]), period[1]);
                double dz = minimum_image((double)(float)(c2[2] - c1[2]), period[2]);
                double rsq = dx * dx + dy * dy + dz * dz;
                double rinv = 1.0 / sqrt(rsq), rr = rinv * rsq, dr = rr - w->bond_r0[ty];
                double fbond = 2.0 * w->bond_k[ty] * dr * rinv;
                fx += dx * fbond; fy += dy * fbond; fz += dz * fbond;
                e += w->bond_k[ty] * dr * dr;
            }
            r->f[3 * i] += fx; r->f[3 * i + 1] += fy; r->f[3 * i + 2] += fz;
            if (ev) {
                double *vv = r->virial + 6 * i;
                vv[0] += c1[0] * fx; vv[1] += c1[1] * fy; vv[2] += c1[2] * fz;
                vv[3] += c1[0] * fy; vv[4] += c1[0] * fz; vv[5] += c1[1] * fz;
                r->e_bond[i] = e * 0.5;
            }
        }
        free(map);
    }
}

double orc_bond_energy(orc_world *w)
{
    double e = 0;
    if (!bonds_on(w)) return 0;
    for (int ir = 0; ir < w->nranks; ir++) for (int i = 0; i < w->rk[ir].nlocal; i++) e += w->rk[ir].e_bond[i];
    return e;
}

/* ---------------------------------------------------------------------- */
/* channel fixes                                        */
/* ---------------------------------------------------------------------- */
enum { FIX_WALL = 1, FIX_SOLID_BOUND = 2, FIX_ADDFORCE = 3, FIX_POIS = 4, FIX_RDF = 5 };

int orc_fix_add(orc_world *w, int kind, int groupbit, int dims, const double *p4)
{
    if (w->nfix >= 8) FAIL("too many fixes");
    w->fix[w->nfix].kind = kind; w->fix[w->nfix].groupbit = groupbit; w->fix[w->nfix].dims = dims;
    for (int q = 0; q < 4; q++) w->fix[w->nfix].p[q] = p4 ? p4[q] : 0.0;
    w->fix[w->nfix].hist = NULL; w->fix[w->nfix].samples = 0;
    if (kind == FIX_RDF) w->fix[w->nfix].hist = calloc(dims > 0 ? dims : 1, sizeof(unsigned long long));   /* dims = nbin */
    return w->nfix++;
}
void orc_fix_clear(orc_world *w)
{
    for (int k = 0; k < w->nfix; k++) { free(w->fix[k].hist); w->fix[k].hist = NULL; }
    w->nfix = 0;
}

/* gpu_calc_rdf, UM/fix_rdf_fast_meso.cu:102-150: p = {every, j_groupbit, rc}; fp32 on the packed coordinates over the stored
 * neighbor table, bin = floorf(r * nbin / rc) with r = rsq * rsqrtf(rsq) (rsqrtf is approximate on the device: a distance
 * within an ulp of a bin edge may land in the neighboring bin there) */
static void rdf_sample(orc_world *w, int k)
{
    const int nbin = w->fix[k].dims, gi = w->fix[k].groupbit, gj = (int)w->fix[k].p[1];
    const float rc = (float)w->fix[k].p[2], bin_sz_inv = (float)nbin / rc;
    for (int ir = 0; ir < w->nranks; ir++) {
        orc_rank *r = &w->rk[ir];
        for (int i = 0; i < r->nlocal; i++) {
            if (!(r->mask[i] & gi)) continue;
            const float *c1 = r->coord4 + 4 * i;
            const int *row = r->pair_rows + (size_t)i * r->n_col;
            for (int q = 0; q < r->pair_count[i]; q++) {
                int j = row[q];
                if (!(r->mask[j] & gj)) continue;
                const float *c2 = r->coord4 + 4 * j;
                float dx = c1[0] - c2[0], dy = c1[1] - c2[1], dz = c1[2] - c2[2];
                float rsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (rsq < rc * rc) {
                    float rr = rsq * (1.0f / sqrtf(rsq));
                    int bid = (int)floorf(rr * bin_sz_inv);
                    if (bid >= 0 && bid < nbin) w->fix[k].hist[bid]++;
                }
            }
        }
    }
    w->fix[k].samples++;
}

/* histogram[nbin] as doubles, samples, sizes of the two groups */
int orc_fix_rdf_read(orc_world *w, int k, double *hist, double *samples, double *ni, double *nj)
{
    if (k < 0 || k >= w->nfix || w->fix[k].kind != FIX_RDF) FAIL("not an rdf fix");
    for (int b = 0; b < w->fix[k].dims; b++) hist[b] = (double)w->fix[k].hist[b];
    *samples = (double)w->fix[k].samples; *ni = 0; *nj = 0;
    for (int ir = 0; ir < w->nranks; ir++)
        for (int i = 0; i < w->rk[ir].nlocal; i++) {
            if (w->rk[ir].mask[i] & w->fix[k].groupbit) *ni += 1;
            if (w->rk[ir].mask[i] & (int)w->fix[k].p[1]) *nj += 1;
        }
    return 0;
}
void orc_set_integrate_group(orc_world *w, int groupbit) { w->integrate_groupbit = groupbit; }

/* Rho5rc1s1::operator(), UM/fix_solid_bound_meso.h:42-56 (nvcc contracts every s*h + c of the Horner form into one fma) */
static double rho5rc1s1(double h)
{
    double s = +0.282625;
    s = fma(s, h, -1.39021);
    s = fma(s, h, +2.70259);
    s = fma(s, h, -2.47678);
    s = fma(s, h, +0.863184);
    s = fma(s, h, +0.0664266);
    s = fma(s, h, +0.0247250);
    s = fma(s, h, +0.00856667);
    s = fma(s, h, -0.116714);
    s = fma(s * h, h, 0.0355959);
    return 75.0 * 6.2831853071796 * s;
}

/* Modify::post_force over the registered fixes: gpu_fix_wall_force UM/fix_wall_meso.cu:74-117, gpu_fix_solid_wall_force
 * UM/fix_solid_bound_meso.cu:72-113, gpu_fix_add_force UM/fix_addforce_meso.cu:72-90, gpu_fix_pois_post_force
 * UM/fix_poiseuille_meso.cu:71-92.  only < 0: all fixes */
void orc_fix_post_force(orc_world *w, int only)
{
    const double SQRT3 = 1.732050808;
    for (int k = 0; k < w->nfix; k++) {
        if (only >= 0 && only != k) continue;
        const int kind = w->fix[k].kind, gb = w->fix[k].groupbit, dims = w->fix[k].dims;
        const double *p = w->fix[k].p;
        if (kind == FIX_RDF) {                      /* MesoFixRDFFast::post_force, UM/fix_rdf_fast_meso.cu:152-180 */
            if (w->ntimestep % (long)(p[0] < 1 ? 1 : p[0]) == 0) rdf_sample(w, k);
            continue;
        }
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *r = &w->rk[ir];
            for (int i = 0; i < r->nlocal; i++) {
                if (!(r->mask[i] & gb)) continue;
                double *f = r->f + 3 * i; const double *x = r->x + 3 * i;
                if (kind == FIX_WALL) {
                    double d = p[0], dinv = 1.0 / d, ff = p[1];
                    if (ff == 0.0) continue;
                    for (int a = 0; a < 3; a++) {
                        if (!((dims >> a) & 1)) continue;
                        double h = x[a] - w->boxlo[a];
                        if (h <= d) f[a] = fma(ff, (double)erfcf((float)((h - 0.5 * d) * dinv * SQRT3)), f[a]);
                        h = w->boxhi[a] - x[a];
                        if (h <= d) f[a] = fma(-ff, (double)erfcf((float)((h - 0.5 * d) * dinv * SQRT3)), f[a]);
                    }
                } else if (kind == FIX_SOLID_BOUND) {
                    for (int a = 0; a < 3; a++) {
                        if (!((dims >> a) & 1)) continue;
                        double h = x[a] - w->boxlo[a];
                        if (h <= 1.0) f[a] += rho5rc1s1(h);
                        h = w->boxhi[a] - x[a];
                        if (h <= 1.0) f[a] -= rho5rc1s1(h);
                    }
                } else if (kind == FIX_ADDFORCE) {
                    f[0] += p[0]; f[1] += p[1]; f[2] += p[2];
                } else if (kind == FIX_POIS) {
                    int dim_ortho = dims & 3, dim_force = (dims >> 2) & 3;
                    double lower = w->boxlo[dim_ortho], upper = w->boxhi[dim_ortho];
                    double bisect = p[1] * upper + (1.0 - p[1]) * lower, rr = x[dim_ortho];
                    if ((rr < bisect && rr >= lower) || rr >= upper) f[dim_force] += p[0];
                    else f[dim_force] -= p[0];
                }
            }
        }
    }
}

/* bounce-forward, gpu_fix_wall_bounce UM/fix_wall_meso.cu:147-193 (= gpu_fix_solid_wall_bounce) */
void orc_fix_bounce(orc_world *w, int only)
{
    for (int k = 0; k < w->nfix; k++) {
        if (only >= 0 && only != k) continue;
        if (w->fix[k].kind != FIX_WALL && w->fix[k].kind != FIX_SOLID_BOUND) continue;
        for (int ir = 0; ir < w->nranks; ir++) {
            orc_rank *r = &w->rk[ir];
            for (int i = 0; i < r->nlocal; i++) {
                if (!(r->mask[i] & w->fix[k].groupbit)) continue;
                for (int a = 0; a < 3; a++) {
                    if (!((w->fix[k].dims >> a) & 1)) continue;
                    double *x = r->x + 3 * i + a, *v = r->v + 3 * i + a;
                    if (*x <= w->boxlo[a]) { *v = fabs(*v); *x = 2. * w->boxlo[a] - *x; }
                    else if (*x >= w->boxhi[a]) { *v = -fabs(*v); *x = 2. * w->boxhi[a] - *x; }
                }
            }
        }
    }
}

/* ---------------------------------------------------------------------- */
/* drivers: ModifiedVerlet::setup / ::run, UM/mvv_meso.cu:139-219,243-425  */
/* ---------------------------------------------------------------------- */
int orc_world_setup(orc_world *w, int eflag, int vflag)
{
    for (int ir = 0; ir < w->nranks; ir++) w->rk[ir].bins_ready = 0;
    if (orc_rebuild(w)) return -1;
    orc_force_clear(w);
    orc_pair_compute(w, eflag, vflag);
    orc_bond_compute(w, eflag, vflag);
    for (int k = 0; k < w->nfix; k++)              /* modify->setup -> Fix::setup -> post_force (UM/fix_wall_meso.cu:66-72); */
        if (w->fix[k].kind != FIX_RDF) orc_fix_post_force(w, k);   /* MesoFixRDFFast::setup is empty */
    return 0;
}

int orc_world_run(orc_world *w, int nsteps, int eflag, int vflag)
{
    const int gb = w->integrate_groupbit ? w->integrate_groupbit : 1;
    for (int s = 0; s < nsteps; s++) {
        w->ntimestep++;
        orc_initial_integrate(w, gb);
        w->ago++;                                  /* Neighbor::decide, neighbor.cpp:1216-1231 (delay 0, check no) */
        if (w->ago % w->every == 0) {
            orc_fix_bounce(w, -1);                 /* modify->pre_exchange, UM/mvv_meso.cu:273 */
            if (orc_rebuild(w)) return -1;
        }
        else orc_forward_comm(w);
        orc_force_clear(w);
        orc_pair_compute(w, eflag, vflag);
        orc_bond_compute(w, eflag, vflag);
        orc_fix_post_force(w, -1);                 /* modify->post_force, UM/mvv_meso.cu:396 */
        orc_final_integrate(w, gb);
        orc_fix_bounce(w, -1);                     /* modify->end_of_step, UM/mvv_meso.cu:399 */
    }
    return 0;
}

/* ---------------------------------------------------------------------- */
/* queries                                                                 */
/* ---------------------------------------------------------------------- */
void orc_counts(orc_world *w, int ir, int *nlocal, int *nghost, int *n_bulk, int *n_border, int *n_col)
{
    orc_rank *r = &w->rk[ir];
    *nlocal = r->nlocal; *nghost = r->nghost; *n_bulk = r->n_bulk; *n_border = r->n_border; *n_col = r->n_col;
}
void orc_bins(orc_world *w, int ir, int m[3], double binsize[3], double bininv[3])
{
    orc_rank *r = &w->rk[ir];
    for (int d = 0; d < 3; d++) { m[d] = r->m[d]; binsize[d] = r->binsize[d]; bininv[d] = r->bininv[d]; }
}
void orc_get_atoms(orc_world *w, int ir, double *x, double *v, double *f, int *tag, int *type, int *mask, int *image)
{
    orc_rank *r = &w->rk[ir]; int n = r->nlocal + r->nghost;
    if (x) memcpy(x, r->x, sizeof(double) * 3 * n);
    if (v) memcpy(v, r->v, sizeof(double) * 3 * n);
    if (f) memcpy(f, r->f, sizeof(double) * 3 * r->nlocal);
    if (tag) memcpy(tag, r->tag, sizeof(int) * n);
    if (type) memcpy(type, r->type, sizeof(int) * n);
    if (mask) memcpy(mask, r->mask, sizeof(int) * n);
    if (image) memcpy(image, r->image, sizeof(int) * r->nlocal);
}
void orc_get_packed(orc_world *w, int ir, float *coord4, float *veloc4)
{
    orc_rank *r = &w->rk[ir]; int n = r->nlocal + r->nghost;
    memcpy(coord4, r->coord4, sizeof(float) * 4 * n); memcpy(veloc4, r->veloc4, sizeof(float) * 4 * n);
}
void orc_get_virial(orc_world *w, int ir, double *virial6, double *e_pair)
{
    orc_rank *r = &w->rk[ir];
    if (virial6) memcpy(virial6, r->virial, sizeof(double) * 6 * r->nlocal);
    if (e_pair) memcpy(e_pair, r->e_pair, sizeof(double) * r->nlocal);
}
void orc_get_reorder(orc_world *w, int ir, uint64_t *key, int *permute_from)
{
    orc_rank *r = &w->rk[ir];
    if (key) memcpy(key, r->key, sizeof(uint64_t) * r->nlocal);
    if (permute_from) memcpy(permute_from, r->perm_from, sizeof(int) * r->nlocal);
}
void orc_get_cells(orc_world *w, int ir, int *cell_start, int *cell_atoms)
{
    orc_rank *r = &w->rk[ir];
    if (cell_start) memcpy(cell_start, r->cell_start, sizeof(int) * (r->ncell + 1));
    if (cell_atoms) memcpy(cell_atoms, r->cell_atoms, sizeof(int) * (r->nlocal + r->nghost));
}
void orc_get_neighbors(orc_world *w, int ir, int *pair_count, int *pair_rows)
{
    orc_rank *r = &w->rk[ir];
    if (pair_count) memcpy(pair_count, r->pair_count, sizeof(int) * r->nlocal);
    if (pair_rows) memcpy(pair_rows, r->pair_rows, sizeof(int) * (size_t)r->nlocal * r->n_col);
}
/* UM/neigh_list_meso.cu:97-102: k-th neighbor of i at ((i&~31)+(k&31))*n_col + (k>>5)*32 + (i&31) */
void orc_get_neighbors_transposed(orc_world *w, int ir, int *pair_table)
{
    orc_rank *r = &w->rk[ir];
    for (int i = 0; i < r->nlocal; i++)
        for (int k = 0; k < r->pair_count[i]; k++)
            pair_table[(size_t)((i & ~31) + (k & 31)) * r->n_col + (k >> 5) * 32 + (i & 31)] = r->pair_rows[(size_t)i * r->n_col + k];
}
