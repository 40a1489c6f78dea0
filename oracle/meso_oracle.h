/* meso_oracle.h -- CPU restatement of the USER-MESO DPD time-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under meso_b200/ may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src, UM/ = USER-MESO/).  Integer work (TEA, signatures, sort
 * keys, cells, neighbor lists) is restated bit-exactly; floating-point work is
 * restated with the contraction spelled out (fma/fmaf) where membership
 * decisions depend on it.
 *
 * Pinning: the reference ships no tests/golden vectors for this path
 * (SURVEY.md s4).  The integer core is pinned against the reference's own
 * __host__ __device__ code compiled host-side (oracle/_ref/libref_math.so, see
 * oracle/Makefile) and the known-answer vectors recorded from it in
 * tests/golden/tea_kat.json; conservative forces are pinned against stock
 * LAMMPS pair_style dpd built from the same tree (oracle/_ref/lmp_serial).
 */
#ifndef MESO_ORACLE_H
#define MESO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- integer core (UM/math_meso.h:166-178, 436-464) ---- */
void     orc_tea(uint32_t *v0, uint32_t *v1, int rounds);
uint32_t orc_premix_tea(uint32_t v0, uint32_t v1, int rounds);
uint32_t orc_bit_space3(uint32_t x);
uint32_t orc_interleave3(uint32_t i, uint32_t j, uint32_t k);
uint32_t orc_mantissa(float u, float v, float w);
uint32_t orc_brev(uint32_t x);
uint32_t orc_seed_now(uint32_t seed, uint32_t ntimestep);
uint32_t orc_signature(uint32_t seed_now, int tag, float vx, float vy, float vz);

/* ---- per-pair Gaussians (UM/math_meso.h:466-484) ---- */
float  orc_gaussian_sp(uint32_t sig_i, uint32_t sig_j);
double orc_gaussian_dp(uint32_t sig_i, uint32_t sig_j);

/* ---- branch-free fp64 transcendentals (UM/math_meso.h:204-424) ---- */
double orc_rsqrt(double x);
double orc_sqrtd(double x);
double orc_rcp(double x);
double orc_log2d_frac(double x);
double orc_exp2d_frac(double x);
double orc_powd(double a, double b);
double orc_sinpi(double x);
double orc_cospi(double x);
double orc_log2u(uint32_t x);

/* ---- world: R = px*py*pz simulated ranks in one process ---- */
typedef struct orc_world orc_world;

/* coeff7: [ntypes*ntypes][7] = {cut,cutsq,cutinv,expw,a0,gamma,sigma}
 * (UM/pair_dpd_meso.h:15-24).  mass: [ntypes+1], 1-based like LAMMPS.
 * precision: 0 = dpd/fast/meso (fp32), 1 = dpd/meso (fp64). */
orc_world *orc_world_create(const double boxlo[3], const double boxhi[3],
                            const int periodic[3], const int procgrid[3],
                            int ntypes, const double *mass, const double *coeff7,
                            double cut_max, double skin, int every,
                            uint32_t seed, double dt, int precision);
void orc_world_destroy(orc_world *w);
const char *orc_last_error(void);

/* atoms in input order; x,v AoS [n][3] (LAMMPS host layout) */
int orc_world_set_atoms(orc_world *w, int n, const double *x, const double *v,
                        const int *tag, const int *type, const int *mask,
                        const int *image);

/* ModifiedVerlet::setup (UM/mvv_meso.cu:139-219) and ::run (:243-425) */
int orc_world_setup(orc_world *w, int eflag, int vflag);
int orc_world_run(orc_world *w, int nsteps, int eflag, int vflag);

/* fine-grained phases, same order as UM/mvv_meso.cu:256-418 */
void orc_initial_integrate(orc_world *w, int groupbit);
void orc_final_integrate(orc_world *w, int groupbit);
int  orc_rebuild(orc_world *w);       /* pbc, exchange, sort_local, borders, neighbor build */
void orc_forward_comm(orc_world *w);
void orc_force_clear(orc_world *w);
void orc_pack(orc_world *w, uint32_t seed_now);   /* dp2sp_merged over all atoms */
void orc_pair_compute(orc_world *w, int eflag, int vflag); /* uses w->ntimestep */
double orc_temperature(orc_world *w, int groupbit);        /* scalar T, lj units */
void orc_set_timestep(orc_world *w, long ntimestep);
long orc_get_timestep(orc_world *w);

/* bead-spring topology: per-atom tables in LAMMPS' layout for the atoms of set_atoms (same order);
 * harmonic bonds UM/bond_harmonic_meso.cu:46-117, exclusion filter UM/neigh_build_meso.cu:497-544 */
int  orc_world_set_bonds(orc_world *w, int n, int bond_per_atom, const int *tag, const int *num_bond,
                         const int *bond_type, const int *bond_atom);
void orc_set_bond_coeff(orc_world *w, int nbondtypes, const double *k, const double *r0);
void orc_set_special_lj12(orc_world *w, double v);
void orc_bond_compute(orc_world *w, int eflag, int vflag);
double orc_bond_energy(orc_world *w);

/* channel fixes: kind 1 wall/meso {d, f}, 2 solid_bound/meso (rho5rc1s1), 3 addforce/meso {fx,fy,fz},
 * 4 pois/meso {strength, bisect_frac} with dims = dim_ortho | dim_force << 2; wall kinds: dims bit a = walls across a.
 * UM/fix_wall_meso.cu, UM/fix_solid_bound_meso.{h,cu}, UM/fix_addforce_meso.cu, UM/fix_poiseuille_meso.cu */
int  orc_fix_add(orc_world *w, int kind, int groupbit, int dims, const double *p4);
/* kind 5 rdf/fast/meso: dims = nbin, p = {every, j_groupbit, rc} (UM/fix_rdf_fast_meso.cu:102-180) */
int  orc_fix_rdf_read(orc_world *w, int k, double *hist, double *samples, double *ni, double *nj);
void orc_fix_clear(orc_world *w);
void orc_fix_post_force(orc_world *w, int only);   /* only < 0: every fix in registration order */
void orc_fix_bounce(orc_world *w, int only);
void orc_set_integrate_group(orc_world *w, int groupbit);   /* group of fix nve/meso used by orc_world_run (default 1) */

/* queries (rank r) */
int  orc_nranks(orc_world *w);
void orc_counts(orc_world *w, int r, int *nlocal, int *nghost, int *n_bulk,
                int *n_border, int *n_col);
void orc_bins(orc_world *w, int r, int m[3], double binsize[3], double bininv[3]);
void orc_get_atoms(orc_world *w, int r, double *x, double *v, double *f,
                   int *tag, int *type, int *mask, int *image); /* nlocal+nghost */
void orc_get_packed(orc_world *w, int r, float *coord4, float *veloc4);
void orc_get_virial(orc_world *w, int r, double *virial6, double *e_pair); /* per local atom */
void orc_get_reorder(orc_world *w, int r, uint64_t *key, int *permute_from); /* last sort_local, nlocal */
void orc_get_cells(orc_world *w, int r, int *cell_start /*ncell+1*/, int *cell_atoms /*nall*/);
int  orc_get_stencil(orc_world *w, int r, int cell, int *out27);
/* neighbor list: counts [nlocal], rows row-major [nlocal][n_col] (entries = local/ghost indices) */
void orc_get_neighbors(orc_world *w, int r, int *pair_count, int *pair_rows);
/* same table in the reference's tile-transposed layout (UM/neigh_list_meso.cu:97-102) */
void orc_get_neighbors_transposed(orc_world *w, int r, int *pair_table /* ceil32(nlocal)*n_col */);

#ifdef __cplusplus
}
#endif
#endif
