/* meso_b200.h -- C ABI of the B200-native USER-MESO DPD time-step library.
 *
 * This is the drop-in boundary: the LAMMPS package in lammps/USER-MESO-B200/
 * (same style names as the reference: dpd/atomic/meso, mvv/meso, dpd/meso,
 * dpd/fast/meso, nve/meso, temp/meso) and the Python mirror in
 * meso_b200/engine.py call nothing but these functions.  Plain pointers and
 * sizes only; every function returns 0 on success or a negative MESO_E* code
 * (text from meso_last_error).  One caller thread per context; calls are
 * asynchronous on the context's stream unless they return host data.
 *
 * Each entry cites the reference interface it replaces
 * (paths relative to /root/reference/src, UM/ = USER-MESO/).
 *
 * There is no CPU fallback: every compute entry fails with MESO_ENODEV when no
 * sm_100 device is usable.
 */
#ifndef MESO_B200_H
#define MESO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct meso_ctx meso_ctx;

enum {
    MESO_OK = 0,
    MESO_ENODEV = -1,    /* no usable CUDA device */
    MESO_ECUDA = -2,     /* CUDA runtime error */
    MESO_EINVAL = -3,    /* bad argument / call order */
    MESO_ECAPACITY = -4, /* ghost / pair-table capacity exceeded (UM/neigh_build_meso.cu:242-252 only printf'd) */
    MESO_ENCCL = -5
};

/* atom ranges, UM/util_meso.h:43-74 (AtomAttribute::BULK/BORDER/LOCAL/GHOST) */
enum { MESO_BULK = 1, MESO_BORDER = 2, MESO_LOCAL = 3, MESO_GHOST = 4, MESO_ALL = 7 };
/* pair precision: dpd/fast/meso (UM/pair_dpd_fast_meso.cu) vs dpd/meso (UM/pair_dpd_meso.cu) */
enum { MESO_SP = 0, MESO_DP = 1 };

/* ---- device runtime: MesoDevice, UM/engine_meso.cu:38-119 ---- */
int  meso_device_count(void);                       /* src/lammps.cpp:432-452 device pick */
int  meso_create(meso_ctx **out, int device);       /* new MesoDevice(lmp, gpu, profile) */
/* Several GPUs of one node behind ONE handle (one process, one host thread per GPU): replaces the reference's one MPI rank
 * per GPU (src/lammps.cpp:432-452 device pick, src/comm.cpp:201-287 processor grid, UM/comm_meso.cu:41-254 exchange/borders
 * through the host) for a host that is a single process.  meso_set_box splits the box into ndev uniform bricks, uploads are
 * dealt out by brick, downloads come back brick after brick, reductions are whole-box sums; every other entry point behaves as
 * on a single device except the meso_export_* parity exports (per-brick data: not available).  ndev == 1: meso_create. */
int  meso_create_gang(meso_ctx **out, int ndev, const int *devices);
int  meso_gang_size(meso_ctx *ctx);                 /* bricks behind the handle (1 for meso_create) */
/* the gang's host-side rules without a device: processor grid for ndev bricks (Comm::set_procs' surface rule,
 * src/comm.cpp:201-287) and the brick (x-major rank) every position x[3n] is dealt to (Domain::set_local_box's uniform split) */
int  meso_gang_layout(int ndev, const double boxlo[3], const double boxhi[3], const int periodic[3], int grid[3], int n,
                      const double *x, int *owner);
void meso_destroy(meso_ctx *ctx);                   /* MesoDevice::destroy */
const char *meso_last_error(meso_ctx *ctx);         /* NULL ctx: last creation error */
int  meso_sync(meso_ctx *ctx);                      /* MesoDevice::sync_device */
void *meso_stream(meso_ctx *ctx);                   /* cudaStream_t of the compute stream, UM/engine_meso.h stream() */
int  meso_profiler(meso_ctx *ctx, int start);       /* MesoDevice::configure_profiler, UM/engine_meso.cu:155-191 */
int  meso_memory_usage(meso_ctx *ctx, uint64_t *bytes); /* MesoDevice::print_memory_usage, UM/engine_meso.cu:123-144 */

/* ---- domain + decomposition: Domain/MesoDomain, Comm::setup (src/comm.cpp:393-640) ---- */
int meso_set_box(meso_ctx *ctx, const double boxlo[3], const double boxhi[3], const int periodic[3]);
/* rank's brick in a procgrid; rank order (ix*py+iy)*pz+iz.  nccl_id: 128-byte ncclUniqueId shared by all
 * ranks (from meso_comm_unique_id on rank 0), NULL when nranks==1.  Replaces the MPI world of the reference. */
int meso_comm_unique_id(void *id128);
int meso_set_decomposition(meso_ctx *ctx, int rank, const int procgrid[3], const void *nccl_id);
/* Halo bootstrap without NCCL (nccl_id == NULL above): ranks that live in ONE process (one context per GPU) or share ONE GPU.
 * The halo itself never uses NCCL: every rank owns a receive arena in its HBM and its neighbors store their records there
 * directly (CUDA IPC between processes, peer pointers inside one; meso_b200/csrc/comm.cu).  After meso_atoms_upload (and
 * meso_bonds_upload) on every rank: blob[r] = meso_comm_export(rank r) (meso_comm_blob_size() bytes each), then every rank
 * meso_comm_import(all blobs in rank order), then meso_setup.  Reductions (meso_compute_ke ...) then return the rank's own
 * part, as with meso_set_reduce_scope(1).  Replaces MPI_Init / the communicator of the reference (src/lammps.cpp:432-452). */
int meso_comm_blob_size(void);
int meso_comm_export(meso_ctx *ctx, void *blob);
int meso_comm_import(meso_ctx *ctx, const void *blobs, int nranks);

/* ---- styles' settings ---- */
/* neighbor <skin> bin; neigh_modify delay 0 every N check no (src/neighbor.cpp:1216-1231) */
int meso_set_neighbor(meso_ctx *ctx, double skin, int every);
/* mass per type, 1-based [ntypes+1] (Atom::mass; unpack_by_type UM/atom_vec_meso.h:90-104) */
int meso_set_types(meso_ctx *ctx, int ntypes, const double *mass);
/* pair_style dpd/meso|dpd/fast/meso <cut_global> <seed>: MesoPairDPD::settings UM/pair_dpd_meso.cu:272-288 */
int meso_pair_dpd_settings(meso_ctx *ctx, int precision, double cut_global, int seed);
/* coefficient table [ntypes*ntypes][7] = {cut,cutsq,cutinv,expw,a0,gamma,sigma}:
 * MesoPairDPD::prepare_coeff UM/pair_dpd_meso.cu:68-89, layout UM/pair_dpd_meso.h:15-24 */
int meso_pair_dpd_coeff(meso_ctx *ctx, const double *coeff7);
int meso_set_timestep_size(meso_ctx *ctx, double dt);     /* update->dt; FixNVEMeso::reset_dt */
/* force->ftm2v of the host's unit system (FixNVEMeso::init UM/fix_nve_meso.cu:42-46: dtf = 0.5*dt*ftm2v); default 1 (lj) */
int meso_set_force_units(meso_ctx *ctx, double ftm2v);
int meso_set_ntimestep(meso_ctx *ctx, int64_t ntimestep); /* update->ntimestep */
int64_t meso_get_ntimestep(meso_ctx *ctx);

/* ---- atom store: MesoAtomVec / MesoAtom::transfer_*, UM/atom_meso.cu:152-266 ---- */
/* host AoS (LAMMPS layout x[n][3]) -> device SoA; replaces all local atoms of this rank.
 * v/tag/type/mask/image may be NULL (0, 1..n, 1, 1, centre image). */
int meso_atoms_upload(meso_ctx *ctx, int nlocal, const double *x, const double *v, const int *tag,
                      const int *type, const int *mask, const int *image);
/* transfer_pre_output: device SoA -> host AoS for LOCAL atoms in device (sorted) order; any pointer may be NULL.
 * Returns after the copy completed. */
int meso_atoms_download(meso_ctx *ctx, int nmax, double *x, double *v, double *f, int *tag, int *type,
                        int *mask, int *image);
/* page-lock / release a borrowed host array in place: Pinned<T>, UM/memory_meso.h:224-243 (cudaHostRegister) */
int meso_host_register(meso_ctx *ctx, void *ptr, uint64_t bytes);
int meso_host_unregister(meso_ctx *ctx, void *ptr);
/* counts after the last rebuild (host mirror; syncs the stream) */
int meso_counts(meso_ctx *ctx, int *nlocal, int *nghost, int *n_bulk, int *n_border);
int64_t meso_natoms_global(meso_ctx *ctx);

/* ---- time-step phases, in the order of ModifiedVerlet::run UM/mvv_meso.cu:243-425 ---- */
/* FixNVEMeso::initial_integrate UM/fix_nve_meso.cu:62-155 (dtf, dtv from meso_set_timestep_size) */
int meso_initial_integrate(meso_ctx *ctx, int groupbit);
/* Neighbor::decide: 1 if this step rebuilds (advances the "ago" counter) */
int meso_neighbor_decide(meso_ctx *ctx);
/* rebuild: Domain::pbc + Comm::exchange + MesoAtom::sort_local + MesoComm::borders + map + MesoNeighbor::build
 * (UM/mvv_meso.cu:270-326).  All on device. */
int meso_rebuild(meso_ctx *ctx);
/* Comm::forward_comm with ghost velocities (src/comm.cpp:686-753; UM/mvv_meso.cu:338-358) */
int meso_forward_comm(meso_ctx *ctx);
/* MesoAtomVec::force_clear UM/atom_vec_meso.cu:325-336 */
int meso_force_clear(meso_ctx *ctx, int range, int vflag);
/* MesoPairDPD::compute / compute_bulk / compute_border UM/pair_dpd_meso.cu:241-266:
 * packs (dp2sp_merged with seed_now = premix_TEA<64>(seed, ntimestep)) then runs the force kernel on `range`. */
int meso_pair_compute(meso_ctx *ctx, int range, int eflag, int vflag);
/* FixNVEMeso::final_integrate UM/fix_nve_meso.cu:157-198 */
int meso_final_integrate(meso_ctx *ctx, int groupbit);
/* MesoComputeTemp::compute_scalar UM/compute_temp_meso.cu:77-101: sum_i m v^2 over group (all ranks), and group count */
int meso_compute_ke(meso_ctx *ctx, int groupbit, double *mv2_sum, double *count);
/* local_only = 1: meso_compute_ke / meso_compute_virial return this rank's part only, for a host that sums over its
 * own MPI ranks (MPI_Allreduce in UM/compute_temp_meso.cu:97); default 0 = summed over all ranks of the decomposition */
int meso_set_reduce_scope(meso_ctx *ctx, int local_only);
/* virial[6] (xx,yy,zz,xy,xz,yz) and pair energy summed over local atoms, all ranks
 * (the reference accumulates per-atom virial UM/pair_dpd_meso.cu:180-186 but never reduces it) */
int meso_compute_virial(meso_ctx *ctx, double virial6[6], double *e_pair);

/* ---- bead-spring topology: atom_style dpd/bond/meso + bond_style harmonic/meso ----
 * On a decomposition every rank makes the same calls with the same bond_per_atom / tag_max (a rank without bonded atoms
 * passes its zero counts): the table rides the migration messages, partners across a brick face are found among the ghosts. */
/* bond_coeff N k r0 for N = 1..nbondtypes, arrays [nbondtypes+1] (MesoBondHarmonic::alloc_coeff UM/bond_harmonic_meso.cu:34-44) */
int meso_bond_harmonic_coeff(meso_ctx *ctx, int nbondtypes, const double *k, const double *r0);
/* special_bonds lj <w12> ...: 1 keeps 1-2 pairs in the neighbor list, 0 filters them out
 * (MesoNeighbor::filter_exclusion_meso UM/neigh_build_meso.cu:546-569) */
int meso_set_special_bonds(meso_ctx *ctx, double lj12);
/* per-atom bond table in LAMMPS' host layout (Atom::num_bond[i], bond_type[i][slot], bond_atom[i][slot] = partner TAG,
 * rows of bond_per_atom ints), for the atoms just passed to meso_atoms_upload, newton_bond off (both atoms hold the bond):
 * AtomVecDPDBond device table UM/atom_vec_dpd_bond_meso.h:33-36.  tag_max = largest tag (size of the tag -> index array,
 * MesoAtom::map_set_device UM/atom_meso.cu:109-128). */
int meso_bonds_upload(meso_ctx *ctx, int nlocal, int bond_per_atom, const int *num_bond, const int *bond_type,
                      const int *bond_atom, int tag_max);
/* MesoBondHarmonic::compute UM/bond_harmonic_meso.cu:119-170: f += bonded forces (tallies per-atom energy/virial if flagged) */
int meso_bond_compute(meso_ctx *ctx, int eflag, int vflag);
int meso_compute_bond_energy(meso_ctx *ctx, double *e_bond);   /* all ranks (collective) unless meso_set_reduce_scope(1) */

/* ---- device-resident fixes of the channel decks: each call registers one fix and returns its handle (>= 0) ---- */
/* fix ID group wall/meso [x] [y] [z] d <d> f <f> (MesoFixWall, UM/fix_wall_meso.cu:20-46): dims bit 0/1/2 = x/y/z;
 * post_force: f -/+= F erfcf(sqrt3 (h - d/2)/d) within d of the box faces (:74-117); bounce-forward at the faces in
 * pre_exchange and end_of_step (:147-238) */
int meso_fix_wall(meso_ctx *ctx, int groupbit, int dims, double d, double f);
/* fix ID group solid_bound/meso [x] [y] [z] rho5rc1s1 (MesoFixSolidBound, UM/fix_solid_bound_meso.cu:20-46;
 * force_kernel 1 = Rho5rc1s1, UM/fix_solid_bound_meso.h:42-63) */
int meso_fix_solid_bound(meso_ctx *ctx, int groupbit, int dims, int force_kernel);
/* fix ID group addforce/meso fx fy fz (MesoFixAddForce, UM/fix_addforce_meso.cu:20-90) */
int meso_fix_addforce(meso_ctx *ctx, int groupbit, double fx, double fy, double fz);
/* fix ID group pois/meso <dim_ortho> <dim_force> <strength> [bisect_frac = 0.5] (MesoFixPoiseuille,
 * UM/fix_poiseuille_meso.cu:20-113): +strength below the bisection plane, -strength above it */
int meso_fix_pois(meso_ctx *ctx, int groupbit, int dim_ortho, int dim_force, double strength, double bisect_frac);
/* fix ID group rdf/fast/meso output <file> nbin <n> [every <k>] [other <group>] (MesoFixRDFFast, UM/fix_rdf_fast_meso.cu:42-180):
 * every `every` steps, in the post_force slot, histogram of the distances r < pair cutoff over the neighbor table between
 * atoms of `groupbit` (rows) and `j_groupbit` (entries); call after meso_pair_dpd_settings (rc = the pair style's global cutoff) */
int meso_fix_rdf(meso_ctx *ctx, int groupbit, int j_groupbit, int nbin, int every);
/* accumulated pair counts per bin (as doubles), number of samples, sizes of the two groups (the inputs of MesoFixRDFFast::dump,
 * UM/fix_rdf_fast_meso.cu:182-219); all ranks (collective) unless meso_set_reduce_scope(1) */
int meso_fix_rdf_read(meso_ctx *ctx, int handle, int nbin, double *histogram, double *n_samples, double *ni, double *nj);
int meso_fix_clear(meso_ctx *ctx);                                  /* unfix: drops every registered fix */
/* hooks for a host that drives the step by phases (Modify::post_force / pre_exchange / end_of_step, UM/mvv_meso.cu:273,396,399);
 * handle < 0 applies every registered fix in registration order.  meso_setup and meso_run call them themselves. */
int meso_fix_post_force(meso_ctx *ctx, int handle);
int meso_fix_bounce(meso_ctx *ctx, int handle);

/* ---- whole-run drivers (same results as the phase calls above, fewer launches) ---- */
/* ModifiedVerlet::setup UM/mvv_meso.cu:139-219 */
int meso_setup(meso_ctx *ctx, int eflag, int vflag);
/* ModifiedVerlet::run(n) for the deck's fix list {nve/meso on group `groupbit`}.
 * Evaluates each local pair once and adds both halves with atomic reductions (fp32 style: fp32 atomics into a per-atom
 * accumulator; fp64 style: fp64 atomics): every addend is bit-identical to the two-sided evaluation, the ORDER of the additions
 * is not fixed, so results differ in the last bits from run to run (the reference adds per-atom sums in a fixed order).
 * MESO_PAIR_ONCE=0 in the environment selects the deterministic two-sided kernel (bit-reproducible; what the bit-for-bit
 * tests use). */
int meso_run(meso_ctx *ctx, int nsteps, int groupbit);

/* ---- exports for parity tests (device -> host, synchronous) ---- */
int meso_export_bins(meso_ctx *ctx, int m[3], double binsize[3], double bininv[3], int *n_col); /* setup_bins */
int meso_export_reorder(meso_ctx *ctx, int nmax, uint64_t *key_sorted, int *permute_from);      /* sort_local */
int meso_export_packed(meso_ctx *ctx, int nmax, float *coord4, float *veloc4);                  /* dp2sp_merged, nlocal+nghost */
int meso_export_ghosts(meso_ctx *ctx, int nmax, double *x, double *v, int *tag, int *type);     /* borders */
int meso_export_cells(meso_ctx *ctx, int ncell_plus1, int *cell_start, int nmax, int *cell_atoms); /* binning_meso */
int meso_export_stencil(meso_ctx *ctx, int cell, int out27[27]);                                /* gpu_stencil_full_bin_3d */
int meso_export_pair_count(meso_ctx *ctx, int nmax, int *pair_count);
/* tile-transposed table, UM/neigh_list_meso.cu:97-102: ceil32(nlocal)*n_col ints, rows in the reference's order
 * (core entries in stencil order, then skin entries reversed: UM/neigh_build_meso.cu:58-117,166-200) */
int meso_export_pair_table(meso_ctx *ctx, int64_t nmax, int *pair_table);
/* the same table as the force kernels read it: rows [owned][other] (same sets; traversal order inside a part);
 * owned_count[i] = leading entries whose pair row i evaluates in the pair-once force kernel: ghost partners, and local
 * partners j with (i+j) odd ? i<j : i>j.  Any pointer may be NULL. */
int meso_export_pair_rows(meso_ctx *ctx, int64_t nmax, int *pair_table, int *owned_count);
int meso_export_virial(meso_ctx *ctx, int nmax, double *virial6, double *e_pair);
/* device-side evaluation of the per-pair Gaussians on n signature pairs (A8) */
int meso_eval_gaussian(meso_ctx *ctx, int n, const uint32_t *sig_i, const uint32_t *sig_j, float *out_sp, double *out_dp);
/* device-side evaluation of fn in {0:rsqrt,1:rcp,2:log2d_frac,3:exp2d_frac,4:sinpi,5:cospi,6:sqrtd} and powd(a,b) (A9) */
int meso_eval_math(meso_ctx *ctx, int fn, int n, const double *a, const double *b, double *out);
int meso_eval_log2u(meso_ctx *ctx, int n, const uint32_t *a, double *out);

/* ---- timing helper: CUDA-event time of a named phase accumulated on the compute stream ---- */
enum { MESO_T_INTEGRATE = 0, MESO_T_FORWARD, MESO_T_PAIR, MESO_T_REBUILD, MESO_T_NEIGH, MESO_T_COUNT };
int meso_timers_enable(meso_ctx *ctx, int on);       /* inserts event pairs around phases inside meso_run */
int meso_timers_read(meso_ctx *ctx, double ms[MESO_T_COUNT], int64_t calls[MESO_T_COUNT], int reset);
/* number of kernels this library launched on the context's streams since the last reset (NCCL's own kernels not included) */
int meso_launch_count(meso_ctx *ctx, int64_t *n, int reset);

#ifdef __cplusplus
}
#endif
#endif
