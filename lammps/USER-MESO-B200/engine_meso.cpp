#include "mpi.h"
#include "stdlib.h"
#include "string.h"
#include "engine_meso.h"
#include "meso_bridge.h"
#include "atom.h"
#include "atom_vec.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "neighbor.h"
#include "timer.h"
#include "universe.h"
#include "update.h"

using namespace LAMMPS_NS;

/* ---------------------------------------------------------------------- */

MesoDevice::MesoDevice(LAMMPS *lmp, int device, std::string profile) :
  Pointers(lmp), dummy(false), ctx(NULL), profile_mode(profile),
  profile_lo(0), profile_hi(0), profiling(false), on_device(false), npinned(0), pinned_nmax(0),
  bonded(false), topo_tag_max(0), topo_bpa(0), topo_maxspecial(0)
{
  int ndev = meso_device_count();
  if (ndev <= 0)
    error->one(FLERR,"<MESO> no CUDA device found: USER-MESO-B200 has no CPU path (run with -meso off for stock styles)");
  if (device < 0 || device >= ndev) error->one(FLERR,"<MESO> -device index out of range");
  // MESO_DEVICES=0-7 | 0,1,2,3 | 0,0 : this ONE process drives several GPUs (the image has no MPI; `-device N` of
  // src/lammps.cpp:177-188 carries a single integer).  The library splits the box into one brick per listed device
  // (a device may be listed more than once: several bricks share it) and LAMMPS keeps seeing one rank with all atoms.
  int devs[64], ngang = 0;
  const char *list = getenv("MESO_DEVICES");
  if (list && *list) {
    // (Comm does not exist yet when src/lammps.cpp:455-465 builds this object: Universe does)
    if (universe->nprocs > 1) error->all(FLERR,"<MESO> MESO_DEVICES is for single-process runs: with MPI use one rank per GPU");
    const char *s = list;
    while (*s && ngang < 64) {
      char *end;
      long a = strtol(s,&end,10), b = a;
      if (end == s) error->one(FLERR,"<MESO> MESO_DEVICES: expected a list such as 0-7 or 0,1,2,3");
      s = end;
      if (*s == '-') { b = strtol(s+1,&end,10); if (end == s+1) error->one(FLERR,"<MESO> MESO_DEVICES: bad range"); s = end; }
      for (long d = a; d <= b && ngang < 64; d++) {
        if (d < 0 || d >= ndev) error->one(FLERR,"<MESO> MESO_DEVICES: device index out of range");
        devs[ngang++] = (int) d;
      }
      if (*s == ',') s++;
      else if (*s) error->one(FLERR,"<MESO> MESO_DEVICES: expected a list such as 0-7 or 0,1,2,3");
    }
  }
  int rc = ngang > 1 ? meso_create_gang(&ctx,ngang,devs) : meso_create(&ctx,ngang == 1 ? devs[0] : device);
  if (rc == MESO_OK && ngang > 1 && universe->me == 0 && screen)
    fprintf(screen,"<MESO> %d bricks on %d listed device(s), one host thread each\n",meso_gang_size(ctx),ngang);
  if (rc != MESO_OK) {
    char msg[512];
    sprintf(msg,"<MESO> cannot create device context: %s",meso_last_error(NULL));
    error->one(FLERR,msg);
  }
  if (profile_mode.compare(0,8,"interval") == 0) {
    const char *s = profile_mode.c_str() + 8;
    profile_lo = ATOBIGINT(s);
    const char *dash = strchr(s,'-');
    profile_hi = dash ? ATOBIGINT(dash+1) : profile_lo;
  } else if (profile_mode == "all") {
    meso_profiler(ctx,1);
    profiling = true;
  }
}

MesoDevice::~MesoDevice()
{
  if (ctx) {
    if (profiling) meso_profiler(ctx,0);
    unpin_host_arrays();
    meso_destroy(ctx);
  }
}

void MesoDevice::check(int rc, const char *file, int line)
{
  if (rc == MESO_OK) return;
  char msg[1024];
  const char *txt = meso_last_error(ctx);
  sprintf(msg,"<MESO> device library error %d: %.900s",rc,txt ? txt : "(no message)");
  error->one(file,line,msg);
}

/* ----------------------------------------------------------------------
   everything the library must know that is not per-atom data
------------------------------------------------------------------------- */

void MesoDevice::push_settings()
{
  if (domain->triclinic) error->all(FLERR,"<MESO> triclinic domain not supported in USER-MESO");
  if (domain->dimension != 3) error->all(FLERR,"<MESO> USER-MESO-B200 needs a 3d simulation");
  int periodic[3] = {domain->xperiodic, domain->yperiodic, domain->zperiodic};
  check(meso_set_box(ctx,domain->boxlo,domain->boxhi,periodic),FLERR);

  if (comm->nprocs > 1) {
    // one MPI rank per GPU: the ranks share an ncclUniqueId and keep LAMMPS' brick layout
    char id[128];
    if (comm->me == 0) check(meso_comm_unique_id(id),FLERR);
    MPI_Bcast(id,128,MPI_CHAR,0,world);
    int expect = (comm->myloc[0]*comm->procgrid[1] + comm->myloc[1])*comm->procgrid[2] + comm->myloc[2];
    if (expect != comm->me)
      error->all(FLERR,"<MESO> processor mapping must be x-major (processors * * * map xyz is not supported)");
    check(meso_set_decomposition(ctx,comm->me,comm->procgrid,id),FLERR);
    check(meso_set_reduce_scope(ctx,1),FLERR);      // LAMMPS computes do their own MPI_Allreduce
  }

  if (atom->mass == NULL) error->all(FLERR,"<MESO> per-type masses are required");
  for (int i = 1; i <= atom->ntypes; i++)
    if (!atom->mass_setflag[i]) error->all(FLERR,"All masses are not set");
  check(meso_set_types(ctx,atom->ntypes,atom->mass),FLERR);
  if (atom->molecular) check(meso_set_special_bonds(ctx,force->special_lj[1]),FLERR);

  if (neighbor->delay != 0 || neighbor->dist_check != 0)
    error->all(FLERR,"<MESO> USER-MESO-B200 rebuilds on a fixed cadence: use neigh_modify delay 0 every N check no");
  check(meso_set_neighbor(ctx,neighbor->skin,neighbor->every),FLERR);
  check(meso_set_timestep_size(ctx,update->dt),FLERR);
  check(meso_set_force_units(ctx,force->ftm2v),FLERR);
  check(meso_set_ntimestep(ctx,(int64_t) update->ntimestep),FLERR);
}

/* ---------------------------------------------------------------------- */

void MesoDevice::pin_host_arrays()
{
  if (pinned_nmax == atom->nmax && npinned) return;
  unpin_host_arrays();
  const uint64_t n = (uint64_t) atom->nmax;
  void *p[7] = {atom->x ? atom->x[0] : NULL, atom->v ? atom->v[0] : NULL, atom->f ? atom->f[0] : NULL,
                atom->tag, atom->type, atom->mask, atom->image};
  const uint64_t bytes[7] = {24*n,24*n,24*n,4*n,4*n,4*n,sizeof(tagint)*n};
  for (int k = 0; k < 7; k++) {
    if (!p[k]) continue;
    check(meso_host_register(ctx,p[k],bytes[k]),FLERR);
    pinned[npinned++] = p[k];
  }
  pinned_nmax = atom->nmax;
}

void MesoDevice::unpin_host_arrays()
{
  for (int k = 0; k < npinned; k++) meso_host_unregister(ctx,pinned[k]);
  npinned = 0;
  pinned_nmax = 0;
}

void MesoDevice::upload_atoms()
{
  if (sizeof(tagint) != sizeof(int))
    error->all(FLERR,"<MESO> USER-MESO-B200 needs 32-bit image flags (default -DLAMMPS_SMALLBIG)");
  pin_host_arrays();
  const int n = atom->nlocal;
  check(meso_atoms_upload(ctx,n,n ? atom->x[0] : NULL,n ? atom->v[0] : NULL,atom->tag,atom->type,atom->mask,
                          (const int *) atom->image),FLERR);
  upload_topology(n);
  on_device = true;
}

/* per-atom bond table -> device (AtomVecDPDBond::transfer_bond, UM/atom_vec_dpd_bond_meso.cu); newton_bond must have
   been off when the data file was read, so that both atoms of a bond hold it (UM/mvv_meso.cu:96-103 forces it off) */
void MesoDevice::upload_topology(int n)
{
  bonded = atom->molecular && atom->avec->bonds_allow && atom->bond_per_atom > 0 && atom->num_bond != NULL;
  if (!bonded) return;
  const int bpa = atom->bond_per_atom, ms = atom->maxspecial;
  bigint held = 0;
  int tmax = 0;
  for (int i = 0; i < n; i++) { held += atom->num_bond[i]; if (atom->tag[i] > tmax) tmax = atom->tag[i]; }
  bigint held_all = held;
  int tmax_all = tmax;
  MPI_Allreduce(&held,&held_all,1,MPI_LMP_BIGINT,MPI_SUM,world);
  MPI_Allreduce(&tmax,&tmax_all,1,MPI_INT,MPI_MAX,world);
  if (held_all != 2*atom->nbonds)
    error->all(FLERR,"<MESO> bonds must be stored by both atoms: put `newton off` before read_data");
  topo_tag_max = tmax_all; topo_bpa = bpa; topo_maxspecial = ms;
  topo_molecule.assign(tmax_all+1,0);
  topo_num_bond.assign(tmax_all+1,0);
  topo_bond_type.assign((size_t)(tmax_all+1)*bpa,0);
  topo_bond_atom.assign((size_t)(tmax_all+1)*bpa,0);
  topo_nspecial.assign((size_t)(tmax_all+1)*3,0);
  topo_special.assign((size_t)(tmax_all+1)*(ms > 0 ? ms : 1),0);
  for (int i = 0; i < n; i++) {
    const int t = atom->tag[i];
    topo_molecule[t] = atom->molecule[i];
    topo_num_bond[t] = atom->num_bond[i];
    for (int p = 0; p < atom->num_bond[i]; p++) {
      topo_bond_type[(size_t)t*bpa+p] = atom->bond_type[i][p];
      topo_bond_atom[(size_t)t*bpa+p] = atom->bond_atom[i][p];
    }
    for (int p = 0; p < 3; p++) topo_nspecial[(size_t)t*3+p] = atom->nspecial[i][p];
    for (int p = 0; p < atom->nspecial[i][2] && p < ms; p++) topo_special[(size_t)t*ms+p] = atom->special[i][p];
  }
  // the device migrates atoms between ranks: every rank keeps the host-only topology of EVERY tag (each tag is filled by
  // exactly one rank, the others hold zeros), so atoms that arrive from elsewhere get their bonds back at download time
  if (comm->nprocs > 1) {
    std::vector<int> *tabs[6] = {&topo_molecule,&topo_num_bond,&topo_bond_type,&topo_bond_atom,&topo_nspecial,&topo_special};
    for (int k = 0; k < 6; k++) {
      std::vector<int> part(*tabs[k]);
      MPI_Allreduce(&part[0],&(*tabs[k])[0],(int) part.size(),MPI_INT,MPI_SUM,world);
    }
  }
  check(meso_bonds_upload(ctx,n,bpa,atom->num_bond,n ? atom->bond_type[0] : NULL,n ? atom->bond_atom[0] : NULL,tmax_all),FLERR);
}

void MesoDevice::restore_topology(int n)
{
  if (!bonded) return;
  const int bpa = topo_bpa, ms = topo_maxspecial;
  for (int i = 0; i < n; i++) {
    const int t = atom->tag[i];
    if (t < 0 || t > topo_tag_max) error->one(FLERR,"<MESO> atom tag outside the uploaded topology");
    atom->molecule[i] = topo_molecule[t];
    atom->num_bond[i] = topo_num_bond[t];
    for (int p = 0; p < topo_num_bond[t]; p++) {
      atom->bond_type[i][p] = topo_bond_type[(size_t)t*bpa+p];
      atom->bond_atom[i][p] = topo_bond_atom[(size_t)t*bpa+p];
    }
    for (int p = 0; p < 3; p++) atom->nspecial[i][p] = topo_nspecial[(size_t)t*3+p];
    for (int p = 0; p < atom->nspecial[i][2] && p < ms; p++) atom->special[i][p] = topo_special[(size_t)t*ms+p];
  }
}

void MesoDevice::download_atoms()
{
  if (!on_device) return;
  int nlocal = 0;
  check(meso_counts(ctx,&nlocal,NULL,NULL,NULL),FLERR);
  if (nlocal > atom->nmax) {                     // atoms migrated in (multi-rank): grow the host arrays first
    unpin_host_arrays();
    while (nlocal > atom->nmax) atom->avec->grow(0);
  }
  pin_host_arrays();
  check(meso_atoms_download(ctx,atom->nmax,atom->x[0],atom->v[0],atom->f[0],atom->tag,atom->type,atom->mask,
                            (int *) atom->image),FLERR);
  atom->nlocal = nlocal;
  atom->nghost = 0;                              // ghosts exist on the device only
  restore_topology(nlocal);
  if (atom->map_style) { atom->map_init(); atom->map_set(); }   // the device reorders atoms at every rebuild
}

/* ---------------------------------------------------------------------- */

void MesoDevice::profile_run_begin(bigint first_step)
{
  if (profile_mode == "loop" || profile_mode == "core") { meso_profiler(ctx,1); profiling = true; }
  profile_step(first_step);
}

void MesoDevice::profile_step(bigint step)
{
  if (profile_mode.compare(0,8,"interval") != 0) return;
  if (!profiling && step >= profile_lo && step < profile_hi) { meso_profiler(ctx,1); profiling = true; }
  else if (profiling && step >= profile_hi) { meso_profiler(ctx,0); profiling = false; }
}

void MesoDevice::profile_run_end()
{
  if (profiling && profile_mode != "all") { meso_profiler(ctx,0); profiling = false; }
}

/* ---------------------------------------------------------------------- */

void MesoDevice::timers_begin()
{
  double ms[MESO_T_COUNT];
  int64_t calls[MESO_T_COUNT];
  meso_timers_enable(ctx,1);
  meso_timers_read(ctx,ms,calls,1);
}

void MesoDevice::timers_end()
{
  double ms[MESO_T_COUNT];
  int64_t calls[MESO_T_COUNT];
  if (meso_timers_read(ctx,ms,calls,1) != MESO_OK) return;
  meso_timers_enable(ctx,0);
  // Finish prints TIME_LOOP minus the categories as "Other": the integrator kernels stay there
  timer->array[TIME_PAIR] += 1.0e-3*ms[MESO_T_PAIR];
  timer->array[TIME_NEIGHBOR] += 1.0e-3*(ms[MESO_T_NEIGH] + ms[MESO_T_REBUILD]);
  timer->array[TIME_COMM] += 1.0e-3*ms[MESO_T_FORWARD];
  neighbor->ncalls += (int) calls[MESO_T_NEIGH];        // "Neighbor list builds" of Finish
}

/* ---------------------------------------------------------------------- */

MesoDevice *MesoBridge::mdev(const char *who) const
{
  if (mlmp->meso_device == NULL) {
    char msg[256];
    sprintf(msg,"<MESO> %s needs the device runtime: do not combine it with -meso off",who);
    mlmp->error->all(FLERR,msg);
  }
  return mlmp->meso_device;
}

meso_ctx *MesoBridge::mctx(const char *who) const { return mdev(who)->ctx; }

void MesoBridge::mcheck(int rc, const char *file, int line) const
{
  if (rc != MESO_OK) mlmp->meso_device->check(rc,file,line);
}
