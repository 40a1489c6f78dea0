/* fix_styles_meso.h -- every fix style of the package in one header.

     nve/meso          FixNVEMeso          UM/fix_nve_meso.{h,cu}            velocity-Verlet halves on the device SoA store
     wall/meso         MesoFixWall         UM/fix_wall_meso.{h,cu}           erfc wall force + bounce-forward at the box faces
     solid_bound/meso  MesoFixSolidBound   UM/fix_solid_bound_meso.{h,cu}    polynomial wall force (rho5rc1s1) + bounce-forward
     addforce/meso     MesoFixAddForce     UM/fix_addforce_meso.{h,cu}       constant body force
     pois/meso         MesoFixPoiseuille   UM/fix_poiseuille_meso.{h,cu}     counter-flowing body force across a bisection plane
     rdf/fast/meso     MesoFixRDFFast      UM/fix_rdf_fast_meso.{h,cu}       g(r) up to the pair cutoff from the device neighbor table

   All but nve/meso derive from MesoFixResident: their whole effect lives in the library's device-resident fix list.  A fix
   registers itself with the context in init() (ModifiedVerlet::init() has cleared the list just before: LAMMPS::init runs
   Update::init ahead of Modify::init) and keeps the handle.  With only such fixes next to one nve/meso the fused run loop
   (meso_run) applies them itself; otherwise Modify calls the phase hooks below once per step. */
#ifdef FIX_CLASS

FixStyle(nve/meso,FixNVEMeso)
FixStyle(wall/meso,MesoFixWall)
FixStyle(solid_bound/meso,MesoFixSolidBound)
FixStyle(addforce/meso,MesoFixAddForce)
FixStyle(pois/meso,MesoFixPoiseuille)
FixStyle(rdf/fast/meso,MesoFixRDFFast)

#else
#ifndef LMP_MESO_B200_FIX_STYLES_H
#define LMP_MESO_B200_FIX_STYLES_H

#include <string>
#include <vector>
#include "fix.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

/* dtf = 0.5*dt*ftm2v, dtv = dt, group-mask test.  ModifiedVerlet fuses both halves into neighbouring kernels when this is
   the only integrating fix. */
class FixNVEMeso : public Fix, protected MesoBridge {
 public:
  FixNVEMeso(class LAMMPS *, int, char **);
  int setmask();
  void init();
  void initial_integrate(int);
  void final_integrate();
  void reset_dt();
  int group_bit() const { return groupbit; }
};

class MesoFixResident : public Fix, protected MesoBridge {
 public:
  MesoFixResident(class LAMMPS *lmp, int narg, char **arg) : Fix(lmp,narg,arg), MesoBridge(lmp), handle(-1) {}
  virtual void init();
  virtual void setup(int) {}            // Fix::setup -> post_force of the reference: meso_setup has applied the whole list
  virtual void post_force(int);
 protected:
  int handle;
  virtual int register_fix(meso_ctx *) = 0;
  void bounce();
};

class MesoFixWall : public MesoFixResident {
 public:
  MesoFixWall(class LAMMPS *, int, char **);
  int setmask();
  void pre_exchange() { bounce(); }
  void end_of_step() { bounce(); }
 protected:
  bool x, y, z;
  double d, f;
  int register_fix(meso_ctx *);
};

class MesoFixSolidBound : public MesoFixResident {
 public:
  MesoFixSolidBound(class LAMMPS *, int, char **);
  int setmask();
  void pre_exchange() { bounce(); }
  void end_of_step() { bounce(); }
 protected:
  bool x, y, z;
  int force_kernel;                      // 0 unspecified, 1 rho5rc1s1
  int register_fix(meso_ctx *);
};

class MesoFixAddForce : public MesoFixResident {
 public:
  MesoFixAddForce(class LAMMPS *, int, char **);
  int setmask();
 protected:
  double fx, fy, fz;
  int register_fix(meso_ctx *);
};

class MesoFixPoiseuille : public MesoFixResident {
 public:
  MesoFixPoiseuille(class LAMMPS *, int, char **);
  int setmask();
 protected:
  int dim_ortho, dim_force;
  double strength, bisect_frac;
  int register_fix(meso_ctx *);
};

/* sampled in the post_force slot; harvested from the device at post_run; g(r) is written when the fix is destroyed */
class MesoFixRDFFast : public MesoFixResident {
 public:
  MesoFixRDFFast(class LAMMPS *, int, char **);
  ~MesoFixRDFFast();
  int setmask();
  void init();
  void post_run();
 protected:
  std::string output;
  int n_bin, n_every, j_groupbit;
  double rc, n_i, n_j, n_steps;
  std::vector<double> total, last;       // pair counts per bin: harvested from the device / at the previous harvest
  double last_steps;
  int register_fix(meso_ctx *);
  void harvest();
  void dump();
};

}

#endif
#endif
