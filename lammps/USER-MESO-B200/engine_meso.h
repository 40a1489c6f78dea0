/* engine_meso.h -- MesoDevice, the object src/lammps.cpp:455-465 constructs with -meso on.
   Reference: UM/engine_meso.h:29-338 (streams, events, tagged allocation table, profiler gate).
   Here it is a thin owner of one meso_ctx (the C ABI of include/meso_b200.h): the library owns
   every device byte, stream and event; this class only moves LAMMPS settings and host arrays
   across the boundary. */
#ifndef LMP_MESO_ENGINE
#define LMP_MESO_ENGINE

#include <string>
#include <vector>
#include "pointers.h"
#include "meso_b200.h"

namespace LAMMPS_NS {

class MesoDevice : protected Pointers {
 public:
  const bool dummy;        // read by src/lammps.cpp:456,460
  meso_ctx *ctx;

  MesoDevice(class LAMMPS *, int device, std::string profile);
  ~MesoDevice();

  void check(int rc, const char *file, int line);

  // settings: Domain box + periodicity, masses, neighbor skin/every, dt, units, ntimestep
  void push_settings();
  // transfer_pre_exchange-like: all local atoms host AoS -> device SoA (UM/atom_meso.cu:185-199)
  void upload_atoms();
  // transfer_pre_output: device -> host x,v,f,tag,type,mask,image in device order (UM/atom_meso.cu:258-266)
  void download_atoms();
  bool resident() const { return on_device; }   // device holds the current atoms
  void invalidate() { on_device = false; }

  // -profile all|loop|core|intervalA-B (src/lammps.cpp:187-191; UM/engine_meso.cu:155-191)
  void profile_run_begin(bigint first_step);
  void profile_step(bigint step);
  void profile_run_end();

  // device-event phase times of the last run, folded into LAMMPS' timer categories
  void timers_begin();
  void timers_end();

 private:
  std::string profile_mode;
  bigint profile_lo, profile_hi;
  bool profiling, on_device;
  void *pinned[8];
  int npinned, pinned_nmax;
  void pin_host_arrays();
  void unpin_host_arrays();
  // bead-spring decks: host-only per-atom arrays (molecule, bond and special tables) keyed by tag at upload time,
  // so that they can follow the device's reorder when the atoms come back
  bool bonded;
  int topo_tag_max, topo_bpa, topo_maxspecial;
  std::vector<int> topo_molecule, topo_num_bond, topo_bond_type, topo_bond_atom, topo_nspecial, topo_special;
  void upload_topology(int n);
  void restore_topology(int n);
};

}

#endif
