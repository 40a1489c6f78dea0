"""Generates tests/golden/stock_polymer_conservative_L6.npz with the UNMODIFIED stock LAMMPS of the reference tree
(atom_style bond + bond_style harmonic of src/MOLECULE, pair_style dpd with gamma = 0).

Dev container only (needs oracle/_ref/lmp_serial, built by `make -C oracle ref_lammps`):
    python tests/golden/make_lammps_bond_golden.py
Bead-spring chains in solvent from meso_b200.workload.polymer_melt(6, chain_len=8, seed=5); `run 0` once with
special_bonds lj 1 1 1 (1-2 pairs keep their pair force) and once with lj 0 1 1 (1-2 pairs excluded): forces by atom id,
pair energy, bond energy (src/MOLECULE/bond_harmonic.cpp:46-112: E = K (r - r0)^2).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from meso_b200 import workload  # noqa: E402

LMP = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
L, CHAIN, SEED, K, R0 = 6, 8, 5, 50.0, 0.5
x, typ, tag, nb, bt, ba = workload.polymer_melt(L, chain_len=CHAIN, seed=SEED)
out = {}
with tempfile.TemporaryDirectory() as d:
    nbonds = workload.write_data_bond(os.path.join(d, "p.data"), x, L, typ, 2, nb, bt, ba)
    for lj12 in (1, 0):
        deck = """dimension 3
units lj
newton off
atom_style bond
communicate single vel yes
neighbor 0.3 bin
neigh_modify delay 0 every 5 check no
read_data p.data
bond_style harmonic
bond_coeff 1 %g %g
special_bonds lj %d 1 1
pair_style dpd 1.0 1.0 419084618
pair_coeff 1 1 25 0.0 1.0
pair_coeff 1 2 40 0.0 1.0
pair_coeff 2 2 25 0.0 1.0
fix 3 all nve
thermo_style custom step temp evdwl ebond press
thermo_modify norm no
dump d all custom 1 f.dump id fx fy fz
dump_modify d format "%%d %%.17g %%.17g %%.17g" sort id
timestep 0.005
run 0
""" % (K, R0, lj12)
        open(os.path.join(d, "in.p"), "w").write(deck)
        r = subprocess.run([LMP, "-meso", "off", "-in", "in.p", "-log", "none"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        lines = open(os.path.join(d, "f.dump")).read().split("\n")
        i0 = [i for i, s in enumerate(lines) if s.startswith("ITEM: ATOMS")][0] + 1
        body = np.array([[float(t) for t in s.split()] for s in lines[i0:i0 + len(x)]])
        assert np.array_equal(body[:, 0], np.arange(1, len(x) + 1))
        th = [s for s in r.stdout.split("\n") if s.strip().startswith("0 ")][0].split()
        out["f_lj%d" % lj12] = body[:, 1:4]
        out["evdwl_lj%d" % lj12], out["ebond_lj%d" % lj12], out["press_lj%d" % lj12] = float(th[2]), float(th[3]), float(th[4])
np.savez_compressed(os.path.join(os.path.dirname(__file__), "stock_polymer_conservative_L6.npz"), L=L, chain_len=CHAIN, seed=SEED, k=K, r0=R0,
                    nbonds=nbonds, note="stock LAMMPS 30Sep2013 atom_style bond, bond_style harmonic, pair_style dpd gamma=0, run 0; "
                                        "workload.polymer_melt(6, chain_len=8, seed=5)", **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()}, "bonds", nbonds)
