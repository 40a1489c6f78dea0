"""Generates tests/golden/stock_dpd_equilibrium_L10.npz with the UNMODIFIED stock LAMMPS of the reference tree:
equilibrium observables of the example/simple deck's fluid (rho = 4, a = 15, gamma = 4.5, kT = 1) for the statistical
parity check of BASELINE.json's north_star (temperature, pressure, g(r)).

Dev container only (needs oracle/_ref/lmp_serial):  python tests/golden/make_lammps_rdf_golden.py
Stock pair_style dpd + fix nve on workload.dpd_fluid(10); 1000 steps of equilibration, then 2000 steps with
compute rdf 50 (src/compute_rdf.cpp) averaged by fix ave/time every 10 steps, and thermo temp/press every 10 steps.
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
L, NBIN = 10, 50
with tempfile.TemporaryDirectory() as d:
    workload.write_data(os.path.join(d, "c.data"), workload.dpd_fluid(L), L)
    deck = """dimension 3
units lj
atom_style atomic
communicate single vel yes
neighbor 0.3 bin
neigh_modify delay 0 every 5 check no
read_data c.data
pair_style dpd 1.0 1.0 419084618
pair_coeff 1 1 15 4.5 1.0
velocity all create 1.0 788662042 loop all
fix 3 all nve
timestep 0.005
thermo 100
run 1000
reset_timestep 0
compute g all rdf %d
fix avg all ave/time 10 200 2000 c_g file rdf.txt mode vector
thermo_style custom step temp press
thermo 10
run 2000
""" % NBIN
    open(os.path.join(d, "in.rdf"), "w").write(deck)
    r = subprocess.run([LMP, "-meso", "off", "-in", "in.rdf", "-log", "none"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [s.split() for s in open(os.path.join(d, "rdf.txt")) if not s.startswith("#")]
    body = np.array([[float(t) for t in row] for row in rows if len(row) == 4])      # bin r g(r) coord(r)
    assert body.shape[0] == NBIN, body.shape
    th, on = [], False
    for s in r.stdout.split("\n"):
        t = s.split()
        if s.startswith("Step Temp Press"):
            on = True
        elif s.startswith("Loop time"):
            on = False
        elif on and len(t) == 3:
            th.append([float(v) for v in t])
    th = np.array(th)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "stock_dpd_equilibrium_L10.npz"), L=L, nbin=NBIN, r=body[:, 1], g=body[:, 2],
                    temp=th[:, 1], press=th[:, 2],
                    note="stock LAMMPS 30Sep2013 pair_style dpd 1.0 1.0 seed, pair_coeff 1 1 15 4.5 1.0, fix nve, dt 0.005; steps 1000-3000 of "
                         "workload.dpd_fluid(10); compute rdf 50 averaged every 10 steps; thermo every 10 steps")
print("T %.4f +- %.4f  P %.3f +- %.3f  g(max) %.3f at r %.2f" % (th[:, 1].mean(), th[:, 1].std(), th[:, 2].mean(), th[:, 2].std(),
                                                             body[:, 2].max(), body[np.argmax(body[:, 2]), 1]))
