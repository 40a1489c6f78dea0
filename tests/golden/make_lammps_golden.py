"""Generates tests/golden/stock_dpd_conservative_L8.npz with the UNMODIFIED stock LAMMPS of the reference tree.

Dev container only (needs oracle/_ref/lmp_serial, built by `make -C oracle ref_lammps`):
    python tests/golden/make_lammps_golden.py
Stock pair_style dpd (src/pair_dpd.cpp:91-155) with gamma = 0 (hence sigma = 0) on the synthetic 8^3, rho = 4
fluid of meso_b200.workload.dpd_fluid(8): conservative forces in fp64 on fp64 coordinates, `run 0`.
Also records a short full-DPD run of the 25^3-like deck for the statistical checks (T, P).
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
L = 8
x = workload.dpd_fluid(L)
with tempfile.TemporaryDirectory() as d:
    workload.write_data(os.path.join(d, "8.data"), x, L)
    deck = """dimension 3
units lj
atom_style atomic
communicate single vel yes
neighbor 0.3 bin
neigh_modify delay 0 every 5 check no
read_data 8.data
pair_style dpd 1.0 1.0 419084618
pair_coeff 1 1 15 0.0 1.0
fix 3 all nve
thermo_style custom step temp press pe
dump d all custom 1 f.dump id fx fy fz
dump_modify d format "%d %.17g %.17g %.17g" sort id
timestep 0.005
run 0
"""
    open(os.path.join(d, "in.cons"), "w").write(deck)
    out = subprocess.run([LMP, "-meso", "off", "-in", "in.cons", "-log", "none"], cwd=d, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = open(os.path.join(d, "f.dump")).read().split("\n")
    i0 = [i for i, s in enumerate(lines) if s.startswith("ITEM: ATOMS")][0] + 1
    body = np.array([[float(t) for t in s.split()] for s in lines[i0:i0 + len(x)]])
    assert np.array_equal(body[:, 0], np.arange(1, len(x) + 1))
    f = body[:, 1:4]
    thermo = [s for s in out.stdout.split("\n") if s.strip().startswith("0 ")]
    press, pe = float(thermo[0].split()[2]), float(thermo[0].split()[3])
np.savez_compressed(os.path.join(os.path.dirname(__file__), "stock_dpd_conservative_L8.npz"), f=f, L=L, press=press, pe=pe,
                    note="stock LAMMPS 30Sep2013 pair_style dpd 1.0 1.0 seed; pair_coeff 1 1 15 0.0 1.0; run 0; workload.dpd_fluid(8)")
print("forces", f.shape, "rms", np.sqrt((f * f).sum(1).mean()), "press", press, "pe/atom", pe)
