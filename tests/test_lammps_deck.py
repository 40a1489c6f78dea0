"""The reference's benchmark decks through the LAMMPS plugin surface (lammps/USER-MESO-B200 -> C ABI -> CUDA).

lammps/_build/lmp_meso_b200 is the reference's own LAMMPS core (30 Sep 2013) compiled with this repository's
package instead of the reference's USER-MESO (lammps/Makefile; built in the dev container, shipped to the GPU box).
The decks below are example/simple/sp.run / dp.run command for command (atom_style dpd/atomic/meso, run_style
mvv/meso, pair_style dpd/fast/meso | dpd/meso, compute temp/meso, fix nve/meso), plus a dump so that the
trajectory can be compared with the same library driven through the Python mirror of the C ABI.
"""
import os
import subprocess

import numpy as np
import pytest

from meso_b200 import workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LMP = os.path.join(ROOT, "lammps", "_build", "lmp_meso_b200")

DECK = """# example/simple/{prec}.run, box edge ${{case}}
dimension       3
units           lj
atom_style      dpd/atomic/meso
neighbor        0.3 bin
neigh_modify    delay 0 every 5 check no
read_data       ${{case}}.data
run_style       mvv/meso
pair_style      {pair} 1.0 419084618
pair_coeff      1 1 15 4.5 3.0 1.0 1.0
compute         mythermo all temp/meso
velocity        all create 1.0 788662042 loop all
fix             3 all nve/meso
thermo_style    custom step temp {extra} cpu spcpu
thermo          {thermo}
thermo_modify   temp mythermo
{dump}
timestep        0.005
run             {steps}
"""
DUMP = ("dump            d all custom {steps} traj.txt id x y z vx vy vz fx fy fz\n"
        "dump_modify     d format \"%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\"")


def need_binary():
    if not os.path.exists(LMP):
        pytest.fail("lammps/_build/lmp_meso_b200 is missing: build it in the dev container (make -C lammps); it ships with the snapshot")


def run_deck(tmp_path, L, prec, steps, thermo, extra="", dump=True, args=(), env=None):
    workload.write_data(str(tmp_path / ("%d.data" % L)), workload.dpd_fluid(L), L)
    deck = DECK.format(prec=prec, pair="dpd/fast/meso" if prec == "sp" else "dpd/meso", extra=extra, thermo=thermo, steps=steps,
                       dump=DUMP.format(steps=steps) if dump else "")
    (tmp_path / "in.run").write_text(deck)
    out = subprocess.run([LMP, "-in", "in.run", "-var", "case", str(L), "-log", "none"] + list(args), cwd=str(tmp_path),
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    return out


def frames(path):
    """{step: array sorted by id, columns id x y z vx vy vz fx fy fz}"""
    out, lines = {}, open(path).read().split("\n")
    i = 0
    while i < len(lines):
        if lines[i].startswith("ITEM: TIMESTEP"):
            step, n = int(lines[i + 1]), int(lines[i + 3])
            body = np.array([[float(t) for t in s.split()] for s in lines[i + 9:i + 9 + n]])
            out[step] = body[np.argsort(body[:, 0])]
            i += 9 + n
        else:
            i += 1
    return out


def thermo_rows(stdout, ncol):
    rows, on = [], False
    for s in stdout.split("\n"):
        if s.startswith("Step "):
            on = True
            continue
        if s.startswith("Loop time"):
            on = False
        t = s.split()
        if on and len(t) == ncol:
            try:
                rows.append([float(v) for v in t])
            except ValueError:
                pass
    return np.array(rows)


def mirror(L, prec, fr0):
    from meso_b200.engine import Meso
    m = Meso(0)
    m.box((0.0, 0.0, 0.0), (L, L, L))
    m.masses([0.0, 1.0])
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/fast/meso" if prec == "sp" else "dpd/meso", 1.0, 419084618)
    m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)
    m.upload(np.ascontiguousarray(fr0[:, 1:4]), np.ascontiguousarray(fr0[:, 4:7]), tag=fr0[:, 0].astype(np.int32))
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["sp", "dp"])
def test_benchmark_deck_runs_unchanged_and_matches_the_c_abi_mirror(tmp_path, prec, monkeypatch):
    """bit-for-bit with the deterministic two-sided force kernel (MESO_PAIR_ONCE=0) in both processes; the default
    pair-once loop reduces with atomics (summation order varies run to run) and is compared on observables below"""
    need_binary()
    L, steps = 10, 20
    monkeypatch.setenv("MESO_PAIR_ONCE", "0")
    out = run_deck(tmp_path, L, prec, steps, thermo=10, env={"MESO_PAIR_ONCE": "0"})
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Loop time of" in out.stdout
    th = thermo_rows(out.stdout, 4)
    assert th.shape[0] == 3 and abs(th[0, 1] - 1.0) < 1e-12          # velocity create 1.0, read back through temp/meso
    fr = frames(str(tmp_path / "traj.txt"))
    assert sorted(fr) == [0, steps]
    m = mirror(L, prec, fr[0])
    m.setup()
    a = m.download()
    o = np.argsort(a["tag"])
    assert np.array_equal(a["f"][o], fr[0][:, 7:10]), "setup forces differ between the LAMMPS-driven and the mirror-driven library"
    m.run(steps)
    a = m.download()
    o = np.argsort(a["tag"])
    assert np.array_equal(a["x"][o], fr[steps][:, 1:4])
    assert np.array_equal(a["v"][o], fr[steps][:, 4:7])
    assert np.array_equal(a["f"][o], fr[steps][:, 7:10])
    assert abs(m.temperature() - th[-1, 1]) < 1e-6                      # thermo prints 8 significant digits
    m.close()
    # default loop (each pair once): same deck, same thermo trace to the fp tolerance of the path
    d2 = tmp_path / "once"
    d2.mkdir()
    out2 = run_deck(d2, L, prec, steps, thermo=10, env={"MESO_PAIR_ONCE": "1"})
    assert out2.returncode == 0, out2.stdout[-2000:] + out2.stderr[-2000:]
    th2 = thermo_rows(out2.stdout, 4)
    # fp32: a few atoms per step draw different random numbers (RNG keys on fp32 velocity bits): T of 4000 atoms agrees statistically
    assert np.abs(th2[:, 1] - th[:, 1]).max() < (6e-2 if prec == "sp" else 1e-7), (th2[:, 1], th[:, 1])
    fr2 = frames(str(d2 / "traj.txt"))
    if prec == "dp":
        assert np.abs(fr2[steps][:, 1:7] - fr[steps][:, 1:7]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("once", ["0", "1"])
def test_thermo_energy_and_pressure_go_through_the_phase_entry_points(tmp_path, monkeypatch, once):
    """`pe` and `press` make LAMMPS ask for energy/virial on thermo steps: those steps run phase by phase
    (fix nve/meso -> meso_initial_integrate, pair->compute(eflag,vflag), ...), the others through meso_run."""
    need_binary()
    # once = 0: deterministic two-sided kernel in meso_run, bit-for-bit; once = 1 (default loop): fp64 atomics, 1e-9
    L, steps = 10, 20
    monkeypatch.setenv("MESO_PAIR_ONCE", once)
    out = run_deck(tmp_path, L, "dp", steps, thermo=10, extra="pe press", env={"MESO_PAIR_ONCE": once})
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    th = thermo_rows(out.stdout, 6)
    assert th.shape[0] == 3
    fr = frames(str(tmp_path / "traj.txt"))
    m = mirror(L, "dp", fr[0])
    m.setup(eflag=1, vflag=1)
    n = len(fr[0])
    vol = float(L) ** 3

    def pe_press():
        vir, e = m.virial()
        t = m.temperature()
        return e / n, ((3 * n - 3) * t + vir[0] + vir[1] + vir[2]) / (3.0 * vol)   # thermo normalises pe by N (lj units)

    pe0, p0 = pe_press()
    assert abs(pe0 - th[0, 2]) < 2e-7 * abs(pe0) and abs(p0 - th[0, 3]) < 2e-7 * abs(p0), (pe0, p0, th[0])
    m.run(steps - 1)
    # last step with tallies, phase by phase
    m.ntimestep = steps
    m.initial_integrate()
    if m.neighbor_decide():
        m.rebuild()
    else:
        m.forward_comm()
    m.force_clear(vflag=1)
    m.pair_compute(eflag=1, vflag=1)
    m.final_integrate()
    pe1, p1 = pe_press()
    assert abs(pe1 - th[-1, 2]) < 2e-7 * abs(pe1) and abs(p1 - th[-1, 3]) < 2e-7 * abs(p1), (pe1, p1, th[-1])
    a = m.download()
    o = np.argsort(a["tag"])
    if once == "0":
        assert np.array_equal(a["x"][o], fr[steps][:, 1:4]) and np.array_equal(a["v"][o], fr[steps][:, 4:7])
    else:
        assert np.abs(a["x"][o] - fr[steps][:, 1:4]).max() < 1e-9 and np.abs(a["v"][o] - fr[steps][:, 4:7]).max() < 1e-9
    # equilibrium DPD fluid at rho = 4, a = 15... just started from random positions: pressure is positive and O(10)
    assert 5.0 < p1 < 80.0
    m.close()


CHANNEL_DECK = """# driven channel: walls across z, counter-flowing body force (SURVEY.md s8f N2)
dimension       3
units           lj
boundary        p p f
atom_style      dpd/atomic/meso
neighbor        0.3 bin
neigh_modify    delay 0 every 5 check no
read_data       {L}.data
run_style       mvv/meso
pair_style      dpd/meso 1.0 419084618
pair_coeff      1 1 15 4.5 3.0 1.0 1.0
compute         mythermo all temp/meso
velocity        all create 1.0 788662042 loop all
fix             3 all nve/meso
fix             4 all wall/meso z d 0.5 f 20.0
fix             5 all pois/meso z x 0.3
fix             6 all addforce/meso 0.0 0.1 0.0
fix             7 all rdf/fast/meso output rdf.txt nbin 40 every 5
thermo_style    custom step temp {extra} cpu
thermo          10
thermo_modify   temp mythermo
{dump}
timestep        0.005
run             {steps}
"""


@pytest.mark.gpu
def test_channel_deck_with_device_resident_fixes(tmp_path, monkeypatch):
    """wall/meso + pois/meso + addforce/meso next to nve/meso: the fused loop applies the library's fix list (bit-for-bit
    with the C-ABI mirror given the same list); with `pe press` in the thermo line the thermo steps go through the fixes'
    own post_force / pre_exchange / end_of_step hooks and land on the same trajectory."""
    need_binary()
    monkeypatch.setenv("MESO_PAIR_ONCE", "0")        # deterministic two-sided kernel in both processes (read at meso_create)
    L, steps = 8, 20
    workload.write_data(str(tmp_path / ("%d.data" % L)), workload.dpd_fluid(L), L)

    def run(extra, sub):
        d = tmp_path / sub
        d.mkdir()
        os.symlink(str(tmp_path / ("%d.data" % L)), str(d / ("%d.data" % L)))
        (d / "in.run").write_text(CHANNEL_DECK.format(L=L, extra=extra, steps=steps, dump=DUMP.format(steps=steps)))
        out = subprocess.run([LMP, "-in", "in.run", "-log", "none"], cwd=str(d), capture_output=True, text=True, timeout=600,
                             env=dict(os.environ, MESO_PAIR_ONCE="0"))
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return frames(str(d / "traj.txt")), out

    fr, out = run("", "fused")
    from meso_b200.engine import Meso
    m = Meso(0)
    m.box((0.0, 0.0, 0.0), (L, L, L), (1, 1, 0))
    m.masses([0.0, 1.0])
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/meso", 1.0, 419084618)
    m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)
    m.upload(np.ascontiguousarray(fr[0][:, 1:4]), np.ascontiguousarray(fr[0][:, 4:7]), tag=fr[0][:, 0].astype(np.int32))
    m.fix("wall/meso", "z", "d", 0.5, "f", 20.0)
    m.fix("pois/meso", "z", "x", 0.3)
    m.fix("addforce/meso", 0.0, 0.1, 0.0)
    hr = m.fix("rdf/fast/meso", "output", str(tmp_path / "rdf_mirror.txt"), "nbin", 40, "every", 5)
    m.setup()
    a = m.download()
    o = np.argsort(a["tag"])
    assert np.array_equal(a["f"][o], fr[0][:, 7:10]), "setup forces (pair + fix post_force) differ"
    m.run(steps)
    a = m.download()
    o = np.argsort(a["tag"])
    assert np.array_equal(a["x"][o], fr[steps][:, 1:4]) and np.array_equal(a["v"][o], fr[steps][:, 4:7])
    assert np.array_equal(a["f"][o], fr[steps][:, 7:10])
    assert a["x"][:, 2].min() > 0 and a["x"][:, 2].max() < L
    # rdf/fast/meso: g(r) written when LAMMPS destroys the fix == the mirror's dump of the same four samples (steps 5..20)
    m.rdf_dump(hr, float(L) ** 3)
    assert m.rdf(hr, float(L) ** 3)[3] == 4
    g_lmp, g_mir = np.loadtxt(str(tmp_path / "fused" / "rdf.txt")), np.loadtxt(str(tmp_path / "rdf_mirror.txt"))
    assert g_lmp.shape == (40, 2) and np.allclose(g_lmp, g_mir, rtol=1e-12, atol=0) and g_lmp[:, 1].max() > 0.5
    m.close()
    fr2, out2 = run("pe press", "phases")
    assert np.allclose(np.loadtxt(str(tmp_path / "phases" / "rdf.txt")), g_lmp, rtol=1e-12)   # sampled through Fix::post_force here
    assert np.abs(fr2[steps][:, 1:7] - fr[steps][:, 1:7]).max() < 1e-9


POLYMER_DECK = """# bead-spring chains in solvent (SURVEY.md s8f N1)
dimension       3
units           lj
newton          off
atom_style      dpd/bond/meso
neighbor        0.3 bin
neigh_modify    delay 0 every 5 check no
read_data       p.data
run_style       mvv/meso
bond_style      harmonic/meso
bond_coeff      1 50.0 0.5
special_bonds   lj 0 1 1
pair_style      dpd/meso 1.0 419084618
pair_coeff      1 1 25 4.5 3.0 1.0 1.0
pair_coeff      1 2 40 4.5 3.0 1.0 1.0
pair_coeff      2 2 25 4.5 3.0 1.0 1.0
compute         mythermo all temp/meso
velocity        all create 1.0 788662042 loop all
fix             3 all nve/meso
thermo_style    custom step temp {extra} cpu
thermo          10
thermo_modify   temp mythermo norm no
{dump}
timestep        0.005
run             {steps}
"""


@pytest.mark.gpu
def test_polymer_deck_dpd_bond_meso_and_harmonic_meso(tmp_path, monkeypatch):
    """atom_style dpd/bond/meso + bond_style harmonic/meso + special_bonds lj 0 through lmp_meso_b200: the Bonds section
    reaches the device table, the fused loop adds the bonded forces, the host's bond arrays follow the device reorder
    (the dump after 20 steps is by id), thermo's ebond / pe come from the device reductions."""
    need_binary()
    monkeypatch.setenv("MESO_PAIR_ONCE", "0")
    L, steps = 8, 20
    x, typ, tag, nb, bt, ba = workload.polymer_melt(L, chain_len=8, seed=5)
    workload.write_data_bond(str(tmp_path / "p.data"), x, L, typ, 2, nb, bt, ba)

    def run(extra, sub, ncol):
        d = tmp_path / sub
        d.mkdir()
        os.symlink(str(tmp_path / "p.data"), str(d / "p.data"))
        (d / "in.run").write_text(POLYMER_DECK.format(extra=extra, steps=steps, dump=DUMP.format(steps=steps)))
        out = subprocess.run([LMP, "-in", "in.run", "-log", "none"], cwd=str(d), capture_output=True, text=True, timeout=600,
                             env=dict(os.environ, MESO_PAIR_ONCE="0"))
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return frames(str(d / "traj.txt")), thermo_rows(out.stdout, ncol)

    fr, th = run("", "fused", 3)
    from meso_b200.engine import Meso
    m = Meso(0)
    m.box((0.0, 0.0, 0.0), (L, L, L))
    m.masses([0.0, 1.0, 1.0])
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/meso", 1.0, 419084618)
    m.pair_coeff(1, 1, 25, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(1, 2, 40, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(2, 2, 25, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)
    assert np.array_equal(fr[0][:, 0].astype(np.int32), tag)
    m.upload(np.ascontiguousarray(fr[0][:, 1:4]), np.ascontiguousarray(fr[0][:, 4:7]), tag=tag, type=typ)
    m.bond_style("harmonic/meso", 1)
    m.bond_coeff(1, 50.0, 0.5)
    m.special_bonds(0.0)
    m.bonds(nb, bt, ba)
    m.setup(eflag=1, vflag=1)
    a = m.download()
    o = np.argsort(a["tag"])
    assert np.array_equal(a["f"][o], fr[0][:, 7:10]), "setup forces (pair + bonds) differ"
    eb0 = m.bond_energy()
    _, ep0 = m.virial()
    m.run(steps)
    a = m.download()
    o = np.argsort(a["tag"])
    assert np.array_equal(a["x"][o], fr[steps][:, 1:4]) and np.array_equal(a["v"][o], fr[steps][:, 4:7])
    assert np.array_equal(a["f"][o], fr[steps][:, 7:10])
    m.close()
    # thermo steps by phases, energies from the device reductions; the by-id dump proves the host bond arrays stayed aligned
    fr2, th2 = run("ebond pe", "phases", 5)
    assert np.abs(fr2[steps][:, 1:7] - fr[steps][:, 1:7]).max() < 1e-9
    assert abs(th2[0, 2] - eb0) < 2e-7 * eb0 and abs(th2[0, 3] - (eb0 + ep0)) < 2e-7 * (eb0 + ep0), (th2[0], eb0, ep0)
    assert 0 < th2[-1, 2] < eb0                      # the stretched start (bond length 0.7, r0 0.5) relaxes


@pytest.mark.gpu
def test_restart_round_trip_continues_the_trajectory(tmp_path):
    """SURVEY.md s8f N4 (I/O edges): write_restart after 10 steps hands the device atoms back through transfer_pre_output and
    stores the pair style's restart records (settings {cut_global, seed, mix_flag} + per-pair {a0, gamma, sigma, expw, cut},
    UM/pair_dpd_meso.cu:363-445); a second process reads the file, needs no pair_style / pair_coeff lines, and lands on the
    same step-20 state as the uninterrupted deck (both re-neighbour at step 10: bit-for-bit with the two-sided kernel)."""
    need_binary()
    L = 8
    workload.write_data(str(tmp_path / ("%d.data" % L)), workload.dpd_fluid(L), L)
    env = dict(os.environ, MESO_PAIR_ONCE="0")
    head = ("dimension 3\nunits lj\natom_style dpd/atomic/meso\nneighbor 0.3 bin\nneigh_modify delay 0 every 5 check no\n")
    tail = ("run_style mvv/meso\ncompute mythermo all temp/meso\nfix 3 all nve/meso\nthermo 10\nthermo_modify temp mythermo\ntimestep 0.005\n")
    dump = "dump d all custom 10 %s id x y z vx vy vz fx fy fz\ndump_modify d format \"%%d %%.17g %%.17g %%.17g %%.17g %%.17g %%.17g %%.17g %%.17g %%.17g\"\n"
    (tmp_path / "in.a").write_text(head + "read_data %d.data\npair_style dpd/meso 1.0 419084618\npair_coeff 1 1 15 4.5 3.0 1.0 1.0\n" % L +
                                   "velocity all create 1.0 788662042 loop all\n" + tail + "run 10\nwrite_restart r.bin\n" + dump % "a.txt" + "run 10\n")
    (tmp_path / "in.b").write_text(head + "read_restart r.bin\n" + tail + dump % "b.txt" + "run 10\n")
    for deck in ("in.a", "in.b"):
        out = subprocess.run([LMP, "-in", deck, "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True, timeout=600, env=env)
        assert out.returncode == 0, deck + out.stdout[-2000:] + out.stderr[-2000:]
    fa, fb = frames(str(tmp_path / "a.txt")), frames(str(tmp_path / "b.txt"))
    assert 20 in fa and 20 in fb
    assert np.array_equal(fa[10][:, 1:7], fb[10][:, 1:7])              # the restart file carries x, v exactly
    assert np.array_equal(fa[20][:, 1:10], fb[20][:, 1:10]), np.abs(fa[20][:, 1:10] - fb[20][:, 1:10]).max()


@pytest.mark.gpu
def test_deck_errors_match_the_reference_strings(tmp_path):
    need_binary()
    workload.write_data(str(tmp_path / "4.data"), workload.dpd_fluid(4), 4)
    base = ("units lj\natom_style dpd/atomic/meso\nneighbor 0.3 bin\nneigh_modify delay 0 every 5 check no\nread_data 4.data\n"
            "run_style mvv/meso\n")
    cases = [("pair_style dpd/fast/meso 1.0\n", "Illegal pair_style command"),
             ("pair_style dpd/meso 1.0 1\npair_coeff 1 1 15 4.5 3.0\n", "Incorrect args for pair coefficients"),
             ("pair_style dpd/meso 1.0 1\nfix 3 all nve/meso\nrun 1\n", "All pair coeffs are not set"),
             ("pair_style dpd/meso 1.0 1\npair_coeff 1 1 15 4.5 3.0 1.0\nfix 3 all nve\nrun 1\n", "does not act on device-resident atoms")]
    for tail, msg in cases:
        (tmp_path / "in.err").write_text(base + tail)
        out = subprocess.run([LMP, "-in", "in.err", "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
        assert msg in out.stdout + out.stderr, (tail, out.stdout[-500:], out.stderr[-500:])


def test_binary_keeps_stock_styles_with_meso_off(tmp_path):
    """-meso off: the same binary is plain LAMMPS (and is what a CPU baseline would run); -meso on without a GPU fails loudly."""
    if not os.path.exists(LMP):
        pytest.skip("lammps/_build/lmp_meso_b200 not built (dev container: make -C lammps)")
    workload.write_data(str(tmp_path / "4.data"), workload.dpd_fluid(4), 4)
    (tmp_path / "in.stock").write_text("units lj\natom_style atomic\ncommunicate single vel yes\nneighbor 0.3 bin\nread_data 4.data\n"
                                       "pair_style dpd 1.0 1.0 34387\npair_coeff 1 1 15 4.5 1.0\nvelocity all create 1.0 4928 loop all\n"
                                       "fix 1 all nve\ntimestep 0.005\nrun 5\n")
    out = subprocess.run([LMP, "-meso", "off", "-in", "in.stock", "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "Loop time of" in out.stdout, out.stdout[-800:] + out.stderr[-800:]
    from meso_b200 import lib
    if lib.load().meso_device_count() <= 0:
        (tmp_path / "in.meso").write_text("units lj\natom_style dpd/atomic/meso\n")
        out = subprocess.run([LMP, "-in", "in.meso", "-log", "none"], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
        assert "no CUDA device found" in out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ndev,extra", [("0,1", ""), ("0-3", "pe press"), ("0-7", "")])
def test_deck_on_several_gpus_in_one_process(tmp_path, ndev, extra):
    """MESO_DEVICES lists the GPUs this ONE LAMMPS process drives (the image has no MPI): the library splits the box into a brick
    per listed device and LAMMPS keeps seeing one rank (skipped on boxes with fewer GPUs).  Same deck and the same setup forces as
    on a single context, to the tolerance of re-centring the fp32-packed coordinates on each brick.  With the thermostat on, a
    last-bit difference in a velocity re-draws that atom's pair random numbers (the RNG is keyed on fp32 velocity bits), so
    the trajectories agree statistically: same atoms, same momentum, thermo trace within a few per cent (the bit-level check of
    the decomposed path is tests/test_gang.py against the oracle's simulated ranks).  With `pe press` the thermo steps go
    through the phase entry points and the gang's whole-box sums."""
    need_binary()
    from meso_b200 import lib
    nbrick = int(ndev[-1]) + 1
    if lib.load().meso_device_count() < nbrick:
        pytest.skip("needs %d GPUs" % nbrick)
    L, steps = 12, 20
    ncol = 4 + len(extra.split())
    one = tmp_path / "one"
    one.mkdir()
    out1 = run_deck(one, L, "dp", steps, thermo=10, extra=extra, env={"MESO_PAIR_ONCE": "0"})
    assert out1.returncode == 0, out1.stdout[-2000:] + out1.stderr[-2000:]
    many = tmp_path / "many"
    many.mkdir()
    out2 = run_deck(many, L, "dp", steps, thermo=10, extra=extra, env={"MESO_PAIR_ONCE": "0", "MESO_DEVICES": ndev, "CUDA_DEVICE_MAX_CONNECTIONS": "32",
                                                                              "CUDA_MODULE_LOADING": "EAGER"})
    assert out2.returncode == 0, out2.stdout[-2000:] + out2.stderr[-2000:]
    assert "%d bricks on" % nbrick in out2.stdout
    th1, th2 = thermo_rows(out1.stdout, ncol), thermo_rows(out2.stdout, ncol)
    assert th1.shape == th2.shape == (3, ncol)
    assert np.abs(th1[0, 1:ncol - 2] - th2[0, 1:ncol - 2]).max() < 2e-6 * np.abs(th1[0, 1:ncol - 2]).max(), (th1, th2)   # setup: same state
    assert np.abs(th1[:, 1:ncol - 2] / th2[:, 1:ncol - 2] - 1.0).max() < 5e-2, (th1, th2)
    f1, f2 = frames(str(one / "traj.txt")), frames(str(many / "traj.txt"))
    assert np.array_equal(f1[0][:, :7], f2[0][:, :7])                    # same atoms in, by id
    assert np.abs(f1[0][:, 7:10] - f2[0][:, 7:10]).max() < 1e-5 * np.abs(f1[0][:, 7:10]).max()
    assert np.array_equal(f1[steps][:, 0], f2[steps][:, 0])              # nobody lost across 4 rebuilds with migration
    assert np.abs(f2[steps][:, 4:7].sum(axis=0) - f2[0][:, 4:7].sum(axis=0)).max() < 1e-9      # pair forces cancel across brick faces too
    dx = np.abs(f1[steps][:, 1:4] - f2[steps][:, 1:4])
    assert np.minimum(dx, L - dx).max() < 0.5                            # same fluid 0.1 time units later (minimum image)


@pytest.mark.gpu
def test_thermo_only_output_steps_skip_the_download(tmp_path):
    """A thermo line made of temp/meso, pe and press reduces on the device: mvv/meso then leaves the per-atom arrays where they
    are (the reference copies everything back at every output step, UM/mvv_meso.cu:411-416).  Same numbers as a run whose dump
    forces the download at the same steps."""
    need_binary()
    L, steps = 10, 30
    a, b = tmp_path / "thermo_only", tmp_path / "with_dump"
    a.mkdir(); b.mkdir()
    env = {"MESO_PAIR_ONCE": "0"}
    out_a = run_deck(a, L, "dp", steps, thermo=10, extra="pe press", dump=False, env=env)
    deck_dump = DUMP.format(steps=10)
    workload.write_data(str(b / ("%d.data" % L)), workload.dpd_fluid(L), L)
    (b / "in.run").write_text(DECK.format(prec="dp", pair="dpd/meso", extra="pe press", thermo=10, steps=steps, dump=deck_dump))
    out_b = subprocess.run([LMP, "-in", "in.run", "-var", "case", str(L), "-log", "none"], cwd=str(b), capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, **env))
    assert out_a.returncode == 0 and out_b.returncode == 0, out_a.stdout[-1500:] + out_b.stdout[-1500:]
    ta, tb = thermo_rows(out_a.stdout, 6), thermo_rows(out_b.stdout, 6)
    assert ta.shape == tb.shape == (4, 6)
    assert np.array_equal(ta[:, :4], tb[:, :4]), (ta, tb)
