"""Channel fixes: wall/meso, solid_bound/meso, addforce/meso, pois/meso.

CPU part: the oracle's restatement against an independent numpy evaluation of the reference formulas
(UM/fix_wall_meso.cu:74-193, UM/fix_solid_bound_meso.h:42-56, UM/fix_addforce_meso.cu:72-90,
UM/fix_poiseuille_meso.cu:71-92) and the properties a wall has (nothing ends up outside, reflections are involutions).
GPU part: the CUDA hooks and the fused run loop against the oracle -- fp64 arithmetic, so the bar is 1e-12 on the
force hooks (1e-6 on the erfcf wall: CUDA's and glibc's erfcf differ in the last ulp) and bit-exact on the bounce.
"""
import math

import numpy as np
import pytest

import oracle
from meso_b200 import workload

L = 8
D_WALL, F_WALL = 0.5, 20.0


def channel_atoms(seed=11, hot=1.0):
    x = workload.dpd_fluid(L, seed=seed)
    v = workload.maxwell_velocities(len(x), seed=seed + 1) * hot
    mask = np.where(np.arange(len(x)) % 3 == 0, 3, 1).astype(np.int32)     # every third atom is also in group bit 2
    return x, v, mask


def channel_world(precision=1, fixes=("solid", "add", "pois"), hot=1.0, gamma_sigma=True):
    x, v, mask = channel_atoms(hot=hot)
    g, s = (4.5, 3.0) if gamma_sigma else (0.0, 0.0)
    w = oracle.World((0, 0, 0), (L, L, L), periodic=(1, 1, 0), precision=precision, coeff=oracle.default_coeff(1, 15.0, g, s))
    w.set_atoms(x, v, mask=mask)
    add_fixes(w, fixes, oracle_side=True)
    return w, x, v, mask


def add_fixes(o, fixes, oracle_side):
    """the same fix list on the oracle World or on the Meso mirror (which takes the reference's argument grammar)"""
    for f in fixes:
        if f == "wall":
            o.fix_wall("z", D_WALL, F_WALL) if oracle_side else o.fix("wall/meso", "z", "d", D_WALL, "f", F_WALL)
        elif f == "solid":
            o.fix_solid_bound("z") if oracle_side else o.fix("solid_bound/meso", "z", "rho5rc1s1")
        elif f == "add":
            o.fix_addforce(0.25, -0.5, 0.0, groupbit=2) if oracle_side else o.fix("addforce/meso", 0.25, -0.5, 0.0, groupbit=2)
        elif f == "pois":
            o.fix_pois(2, 0, 0.3) if oracle_side else o.fix("pois/meso", "z", "x", 0.3)


def rho5(h):
    s = 0.282625
    for c in (-1.39021, 2.70259, -2.47678, 0.863184, 0.0664266, 0.0247250, 0.00856667, -0.116714):
        s = s * h + c
    return 75.0 * 6.2831853071796 * (s * h * h + 0.0355959)


def test_oracle_post_force_matches_numpy_formulas():
    from scipy.special import erfc
    w, x, v, mask = channel_world(fixes=("wall", "solid", "add", "pois"))
    w.setup()                                        # includes Fix::setup -> post_force
    a = w.atoms()
    nl = a["nlocal"]
    f_all = a["f"][:nl].copy()
    w.force_clear(); w.pair_compute()
    f_pair = w.atoms()["f"][:nl]
    df = f_all - f_pair
    xs, mk = a["x"][:nl], a["mask"][:nl]
    want = np.zeros_like(df)
    z = xs[:, 2]
    lo, hi = z, L - z
    arg = lambda h: ((h - 0.5 * D_WALL) / D_WALL * 1.732050808).astype(np.float32)
    want[:, 2] += np.where(lo <= D_WALL, F_WALL * erfc(arg(lo)).astype(np.float32), 0.0)
    want[:, 2] -= np.where(hi <= D_WALL, F_WALL * erfc(arg(hi)).astype(np.float32), 0.0)
    want[:, 2] += np.where(lo <= 1.0, rho5(lo), 0.0) - np.where(hi <= 1.0, rho5(hi), 0.0)
    want[:, 0] += np.where(mk & 2, 0.25, 0.0)
    want[:, 1] += np.where(mk & 2, -0.5, 0.0)
    want[:, 0] += np.where(z < 0.5 * L, 0.3, -0.3)
    assert np.abs(df - want).max() < 1e-5 * np.abs(want).max()
    assert np.abs(df[:, :2] - want[:, :2]).max() < 1e-12          # no erfcf in x, y


def test_oracle_bounce_is_a_reflection_and_walls_confine():
    w, x, v, mask = channel_world(fixes=("wall",), hot=3.0)
    w.setup()
    w.run(60)
    a = w.atoms()
    z = a["x"][:a["nlocal"], 2]
    assert z.min() > 0.0 and z.max() < L, (z.min(), z.max())
    assert a["nlocal"] == len(x)
    # one explicit reflection: an atom put outside comes back mirrored with its velocity pointing inwards
    w2 = oracle.World((0, 0, 0), (L, L, L), periodic=(1, 1, 0))
    w2.set_atoms(np.array([[1.0, 1.0, 0.25], [2.0, 2.0, L - 0.5], [3.0, 3.0, 4.0]]), np.array([[0, 0, -100.0], [0, 0, 200.0], [1.0, 1.0, 1.0]]))
    w2.fix_wall("z", 0.0, 0.0)
    w2.initial_integrate()                           # f = 0: pure drift by 0.005 v, two atoms end up beyond the walls
    w2.fix_bounce()
    b = w2.atoms()
    order = np.argsort(b["tag"][:3])
    assert np.allclose(b["x"][:3][order][:, 2], [0.25, L - 0.5, 4.005]) and np.allclose(b["v"][:3][order][:, 2], [100.0, -200.0, 1.0])


def test_oracle_poiseuille_drives_counterflow_and_addforce_adds_momentum():
    w, x, v, mask = channel_world(fixes=("wall", "pois"), precision=0)
    w.setup()
    w.run(150)
    a = w.atoms()
    nl = a["nlocal"]
    z, vx = a["x"][:nl, 2], a["v"][:nl, 0]
    assert vx[z < 0.5 * L].mean() > 0.05 and vx[z >= 0.5 * L].mean() < -0.05
    w, x, v, mask = channel_world(fixes=("solid", "add"), precision=1, gamma_sigma=False)
    w.setup()
    p0 = w.atoms()["v"][:len(x)].sum(axis=0)
    w.run(10)
    a = w.atoms()
    assert a["nlocal"] == len(x)
    p1 = a["v"][:len(x)].sum(axis=0)
    n2 = int((mask & 2).astype(bool).sum())
    # dp/dt = sum of the added forces: n2 * (0.25, -0.5) over 10 steps of 0.005 -- conservative pair forces cancel (to the
    # fp32 packing of the coordinates), the walls act along z only
    assert np.allclose((p1 - p0)[:2], np.array([0.25, -0.5]) * n2 * 10 * 0.005, rtol=1e-4)


# ------------------------------------------------------------------------------------------------ GPU parity
def gpu_pair(precision, fixes, hot=1.0):
    from meso_b200.engine import Meso
    x, v, mask = channel_atoms(hot=hot)
    m = Meso(0)
    m.box((0.0, 0.0, 0.0), (L, L, L), (1, 1, 0))
    m.masses([0.0, 1.0])
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/fast/meso" if precision == "sp" else "dpd/meso", 1.0, 419084618)
    m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)
    m.upload(x, v, mask=mask)
    add_fixes(m, fixes, oracle_side=False)
    w, _, _, _ = channel_world(1 if precision == "dp" else 0, fixes, hot=hot)
    return m, w


def rel_err(a, b):
    d = np.linalg.norm(a - b, axis=1)
    mag = np.linalg.norm(b, axis=1)
    return float((d / np.maximum(mag, mag.mean())).max())


@pytest.mark.gpu
@pytest.mark.parametrize("fixes,tol", [(("solid", "add", "pois"), 1e-12), (("wall",), 1e-6), (("wall", "solid", "add", "pois"), 1e-6)])
def test_gpu_setup_post_force_matches_oracle(fixes, tol):
    m, w = gpu_pair("dp", fixes)
    m.setup(); w.setup()
    ag, ao = m.download(), w.atoms()
    nl = ao["nlocal"]
    assert np.array_equal(ag["tag"], ao["tag"][:nl])
    assert rel_err(ag["f"], ao["f"][:nl]) < tol
    # the stand-alone hook adds the same amount again, fix by fix
    for h in range(len(fixes)):
        m.fix_post_force(h); w.fix_post_force(h)
    assert rel_err(m.download(("f",))["f"], w.atoms()["f"][:nl]) < tol
    m.close()


@pytest.mark.gpu
def test_gpu_bounce_hook_bit_exact():
    m, w = gpu_pair("dp", ("wall", "solid"), hot=4.0)
    m._push_coeff()
    for _ in range(3):                                # no setup: f = 0, pure fp64 drift (bit-identical inputs for the bounce)
        m.initial_integrate(); w.initial_integrate()
    xg0 = m.download(("x",))["x"]
    assert (xg0[:, 2] < 0).any() or (xg0[:, 2] > L).any(), "test needs atoms beyond the walls"
    m.fix_bounce(); w.fix_bounce()
    ag, ao = m.download(("x", "v")), w.atoms()
    nl = ao["nlocal"]
    assert np.array_equal(ag["x"], ao["x"][:nl]) and np.array_equal(ag["v"], ao["v"][:nl])
    assert ag["x"][:, 2].min() >= 0 and ag["x"][:, 2].max() <= L
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("once", ["1", "0"])
def test_gpu_channel_trajectory_fp64_lockstep(monkeypatch, once):
    """17 steps (three rebuilds) of a driven channel: bounce fused into the step-boundary pass, post_force hook after the
    pair kernel, in both run loops; x, v stay on the oracle's fp64 trajectory and nobody leaves the channel."""
    monkeypatch.setenv("MESO_PAIR_ONCE", once)
    m, w = gpu_pair("dp", ("solid", "add", "pois"), hot=2.0)
    m.setup(); w.setup()
    m.run(17); w.run(17)
    ag, ao = m.download(), w.atoms()
    nl = ao["nlocal"]
    assert nl == len(ag["tag"]) and np.array_equal(ag["tag"], ao["tag"][:nl])
    assert np.abs(ag["x"] - ao["x"][:nl]).max() < 1e-9 and np.abs(ag["v"] - ao["v"][:nl]).max() < 1e-9
    assert rel_err(ag["f"], ao["f"][:nl]) < 1e-9
    assert ag["x"][:, 2].min() > 0 and ag["x"][:, 2].max() < L
    m.close()


@pytest.mark.gpu
def test_gpu_channel_fp32_flow_profile_and_grammar_errors():
    from meso_b200.engine import MesoError
    m, w = gpu_pair("sp", ("wall", "pois"))
    m.setup()
    m.run(300)
    a = m.download(("x", "v"))
    z, vx = a["x"][:, 2], a["v"][:, 0]
    assert z.min() > 0 and z.max() < L
    assert vx[z < 0.5 * L].mean() > 0.05 and vx[z >= 0.5 * L].mean() < -0.05     # counter-flowing Poiseuille halves
    assert 0.7 < m.temperature() < 1.6
    for args, msg in ((("wall/meso", "z"), "Illegal fix MesoFixWall"), (("wall/meso", "d", 1.0, "f", 2.0), "insufficient arguments"),
                      (("wall/meso", "z", "f", 1.0, "d"), "after 'd'"), (("solid_bound/meso", "z"), "force kernel unspecified"),
                      (("solid_bound/meso", "rho5rc1s1"), "dimension unspecified"), (("addforce/meso", 1.0, 2.0), "Illegal fix addforce/meso"),
                      (("pois/meso", "z", "x"), "Illegal fix CUDAPoiseuille"), (("nope/meso",), "Invalid fix style")):
        with pytest.raises(MesoError, match=msg):
            m.fix(*args)
    m.unfix_all()
    m.close()


# ------------------------------------------------------------------------------------------------ rdf/fast/meso
GOLD_EQ = None


def gold_eq():
    global GOLD_EQ
    if GOLD_EQ is None:
        import os
        GOLD_EQ = np.load(os.path.join(os.path.dirname(__file__), "golden", "stock_dpd_equilibrium_L10.npz"))
    return GOLD_EQ


def g_of_r(hist, samples, ni, nj, volume, rc=1.0):
    nbin = len(hist)
    b = rc / nbin
    i = np.arange(nbin)
    return hist / ni / samples / (4.0 / 3.0 * 3.1415 * ((b * (i + 1)) ** 3 - (b * i) ** 3)) / (nj / volume)


def test_oracle_rdf_histogram_equals_brute_force_and_group_selection():
    Lb = 6
    x = workload.dpd_fluid(Lb)
    mask = np.where(np.arange(len(x)) % 4 == 0, 3, 1).astype(np.int32)
    w = oracle.World((0, 0, 0), (Lb, Lb, Lb))
    w.set_atoms(x, workload.maxwell_velocities(len(x)), mask=mask)
    h_all = w.fix_rdf(40)
    h_sub = w.fix_rdf(40, groupbit=2, other=1)
    w.setup()
    assert w.rdf(h_all)[1] == 0                      # MesoFixRDFFast::setup is empty: no sample at setup
    w.fix_post_force()
    d = x[:, None, :] - x[None, :, :]
    d -= Lb * np.round(d / Lb)
    r = np.sqrt((d * d).sum(-1))
    np.fill_diagonal(r, 9.0)
    bf = np.histogram(r[r < 1.0], bins=40, range=(0, 1))[0]
    hist, s, ni, nj = w.rdf(h_all)
    assert s == 1 and ni == nj == len(x) and hist.sum() == bf.sum() and np.abs(hist - bf).max() <= 2   # fp32 distances at bin edges
    sub = (mask & 2) != 0
    bf2 = np.histogram(r[sub][r[sub] < 1.0], bins=40, range=(0, 1))[0]
    hist2, s2, ni2, nj2 = w.rdf(h_sub)
    assert ni2 == sub.sum() and nj2 == len(x) and hist2.sum() == bf2.sum() and np.abs(hist2 - bf2).max() <= 2


def test_oracle_equilibrium_matches_stock_lammps_statistically():
    """north_star check (4) on the CPU side: <T>, <P> and g(r) of the MESO algorithm (its own TEA random stream) against
    stock pair_style dpd (RanMars stream) -- fixture from tests/golden/make_lammps_rdf_golden.py."""
    g = gold_eq()
    Lb = int(g["L"])
    x = workload.dpd_fluid(Lb)
    w = oracle.World((0, 0, 0), (Lb, Lb, Lb), precision=0)
    w.set_atoms(x, workload.maxwell_velocities(len(x)))
    w.setup()
    w.run(500)
    h = w.fix_rdf(int(g["nbin"]), every=10)
    ts, ps = [], []
    for _ in range(10):
        w.run(49)
        w.run(1, eflag=1, vflag=1)
        vir, _ = w.virial()
        t = w.temperature()
        ts.append(t)
        ps.append(((3 * len(x) - 3) * t + vir[:, :3].sum()) / (3.0 * Lb ** 3))
    hist, s, ni, nj = w.rdf(h)
    assert s == 50
    gr = g_of_r(hist, s, ni, nj, float(Lb) ** 3)
    sel = g["r"] > 0.3
    assert np.abs(gr[sel] - g["g"][sel]).max() < 0.05, np.abs(gr[sel] - g["g"][sel]).max()
    assert abs(np.mean(ts) - g["temp"].mean()) < 0.02 and abs(np.mean(ps) - g["press"].mean()) < 0.5, (np.mean(ts), np.mean(ps))


@pytest.mark.gpu
def test_gpu_rdf_matches_oracle_and_samples_on_cadence():
    m, w = gpu_pair("sp", ())
    hg = m.fix("rdf/fast/meso", "output", "unused.txt", "nbin", 64, "every", 5)
    hg2 = m.fix("rdf/fast/meso", "output", "unused2.txt", "nbin", 32, "other", 2, groupbit=2)
    ho = w.fix_rdf(64, every=5)
    ho2 = w.fix_rdf(32, groupbit=2, other=2)
    m.setup(); w.setup()
    m.fix_post_force(); w.fix_post_force()           # ntimestep 0: both cadences hit
    vol = float(L) ** 3
    for (a, b) in ((hg, ho), (hg2, ho2)):
        r, gr, hist_g, s_g = m.rdf(a, vol)
        hist_o, s_o, ni, nj = w.rdf(b)
        assert s_g == s_o == 1 and hist_g.sum() == hist_o.sum() and np.abs(hist_g - hist_o).max() <= 2, np.abs(hist_g - hist_o).max()
        assert np.allclose(gr, g_of_r(hist_o, s_o, ni, nj, vol) * hist_g / np.maximum(hist_o, 1), rtol=1e-12, atol=1e-12)
    m.run(20)
    assert m.rdf(hg, vol)[3] == 1 + 4 and m.rdf(hg2, vol)[3] == 1 + 20     # steps 5, 10, 15, 20 / every step
    m.close()


@pytest.mark.gpu
def test_gpu_equilibrium_temperature_pressure_rdf_match_stock_lammps():
    """north_star check (4): trajectory observables of the CUDA path (fp32 style, pair-once loop) against stock LAMMPS."""
    from meso_b200.engine import dpd_fluid_deck
    g = gold_eq()
    Lb = int(g["L"])
    m = dpd_fluid_deck(Lb, "sp")
    h = m.fix("rdf/fast/meso", "output", "unused.txt", "nbin", int(g["nbin"]), "every", 10)
    m.setup()
    m.run(1000)
    base = m.rdf(h, float(Lb) ** 3)
    ts, ps = [], []
    n = 4 * Lb ** 3
    for _ in range(40):
        m.run(49)
        # one step with tallies through the phase entry points
        m.ntimestep = m.ntimestep + 1
        m.initial_integrate()
        if m.neighbor_decide():
            m.rebuild()
        else:
            m.forward_comm()
        m.force_clear(vflag=1)
        m.pair_compute(eflag=1, vflag=1)
        m.fix_post_force()
        m.final_integrate()
        vir, _ = m.virial()
        t = m.temperature()
        ts.append(t)
        ps.append(((3 * n - 3) * t + vir[:3].sum()) / (3.0 * Lb ** 3))
    r, gr, hist, s = m.rdf(h, float(Lb) ** 3)
    hist, s = hist - base[2], s - base[3]
    gr = g_of_r(hist, s, n, n, float(Lb) ** 3)
    sel = g["r"] > 0.3
    assert s == 200 and np.abs(gr[sel] - g["g"][sel]).max() < 0.03, np.abs(gr[sel] - g["g"][sel]).max()
    # 40 samples of 4000 atoms: sigma of the means ~0.002 in T and ~0.07 in P (stock: 0.0126 / 0.43 per sample); the fp32 pair-once loop
    # sums with atomics, so every run is its own realisation -- bounds at ~5 sigma
    assert abs(np.mean(ts) - g["temp"].mean()) < 0.012 and abs(np.mean(ps) - g["press"].mean()) < 0.35, (np.mean(ts), np.mean(ps))
    m.close()


# ------------------------------------------------------------------------------------------------ BASELINE configs[4]
# amphiphilic bead-spring chains in a driven channel: bonds + 1-2 exclusions + three atom types + walls + body force together
AMPHI_COEFF = {(1, 1): 25.0, (2, 2): 25.0, (3, 3): 25.0, (1, 2): 27.0, (1, 3): 40.0, (2, 3): 40.0}


def amphi_worlds(precision, gpu):
    Lc = 8
    x, typ, tag, nb, bt, ba = workload.amphiphilic_channel(Lc)
    v = workload.maxwell_velocities(len(x), seed=99)
    coeff = np.zeros((3, 3, 7))
    for (a, b), a0 in AMPHI_COEFF.items():
        coeff[a - 1, b - 1] = coeff[b - 1, a - 1] = [1.0, 1.0, 1.0, 1.0, a0, 4.5, 3.0]
    w = oracle.World((0, 0, 0), (Lc, Lc, Lc), periodic=(1, 1, 0), ntypes=3, mass=[0, 1, 1, 1], coeff=coeff.reshape(-1, 7), precision=precision)
    w.set_atoms(x, v, tag=tag, type=typ)
    w.set_bonds(nb, bt, ba, tag=tag, k=[0.0, 50.0], r0=[0.0, 0.5], special_lj12=0.0)
    w.fix_solid_bound("z"); w.fix_pois(2, 0, 0.2)
    m = None
    if gpu:
        from meso_b200.engine import Meso
        m = Meso(0)
        m.box((0.0, 0.0, 0.0), (Lc, Lc, Lc), (1, 1, 0))
        m.masses([0.0, 1.0, 1.0, 1.0])
        m.neighbor(0.3, "bin")
        m.neigh_modify(delay=0, every=5, check=False)
        m.pair_style("dpd/fast/meso" if precision == 0 else "dpd/meso", 1.0, 419084618)
        for (a, b), a0 in AMPHI_COEFF.items():
            m.pair_coeff(a, b, a0, 4.5, 3.0, 1.0, 1.0)
        m.timestep(0.005)
        m.upload(x, v, tag=tag, type=typ)
        m.bond_style("harmonic/meso", 1)
        m.bond_coeff(1, 50.0, 0.5)
        m.special_bonds(0.0)
        m.bonds(nb, bt, ba)
        m.fix("solid_bound/meso", "z", "rho5rc1s1"); m.fix("pois/meso", "z", "x", 0.2)
    return m, w, Lc, len(x)


def test_oracle_amphiphilic_channel_is_stable():
    _, w, Lc, n = amphi_worlds(0, gpu=False)
    w.setup()
    w.run(60)
    a = w.atoms()
    assert a["nlocal"] == n and a["x"][:n, 2].min() > 0 and a["x"][:n, 2].max() < Lc
    assert 0.7 < w.temperature() < 4.5                   # the random start releases conservative energy; the thermostat needs ~400 steps
    w.bond_compute(1, 0)
    assert w.bond_energy() < 50.0 * 0.04 * 714            # relaxing from the stretched start (714 bonds at r - r0 = 0.2)


@pytest.mark.gpu
def test_gpu_amphiphilic_channel_fp64_lockstep_and_fp32_run():
    m, w, Lc, n = amphi_worlds(1, gpu=True)
    m.setup(); w.setup()
    m.run(12); w.run(12)
    ag, ao = m.download(), w.atoms()
    assert np.array_equal(ag["tag"], ao["tag"][:n])
    assert np.abs(ag["x"] - ao["x"][:n]).max() < 1e-9 and np.abs(ag["v"] - ao["v"][:n]).max() < 1e-9
    assert rel_err(ag["f"], ao["f"][:n]) < 1e-9
    cg, rg = m.neighbors()
    co, ro = w.neighbors()
    msk = np.arange(ro.shape[1])[None, :] < co[:, None]
    assert np.array_equal(cg, co) and np.array_equal(rg[msk], ro[msk])       # exclusion-filtered rows after two rebuilds
    m.close()
    m, _, _, _ = amphi_worlds(0, gpu=True)
    m.setup()
    m.run(600)
    a = m.download(("x", "v"))
    assert a["x"][:, 2].min() > 0 and a["x"][:, 2].max() < Lc and 0.7 < m.temperature() < 1.5
    z, vx = a["x"][:, 2], a["v"][:, 0]
    assert vx[z < 0.5 * Lc].mean() > 0.02 and vx[z >= 0.5 * Lc].mean() < -0.02
    m.close()
