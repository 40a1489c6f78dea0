"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * integer work -- sort keys, permutation, ghost order, cells, neighbor lists (sets AND order),
    packed coordinates, TEA signatures -- bit-exact;
  * dpd/meso (fp64) forces and the fp64 Gaussian/polynomials: 1e-12 relative;
  * dpd/fast/meso (fp32) forces: 1e-5 relative (|dF| <= 1e-5 * max(|F_i|, mean|F|)).
"""
import numpy as np
import pytest

import oracle
from meso_b200 import lib, workload
from meso_b200.engine import Meso, MesoError

pytestmark = pytest.mark.gpu

SP_TOL = 1e-5
DP_TOL = 1e-12


def make_pair(L, precision="sp", seed=419084618, ntypes=1, periodic=(1, 1, 1), coeff=None, mass=None, types=None,
              x=None, v=None, skin=0.3, every=5, mask=None, tag=None, gamma_sigma=True):
    dims = (L, L, L) if np.isscalar(L) else tuple(L)
    if x is None:
        x = workload.dpd_fluid(L if np.isscalar(L) else tuple(int(d) for d in dims))
    if v is None:
        v = workload.maxwell_velocities(len(x))
    mass = [0.0] + [1.0] * ntypes if mass is None else mass
    m = Meso(0)
    m.box((0.0, 0.0, 0.0), dims, periodic)
    m.masses(mass)
    m.neighbor(skin, "bin")
    m.neigh_modify(delay=0, every=every, check=False)
    m.pair_style("dpd/fast/meso" if precision == "sp" else "dpd/meso", 1.0, seed)
    if coeff is None:
        g, s = (4.5, 3.0) if gamma_sigma else (0.0, 0.0)
        m.pair_coeff("*", "*", 15, g, s, 1.0, 1.0)
    else:
        for (i, j, args) in coeff:
            m.pair_coeff(i, j, *args)
    m.timestep(0.005)
    m.upload(x, v, tag=tag, type=types, mask=mask)
    m._push_coeff()
    w = oracle.World((0, 0, 0), dims, periodic=periodic, ntypes=ntypes, mass=mass, coeff=m._coeff.reshape(-1, 7),
                     cut_max=float(m._coeff[..., 0].max()), skin=skin, every=every, seed=seed, dt=0.005,
                     precision=0 if precision == "sp" else 1)
    w.set_atoms(x, v, tag=tag, type=types, mask=mask)
    return m, w


def force_err(fg, fo):
    d = np.linalg.norm(fg - fo, axis=1)
    mag = np.linalg.norm(fo, axis=1)
    scale = np.maximum(mag, mag.mean() if len(mag) else 1.0)
    return float((d / scale).max()) if len(d) else 0.0


def assert_state_identical(m, w, check_forces=True, tol=None, precision="sp"):
    """everything the rebuild produces, bit for bit"""
    cg, co = m.counts(), w.counts()
    for k in ("nlocal", "nghost", "n_bulk", "n_border"):
        assert cg[k] == co[k], (k, cg, co)
    mg, bsg, big, ncol = m.bins()
    mo, bso, bio = w.bins()
    assert mg == mo and bsg == bso and big == bio and ncol == co["n_col"]
    kg, pg = m.reorder()
    ko, po = w.reorder()
    assert np.array_equal(kg, ko), "sorted reorder keys differ"
    assert np.array_equal(pg, po), "permutation differs"
    ag, ao = m.download(), w.atoms()
    nl = cg["nlocal"]
    assert np.array_equal(ag["tag"], ao["tag"][:nl])
    assert np.array_equal(ag["x"], ao["x"][:nl]) and np.array_equal(ag["v"], ao["v"][:nl])
    assert np.array_equal(ag["type"], ao["type"][:nl]) and np.array_equal(ag["mask"], ao["mask"][:nl])
    assert np.array_equal(ag["image"], ao["image"])
    gg = m.ghosts()
    assert np.array_equal(gg["tag"], ao["tag"][nl:]) and np.array_equal(gg["type"], ao["type"][nl:])
    assert np.array_equal(gg["x"], ao["x"][nl:]) and np.array_equal(gg["v"], ao["v"][nl:])
    sg, cag = m.cells()
    so, cao = w.cells()
    assert np.array_equal(sg, so) and np.array_equal(cag, cao)
    rng = np.random.default_rng(0)
    ncell = mg[0] * mg[1] * mg[2]
    for c in list(rng.integers(0, ncell, 40)) + [0, ncell - 1]:
        assert np.array_equal(m.stencil(int(c)), w.stencil(int(c)))
    cntg, rowsg = m.neighbors()
    cnto, rowso = w.neighbors()
    assert np.array_equal(cntg, cnto), "neighbor counts differ"
    mask = np.arange(rowso.shape[1])[None, :] < cnto[:, None]
    assert np.array_equal(rowsg[mask], rowso[mask]), "neighbor lists differ (order included)"
    tg, _ = m.pair_table()
    to = w.neighbors_transposed()
    valid = to >= 0
    assert np.array_equal(tg[valid], to[valid]), "tile-transposed table differs"
    c4g, v4g = m.packed()
    c4o, v4o = w.packed()
    assert np.array_equal(c4g.view(np.uint32), c4o.view(np.uint32)), "packed coordinates differ"
    assert np.array_equal(v4g.view(np.uint32), v4o.view(np.uint32)), "packed velocities / TEA signatures differ"
    if check_forces:
        e = force_err(ag["f"], ao["f"])
        assert e <= (tol if tol is not None else (SP_TOL if precision == "sp" else DP_TOL)), e


# ------------------------------------------------------------------ A8 / A9
def test_device_fp64_math_and_gaussians_match_oracle():
    m = Meso(0)
    L = oracle.lib()
    rng = np.random.default_rng(7)
    x = rng.uniform(1e-3, 50.0, 4096)
    for name, dom in (("rsqrt", x), ("rcp", x), ("sqrtd", x), ("log2d_frac", rng.uniform(1, 2, 4096)),
                      ("exp2d_frac", rng.uniform(0, 1, 4096)), ("sinpi", rng.uniform(0, 1, 4096)), ("cospi", rng.uniform(0, 1, 4096))):
        got = m.eval_math(name, dom)
        ref = np.array([getattr(L, "orc_" + name)(float(t)) for t in dom])
        assert np.array_equal(got, ref), name
    a, b = rng.uniform(1e-6, 1, 4096), rng.choice([0.25, 0.5, 1.0, 2.0, 1.37], 4096)
    assert np.array_equal(m.eval_math("powd", a, b), np.array([L.orc_powd(float(p), float(q)) for p, q in zip(a, b)]))
    u = np.concatenate([rng.integers(1, 2 ** 32, 4092, dtype=np.uint64).astype(np.uint32), np.array([1, 2, 3, 2 ** 32 - 1], np.uint32)])
    assert np.array_equal(m.eval_log2u(u), np.array([L.orc_log2u(int(t)) for t in u]))
    si = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32)
    sj = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32)
    sp, dp = m.eval_gaussian(si, sj)
    rdp = np.array([L.orc_gaussian_dp(int(p), int(q)) for p, q in zip(si, sj)])
    rsp = np.array([L.orc_gaussian_sp(int(p), int(q)) for p, q in zip(si, sj)], dtype=np.float32)
    assert np.array_equal(dp, rdp), "fp64 Gaussian must be bit-exact (same FMA chains)"
    # the fp32 kernel uses MUFU lg2/sin/sqrt (non-branching): ~5e-6 abs on |g| <= 4, inside the 1e-5 force bar
    assert np.abs(sp - rsp).max() <= 2e-5, np.abs(sp - rsp).max()
    # degenerate inputs: v1 == 0 (log2f(0) = -inf) must clamp to +-4, never NaN
    assert np.isfinite(sp).all() and np.abs(sp).max() <= 4.0 and np.abs(dp).max() <= 4.0
    m.close()


# ------------------------------------------------------------------ rebuild + forces at setup
@pytest.mark.parametrize("precision", ["sp", "dp"])
def test_setup_parity_small(precision):
    m, w = make_pair(10, precision)
    m.setup()
    w.setup()
    assert_state_identical(m, w, precision=precision)
    assert abs(m.temperature() - w.temperature()) < 1e-12
    m.close()


@pytest.mark.parametrize("precision", ["sp", "dp"])
def test_setup_parity_case25(precision):
    """BASELINE.json configs[0]: the 25^3 rho=4 DPD fluid (62,500 particles)."""
    m, w = make_pair(25, precision)
    m.setup()
    w.setup()
    assert_state_identical(m, w, precision=precision)
    c = m.counts()
    assert c["nlocal"] == 62500 and m.bins()[0] == [21, 21, 21] and m.bins()[3] == 160
    m.close()


def test_conservative_only_forces():
    """gamma = sigma = 0: the part that is also pinned against stock LAMMPS pair_style dpd (tests/golden)."""
    for precision in ("sp", "dp"):
        m, w = make_pair(8, precision, gamma_sigma=False)
        m.setup(); w.setup()
        e = force_err(m.download()["f"], w.atoms()["f"])
        assert e <= (SP_TOL if precision == "sp" else DP_TOL), e
        m.close()


def test_conservative_forces_match_stock_lammps_golden():
    """CUDA path vs the UNMODIFIED stock LAMMPS pair_style dpd of the reference tree (fixture:
    tests/golden/make_lammps_golden.py), gamma = sigma = 0, within 1e-5 relative."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "stock_dpd_conservative_L8.npz"))
    for precision in ("sp", "dp"):
        m, _ = make_pair(int(gold["L"]), precision, gamma_sigma=False)
        m.setup(eflag=1, vflag=1)
        d = m.download(("f", "tag"))
        f = np.empty_like(d["f"])
        f[d["tag"] - 1] = d["f"]
        assert force_err(f, gold["f"]) <= SP_TOL
        vir, e = m.virial()
        n, L = len(f), float(gold["L"])
        assert abs(e / n - float(gold["pe"])) < 1e-5 * float(gold["pe"])
        assert abs(vir[:3].sum() / (3 * L ** 3) - float(gold["press"])) < 1e-4 * float(gold["press"])
        m.close()


def test_energy_virial():
    for precision in ("sp", "dp"):
        m, w = make_pair(8, precision)
        m.setup(eflag=1, vflag=1); w.setup(eflag=1, vflag=1)
        vg, eg = m.per_atom_virial()
        vo, eo = w.virial()
        tol = 2e-5 if precision == "sp" else 1e-11
        assert np.abs(vg - vo).max() <= tol * max(1.0, np.abs(vo).max())
        assert np.abs(eg - eo).max() <= tol * max(1.0, np.abs(eo).max())
        tot, etot = m.virial()
        # the device-side reduction (new: the reference never reduces its per-atom virial) against its own per-atom data
        assert np.allclose(tot, vg.sum(0), rtol=1e-12, atol=1e-9) and abs(etot - eg.sum()) <= 1e-12 * abs(eg.sum()) + 1e-9
        assert np.allclose(tot, vo.sum(0), rtol=tol * 10, atol=1e-3)
        m.close()


# ------------------------------------------------------------------ edge cases
def test_two_types_masses_ragged_box_and_group_mask():
    rng = np.random.default_rng(3)
    L = (9, 7, 12)
    x = workload.dpd_fluid(L, seed=5)
    n = len(x)
    types = rng.integers(1, 3, n).astype(np.int32)
    mask = np.where(rng.random(n) < 0.9, 1, 2).astype(np.int32)       # 10 % of the atoms outside group bit 1
    tag = (rng.permutation(n) * 3 + 7).astype(np.int32)               # non-contiguous, shuffled tags
    coeff = [(1, 1, (15, 4.5, 3.0, 1.0, 1.0)), (1, 2, (20, 3.0, 2.449489742783178, 0.5, 0.9)), (2, 2, (25, 6.0, 3.4641016151377544, 2.0, 0.8))]
    for precision in ("sp", "dp"):
        m, w = make_pair(L, precision, ntypes=2, coeff=coeff, mass=[0.0, 1.0, 2.5], types=types, x=x, mask=mask, tag=tag)
        m.setup(); w.setup()
        assert_state_identical(m, w, precision=precision, tol=2e-5 if precision == "sp" else 1e-11)
        m.close()


def test_non_periodic_dimension_and_tiny_box():
    x = workload.dpd_fluid((6, 5, 4), seed=9)
    m, w = make_pair((6, 5, 4), "dp", periodic=(1, 0, 1), x=x)
    m.setup(); w.setup()
    assert_state_identical(m, w, precision="dp")
    m.close()
    # 4^3: three inner cells per dimension
    m, w = make_pair(4, "dp")
    m.setup(); w.setup()
    assert_state_identical(m, w, precision="dp")
    m.close()
    # 2.5 wide in x: one inner cell, the lo and hi send slabs overlap, so atoms are sent both ways
    dims = np.array([2.5, 4.0, 5.5])
    x = np.random.default_rng(4).uniform(0, 1, (220, 3)) * dims
    m, w = make_pair(tuple(dims), "dp", x=x)
    m.setup(); w.setup()
    assert w.counts()["n_bulk"] == 0 and w.counts()["nghost"] > 3 * 220
    assert_state_identical(m, w, precision="dp")
    m.close()


def test_dense_fluid_rows_longer_than_the_staging_queue():
    """rho = 8: ~72 stored neighbors per atom (> 2 table tiles): the tile build picks a smaller block of cells and deeper hit
    queues for this density; also the 32-slot tile wrap of the transposed table and the early-drain path of the force kernel."""
    x = workload.dpd_fluid(6, rho=8, seed=13)
    for precision in ("sp", "dp"):
        m, w = make_pair(6, precision, x=x)
        m.setup(); w.setup()
        assert w.neighbors()[0].max() > 64 and w.counts()["n_col"] == 320
        assert_state_identical(m, w, precision=precision, tol=2e-5 if precision == "sp" else 1e-11)
        m.close()


def test_neighbor_build_plain_walk_gives_the_same_table(monkeypatch):
    """MESO_NB_SLOW=1 builds every row with the plain walk of the 27 stencil cells in global memory (the fall-back kernel of
    the tile build): same counts, same canonical rows, same production rows entry for entry -- and the same state after a run
    across a rebuild."""
    for L, precision in ((9, "dp"), ((7, 9, 12), "dp")):      # fp64: the run stays in lockstep with the oracle
        out = []
        for slow in ("0", "1"):
            monkeypatch.setenv("MESO_NB_SLOW", slow)
            m, w = make_pair(L, precision)
            m.setup(); w.setup()
            assert_state_identical(m, w, precision=precision)
            cnt, own, rows = m.pair_rows()
            out.append((cnt, own, rows))
            m.run(6); w.run(6)
            cntg, rowsg = m.neighbors()
            cnto, rowso = w.neighbors()
            mask = np.arange(rowso.shape[1])[None, :] < cnto[:, None]
            assert np.array_equal(cntg, cnto) and np.array_equal(rowsg[mask], rowso[mask])
            m.close()
        for a, b in zip(out[0], out[1]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("L", [10, (7, 9, 12)])
def test_production_rows_are_the_canonical_rows_in_two_parts(L):
    """The table the force kernels read: [owned][other].  A row holds exactly the reference's entries; "owned" is exactly the
    rule of the pair-once kernel (ghost j, or (i+j) odd ? i<j : i>j), so every local pair is owned by exactly one of its two
    rows; inside a part the entries follow the walk (ascending position in the cell order)."""
    m, w = make_pair(L, "sp")
    m.setup(); w.setup()
    cnt, own, rows = m.pair_rows()
    cnto, rowso = w.neighbors()                              # reference order: core entries, then skin entries reversed
    nl = len(cnt)
    assert np.array_equal(cnt, cnto)
    cs, ca = w.cells()
    pos = np.empty(len(ca), np.int64)
    pos[ca] = np.arange(len(ca))
    owner = {}
    for i in range(nl):
        r = rows[i, :cnt[i]]
        assert sorted(r.tolist()) == sorted(rowso[i, :cnto[i]].tolist())
        mine = (r >= nl) | np.where((i + r) % 2 == 1, i < r, i > r)
        assert mine[:own[i]].all() and not mine[own[i]:].any(), i
        for part in (r[:own[i]], r[own[i]:]):
            assert np.all(np.diff(pos[part]) > 0), i
        for j in r[:own[i]].tolist():
            if j < nl:
                key = (min(i, j), max(i, j))
                assert key not in owner
                owner[key] = i
    npairs = sum(int((rowso[i, :cnto[i]] < nl).sum()) for i in range(nl)) // 2
    assert len(owner) == npairs
    m.close()


def test_atoms_on_cell_and_box_boundaries():
    x = workload.dpd_fluid(6, seed=2)
    x[:50] = np.round(x[:50])            # exactly on unit-lattice planes, including 0.0
    x[50:60, 0] = 6.0                    # exactly boxhi: wraps to 0 with an image flag
    x[60:70, 1] = -0.25                  # outside: wraps up
    x[70:80, 2] = 6.0 + 1e-12
    x[80:90] = x[90:100] + 1e-7          # near-coincident pairs (rsq ~ 1e-14 >= EPSILON_SQ) -> huge but finite forces
    m, w = make_pair(6, "dp", x=x)
    m.setup(); w.setup()
    assert_state_identical(m, w, precision="dp", tol=1e-11)
    m.close()


def test_empty_and_single_atom():
    for n in (0, 1):
        x = np.full((n, 3), 2.5)
        m, w = make_pair(5, "sp", x=x, v=np.zeros((n, 3)))
        m.setup(); w.setup()
        c = m.counts()
        assert c["nlocal"] == n and c["nghost"] == w.counts()["nghost"]
        if n:
            assert np.array_equal(m.pair_count(), w.neighbors()[0])
            assert np.abs(m.download()["f"]).max() == 0.0
        m.run(3)
        m.sync()
        m.close()


def test_errors():
    m = Meso(0)
    with pytest.raises(MesoError, match="Illegal pair_style command"):
        m.pair_style("dpd/fast/meso", 1.0)
    m.box((0, 0, 0), (5, 5, 5))
    m.masses([0.0, 1.0, 1.0])
    m.pair_style("dpd/fast/meso", 1.0, 1)
    with pytest.raises(MesoError, match="Incorrect args for pair coefficients"):
        m.pair_coeff(1, 1, 15, 4.5)
    m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0)
    m.upload(workload.dpd_fluid(5))
    with pytest.raises(MesoError, match="All pair coeffs are not set"):
        m.setup()
    m.close()
    m = Meso(0)
    m.box((0, 0, 0), (1.2, 5, 5))
    m.masses([0.0, 1.0])
    m.pair_style("dpd/fast/meso", 1.0, 1)
    m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0)
    m.upload(np.random.default_rng(0).uniform(0, 1, (100, 3)) * np.array([1.2, 5, 5]))
    with pytest.raises(MesoError, match="thinner than the ghost cutoff"):
        m.setup()
    m.close()


# ------------------------------------------------------------------ time stepping
def test_dp_trajectory_lockstep_12_steps():
    """fp64 path stays locked to the oracle across two rebuilds: x, v to 1e-10, lists and signatures bit-exact."""
    m, w = make_pair(10, "dp")
    m.setup(); w.setup()
    m.run(12); w.run(12)
    assert m.ntimestep == 12 and w.ntimestep == 12
    ag, ao = m.download(), w.atoms()
    nl = ao["nlocal"]
    assert np.array_equal(ag["tag"], ao["tag"][:nl])
    assert np.abs(ag["x"] - ao["x"][:nl]).max() < 1e-10 and np.abs(ag["v"] - ao["v"][:nl]).max() < 1e-10
    assert force_err(ag["f"], ao["f"]) < 1e-10
    cntg, rowsg = m.neighbors()
    cnto, rowso = w.neighbors()
    assert np.array_equal(cntg, cnto)
    mask = np.arange(rowso.shape[1])[None, :] < cnto[:, None]
    assert np.array_equal(rowsg[mask], rowso[mask])
    c4g, v4g = m.packed()
    c4o, v4o = w.packed()
    assert np.array_equal(v4g.view(np.uint32)[:, 3], v4o.view(np.uint32)[:, 3]), "signatures drifted"
    assert abs(m.temperature() - w.temperature()) < 1e-10
    m.close()


def test_phase_api_equals_fused_run(monkeypatch):
    """ModifiedVerlet::run spelled phase by phase (bulk/border split, separate clears) == meso_run, bit for bit, when
    meso_run uses the same two-sided force kernel as the phase entry points (MESO_PAIR_ONCE=0).  The default run loop
    evaluates each local pair once and reduces with atomics, so there only the summation order differs:
    see test_pair_once_run_matches_two_sided_and_oracle."""
    monkeypatch.setenv("MESO_PAIR_ONCE", "0")
    B, O, LOC = lib.MESO_BULK, lib.MESO_BORDER, lib.MESO_LOCAL
    for precision in ("sp", "dp"):
        a, _ = make_pair(8, precision)
        b, _ = make_pair(8, precision)
        a.setup(); b.setup()
        a.run(11)
        for _ in range(11):
            b.ntimestep = b.ntimestep + 1
            b.initial_integrate()
            if b.neighbor_decide():
                b.rebuild()
                b.force_clear(LOC)
                b.pair_compute(B)
            else:
                b.force_clear(B)
                b.pair_compute(B)
                b.forward_comm()
                b.force_clear(O)
            b.pair_compute(O)
            b.final_integrate()
        da, db = a.download(), b.download()
        for k in ("x", "v", "f", "tag", "image"):
            assert np.array_equal(da[k], db[k]), (precision, k)
        a.close(); b.close()


def test_pair_once_run_matches_two_sided_and_oracle(monkeypatch):
    """The default run loop (pair.cu:k_dpd_once: each local pair evaluated once, j side reduced with REDG, second
    half-kick fused into the next step's streaming pass) against the two-sided kernel and the oracle.
    One step from the same state: x and the half-kicked v are bit-identical (they depend on the setup force only),
    the new force agrees to the fp tolerance (every addend is bit-identical, only the summation order differs).
    fp64 additionally stays locked over 11 steps / two rebuilds (x, v to 1e-10, lists + signatures bit-exact in
    test_dp_trajectory_lockstep_12_steps); fp32 trajectories cannot be locked (RNG keys on fp32 velocity bits)."""
    for precision, tol in (("sp", SP_TOL), ("dp", DP_TOL)):
        for ntypes in (1, 2):
            kw = {}
            if ntypes == 2:
                n = 4 * 9 ** 3
                kw = dict(ntypes=2, types=(np.arange(n) % 2 + 1).astype(np.int32), mass=[0.0, 1.0, 2.5],
                          coeff=[(1, 1, (15, 4.5, 3.0, 1.0, 1.0)), (1, 2, (20, 3.0, 2.0, 0.5, 0.9)), (2, 2, (25, 5.0, 3.2, 2.0, 1.0))])
            monkeypatch.setenv("MESO_PAIR_ONCE", "1")
            a, w = make_pair(9, precision, **kw)
            monkeypatch.setenv("MESO_PAIR_ONCE", "0")
            b, _ = make_pair(9, precision, **kw)
            a.setup(); b.setup(); w.setup()
            a.run(1); b.run(1); w.run(1)
            da, db, ao = a.download(), b.download(), w.atoms()
            assert np.array_equal(da["x"], db["x"]) and np.array_equal(da["tag"], db["tag"])
            e_ab = force_err(da["f"], db["f"])
            assert e_ab <= tol, (precision, ntypes, e_ab)
            if precision == "dp":
                assert force_err(da["f"], ao["f"]) <= 1e-10
            else:
                # vs the oracle a few atoms draw different random numbers: the setup forces agree to 1e-6 only (MUFU), so
                # the half-kicked fp32 velocity of a handful of atoms rounds across one of the signature's mantissa bits
                mag = np.linalg.norm(ao["f"], axis=1)
                derr = np.linalg.norm(da["f"] - ao["f"], axis=1) / np.maximum(mag, mag.mean())
                assert (derr > tol).mean() <= 0.12 and np.median(derr) <= tol, (derr > tol).sum()   # ~0.25 % of atoms flip x ~17 pairs each
            assert np.abs(da["v"] - db["v"]).max() <= 0.5 * 0.005 * 40 * tol * 10
            # pairwise antisymmetry; pairs across the periodic boundary see separately fp32-packed image coordinates on
            # the two sides (|dF| ~ 1e-7 |F| each, also in the oracle), everything else cancels to rounding
            assert np.abs(da["f"].sum(0)).max() < 5e-3
            assert np.abs(da["f"].sum(0) - ao["f"][:len(da["f"])].sum(0)).max() < (1e-3 if precision == "sp" else 1e-8)
            # a second run call: the accumulator was handed back clean and f holds the fp64 mirror
            a.run(4); b.run(4)
            if precision == "dp":
                da, db = a.download(), b.download()
                assert np.abs(da["x"] - db["x"]).max() < 1e-10 and np.abs(da["v"] - db["v"]).max() < 1e-10
                assert force_err(da["f"], db["f"]) < 1e-9
            a.close(); b.close()


def test_message_based_halo_path_on_one_rank(monkeypatch):
    """MESO_FORCE_COMM_PATH=1: the multi-GPU halo code (ordered pack -> message -> unpack, sendlists, side-stream
    overlap of the refresh with the bulk force kernel) with every swap partnered to the rank itself.  No migration
    happens on one rank, so even the ghost ORDER must equal the oracle's."""
    monkeypatch.setenv("MESO_FORCE_COMM_PATH", "1")
    for L, precision, periodic in ((10, "dp", (1, 1, 1)), ((9, 7, 12), "sp", (1, 1, 1)), ((6, 5, 4), "dp", (1, 0, 1))):
        m, w = make_pair(L, precision, periodic=periodic)
        m.setup(); w.setup()
        assert_state_identical(m, w, precision=precision)
        # (no dynamics with an open face: the reference DROPS atoms that cross a non-periodic face in Comm::exchange and
        #  LAMMPS then aborts with "Lost atoms"; keeping such runs inside the box is the wall fixes' job)
        if precision == "dp" and all(periodic):
            m.run(12); w.run(12)
            ag, ao = m.download(), w.atoms()
            nl = ao["nlocal"]
            assert np.array_equal(ag["tag"], ao["tag"][:nl])
            assert np.abs(ag["x"] - ao["x"][:nl]).max() < 1e-10 and np.abs(ag["v"] - ao["v"][:nl]).max() < 1e-10
            cntg, rowsg = m.neighbors()
            cnto, rowso = w.neighbors()
            mask = np.arange(rowso.shape[1])[None, :] < cnto[:, None]
            assert np.array_equal(cntg, cnto) and np.array_equal(rowsg[mask], rowso[mask])
            gg = m.ghosts()
            assert np.array_equal(gg["tag"], ao["tag"][nl:])
        m.close()
    # and the two halo implementations agree with each other bit for bit over a run (deterministic two-sided kernel)
    monkeypatch.setenv("MESO_PAIR_ONCE", "0")
    a, _ = make_pair(8, "sp")
    monkeypatch.setenv("MESO_FORCE_COMM_PATH", "0")
    b, _ = make_pair(8, "sp")
    a.setup(); b.setup()
    a.run(23); b.run(23)
    da, db = a.download(), b.download()
    for k in ("x", "v", "f", "tag"):
        assert np.array_equal(da[k], db[k]), k
    a.close(); b.close()
    # same in the default pair-once loop (bulk kernel || halo, then border kernel, both reducing into one accumulator): fp64, 1e-10
    monkeypatch.setenv("MESO_PAIR_ONCE", "1")
    monkeypatch.setenv("MESO_FORCE_COMM_PATH", "1")
    a, w = make_pair(8, "dp")
    a.setup(); w.setup()
    a.run(12); w.run(12)
    da, ao = a.download(), w.atoms()
    assert np.array_equal(da["tag"], ao["tag"][:ao["nlocal"]])
    assert np.abs(da["x"] - ao["x"][:ao["nlocal"]]).max() < 1e-10 and np.abs(da["v"] - ao["v"][:ao["nlocal"]]).max() < 1e-10
    a.close()


def test_sp_run_then_oracle_force_from_device_state():
    """fp32 trajectories cannot be locked (the RNG keys on fp32 velocity mantissa bits), so the
    multi-step fp32 path is checked by recomputing the LAST force on the oracle from the device's own state."""
    m, w = make_pair(10, "sp")
    m.setup()
    nsteps = 13
    m.run(nsteps)
    d = m.download()
    dtfm = 0.5 * 0.005
    v_half = d["v"] - dtfm * d["f"]                     # undo the fused final half-kick
    w.set_atoms(d["x"], v_half, tag=d["tag"], image=d["image"])
    w.ntimestep = nsteps
    w.rebuild(); w.force_clear(); w.pair_compute()
    ao = w.atoms()
    fo = np.empty_like(d["f"])
    fo[ao["tag"][:ao["nlocal"]] - 1] = ao["f"]
    fg = np.empty_like(d["f"])
    fg[d["tag"] - 1] = d["f"]
    # a handful of atoms may see a different RNG draw if (float)v_half rounds differently: allow 1e-4 of them
    derr = np.linalg.norm(fg - fo, axis=1) / np.maximum(np.linalg.norm(fo, axis=1), np.linalg.norm(fo, axis=1).mean())
    assert (derr > SP_TOL).mean() <= 1e-4, (derr > SP_TOL).sum()
    m.close()


def test_thermostat_equilibrium_statistics():
    """sigma^2 = 2 gamma kT: the device trajectory thermalises at T = 1 (stock LAMMPS: 1.000 +- 0.005)."""
    m, _ = make_pair(16, "sp")
    m.setup()
    m.run(600)
    ts = []
    for _ in range(20):
        m.run(20)
        ts.append(m.temperature())
    assert abs(np.mean(ts) - 1.0) < 0.02, np.mean(ts)
    d = m.download(("x", "v"))
    assert np.abs((d["v"]).sum(0)).max() < 1e-6 * len(d["v"])      # momentum conserved (pairwise antisymmetric forces)
    assert d["x"].min() >= -1.0 and d["x"].max() <= 17.0
    m.close()


# ------------------------------------------------------------------ the benchmarked sizes against the oracle (BASELINE configs[1], [2])
@pytest.mark.parametrize("L,precision", [(64, "sp"), (48, "dp")])
def test_full_size_setup_parity_against_the_oracle(L, precision):
    """sp.run case 64 (1,048,576 particles) and dp.run case 48 (442,368): everything the rebuild produces bit for bit
    against the C oracle -- sort keys, permutation, ghosts, cells, ORDERED neighbor rows, the tile-transposed table,
    packed coordinates and TEA signatures -- and the setup forces to 1e-5 (fp32) / 1e-12 (fp64); then size-independent
    properties of a short run (sum F = 0, temperature)."""
    m, w = make_pair(L, precision)
    m.setup()
    w.setup()
    assert_state_identical(m, w, precision=precision)
    c = m.counts()
    assert c["nlocal"] == 4 * L ** 3
    cnt = m.pair_count()
    # stock LAMMPS on 25.data: 17.92 half-list neighbors per atom at r_n = 1.3 -> 35.84 full
    assert abs(cnt.mean() - 35.84) < 0.2, cnt.mean()
    d = m.download(("f",))
    assert np.abs(d["f"].sum(0)).max() < (2e-2 if precision == "sp" else 1e-8) * np.sqrt(len(cnt))   # sum F = 0
    m.run(10)
    T = m.temperature()
    assert 0.5 < T < 2.5
    m.close()


# ------------------------------------------------------------------ bead-spring polymers
def make_polymer_pair(L, precision, lj12=1.0, monkeypatch=None, once="1"):
    x, typ, tag, nb, bt, ba = workload.polymer_melt(L, chain_len=8, seed=5)
    v = workload.maxwell_velocities(len(x))
    coeff = [(1, 1, (25, 4.5, 3.0, 1.0, 1.0)), (1, 2, (40, 4.5, 3.0, 1.0, 1.0)), (2, 2, (25, 4.5, 3.0, 1.0, 1.0))]
    m, w = make_pair(L, precision, ntypes=2, mass=[0.0, 1.0, 1.0], coeff=coeff, types=typ, x=x, v=v, tag=tag)
    m.bond_style("harmonic/meso", 1)
    m.bond_coeff(1, 50.0, 0.5)
    m.special_bonds(lj12)
    m.bonds(nb, bt, ba)
    w.set_bonds(nb, bt, ba, tag=tag, k=[0.0, 50.0], r0=[0.0, 0.5], special_lj12=lj12)
    return m, w


@pytest.mark.parametrize("lj12", [1.0, 0.0])
def test_polymer_setup_forces_table_and_bond_energy(lj12):
    """atom_style dpd/bond/meso + bond_style harmonic/meso: bond table rides the reorder, tag map, harmonic bonds,
    1-2 exclusion filter: rows (order included) bit-exact, forces and bond energy to the fp64 bar."""
    for precision in ("dp", "sp"):
        m, w = make_polymer_pair(8, precision, lj12)
        m.setup(eflag=1, vflag=1); w.setup(eflag=1, vflag=1)
        assert_state_identical(m, w, precision=precision, tol=2e-5 if precision == "sp" else 1e-11)
        eg, eo = m.bond_energy(), w.bond_energy()
        assert abs(eg - eo) < 1e-10 * abs(eo), (eg, eo)
        vg, _ = m.per_atom_virial()
        vo, _ = w.virial()
        assert np.abs(vg - vo).max() < (1e-3 if precision == "sp" else 1e-8)
        m.close()


@pytest.mark.parametrize("once", ["1", "0"])
def test_polymer_trajectory_fp64_lockstep(monkeypatch, once):
    """12 steps / two rebuilds of bead-spring chains in solvent: both run loops (pair-once + REDG, and the two-sided
    kernel accumulating into f) stay on the oracle's fp64 trajectory; bond tables survive the reorder."""
    monkeypatch.setenv("MESO_PAIR_ONCE", once)
    m, w = make_polymer_pair(8, "dp")
    m.setup(); w.setup()
    m.run(12); w.run(12)
    ag, ao = m.download(), w.atoms()
    nl = ao["nlocal"]
    assert np.array_equal(ag["tag"], ao["tag"][:nl])
    assert np.abs(ag["x"] - ao["x"][:nl]).max() < 1e-9 and np.abs(ag["v"] - ao["v"][:nl]).max() < 1e-9
    assert force_err(ag["f"], ao["f"]) < 1e-9
    # and in fp32: chains stay bonded (no bond stretches beyond a few r0) and the thermostat holds T
    m.close()
    m, _ = make_polymer_pair(8, "sp")
    m.setup()
    m.run(200)
    d = m.download(("x", "tag"))
    assert 0.5 < m.temperature() < 1.8                           # still cooling from the stretched start; fp32 atomics: +-0.1 run to run
    m.bond_compute(1, 0)
    assert m.bond_energy() < 50.0 * 0.5 * 4 * 8 ** 3            # mean bond extension well below r0
    m.close()


def test_bonds_refused_on_a_decomposition_and_missing_partner_flagged():
    x, typ, tag, nb, bt, ba = workload.polymer_melt(6, chain_len=4, seed=9)
    m, _ = make_pair(6, "dp", ntypes=2, mass=[0.0, 1.0, 1.0], types=typ, x=x, tag=tag,
                     coeff=[("*", "*", (25, 4.5, 3.0, 1.0, 1.0))])
    m.bond_style("harmonic/meso", 1)
    m.bond_coeff(1, 50.0, 0.5)
    ba2 = ba.copy()
    ba2[np.flatnonzero(nb > 0)[0], 0] = len(x) + 7            # a partner tag nobody owns
    m.bonds(nb, bt, ba2, tag_max=len(x) + 8)
    with pytest.raises(MesoError, match="Bond atoms missing"):
        m.setup()
        m.sync()
    m.close()
