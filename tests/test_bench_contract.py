"""CPU-side checks of bench.py: the processor grids it picks, the host-core stand-in for `mpirun -np P`, and the JSON
contract of the reference arm (which runs without a GPU: stock LAMMPS from oracle/_ref, else the oracle port)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_procgrid_matches_lammps_choice_for_cubic_boxes():
    assert bench.procgrid_for(1) == (1, 1, 1) and bench.procgrid_for(2) == (1, 1, 2)
    assert bench.procgrid_for(4) == (1, 2, 2) and bench.procgrid_for(8) == (2, 2, 2)
    for n in (1, 2, 3, 4, 6, 8, 12, 16, 64):
        g = bench.procgrid_for(n)
        assert g[0] * g[1] * g[2] == n and g[0] <= g[1] <= g[2]


def test_host_copies_is_a_power_of_two_with_whole_bricks():
    for L in (8, 25, 48, 64, 100):
        p = bench.host_copies(L)
        assert p >= 1 and (p & (p - 1)) == 0 and p <= 64
        assert all(L % g == 0 and L // g >= 4 for g in bench.procgrid_for(p))
    assert bench.host_copies(25) == 1                    # 25 has no even split


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--case", "8", "--steps", "4", "--warmup", "1",
                          "--ref-steps", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [s for s in out.stdout.split("\n") if s.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/s" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4 and "workload" in d["config"]


def test_workload_files_round_trip(tmp_path):
    """the synthetic data files have the layout of example/simple/25.data and survive a write -> read round trip bit for bit
    (positions are rounded to the %.9e of the file when they are generated)"""
    import numpy as np
    from meso_b200 import workload
    x = workload.dpd_fluid(5)
    assert x.shape == (500, 3) and x.min() >= 0 and x.max() < 5
    cells = np.floor(x).astype(int)
    assert (np.bincount(cells[:, 0] + 5 * (cells[:, 1] + 5 * cells[:, 2]), minlength=125) == 4).all()   # 4 atoms per unit cell, in cell order
    p = str(tmp_path / "5.data")
    workload.write_data(p, x, 5)
    x2, tag, typ, lo, hi, mass = workload.read_data(p)
    assert np.array_equal(x2, x) and np.array_equal(tag, np.arange(1, 501)) and (typ == 1).all() and hi == [5.0, 5.0, 5.0] and mass == [0.0, 1.0]
    v = workload.maxwell_velocities(500)
    assert abs((v * v).sum() / (3 * 500 - 3) - 1.0) < 1e-12 and np.abs(v.sum(0)).max() < 1e-10
    # bead-spring writer: every bond once, both partners hold it in the per-atom table
    xp, typ, tag, nb, bt, ba = workload.polymer_melt(4, chain_len=4, seed=1)
    nbonds = workload.write_data_bond(str(tmp_path / "p.data"), xp, 4, typ, 2, nb, bt, ba)
    assert 2 * nbonds == nb.sum()
    txt = open(str(tmp_path / "p.data")).read()
    assert "%d bonds" % nbonds in txt and "Bonds" in txt and "1 bond types" in txt
    xa, typa, _, nba, _, baa = workload.amphiphilic_channel(6)
    assert set(np.unique(typa)) == {1, 2, 3} and xa[:, 2].min() > 0.29 and xa[:, 2].max() < 5.71
