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
