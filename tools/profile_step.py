"""Runs the bench workload for a few steps inside a cudaProfilerStart/Stop window (for ncu --profile-from-start off).
   python tools/profile_step.py [--case 64] [--precision sp] [--steps 10]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from meso_b200.engine import dpd_fluid_deck  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", type=int, default=64)
ap.add_argument("--precision", default="sp")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=20)
a = ap.parse_args()
m = dpd_fluid_deck(a.case, a.precision)
m.setup()
m.run(a.warmup)
m.sync()
m._chk(m.L.meso_profiler(m.h, 1))
m.run(a.steps)
m.sync()
m._chk(m.L.meso_profiler(m.h, 0))
print("T", m.temperature())
m.close()
