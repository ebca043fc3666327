"""GPU experiment / ncu target: guest-guest pair energies (ceg_guest_env_energy), 10^5 trial CO2 x 10^3 CO2 guests."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl
import bench
r = bench.bench_guest(ceg, wl, 0)
print(r["ms_per_step"], "ms per call; kernel", r["kernel_ms"], "ms;", r["pair_tests_per_s_kernel"] / 1e9, "G pair tests/s (kernel); frac", r["roofline"]["frac"], "checksum", r["checksum_K"])
