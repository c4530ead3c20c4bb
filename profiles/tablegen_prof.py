"""One complete asymmetric-top table (nmv_prop/a-run argument list, theta = 0..180) for ncu captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

gpu = ge.load_package().gpu
r, e, q, info = gpu.gen_asymrho(0.37, 128, -1, 0, 180, 0.6666525, 0.2306476, 0.1769383, 66)
print("device ms (eigen+coeff, phi, chi gemm, combine):", gpu.gen_timing(), "rho(identity) =", r[0, 0, 0])
