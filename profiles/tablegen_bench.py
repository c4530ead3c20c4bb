"""Times the device rho-table generators on the reference's own argument lists (nmv_prop/a-run, symtop_prop/a-run,
linear_prop/README) for the COMPLETE tables (theta = 0..180) and prints one JSON line.  Device times are CUDA-event
times inside the library (pimcgpu_gen_timing); wall times include allocation and the device->host copy of the tables."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

gpu = ge.load_package().gpu
args = (0.37, 128, -1, 0, 180, 0.6666525, 0.2306476, 0.1769383, 66)
gpu.gen_asymrho(*args[:3], 0, 1, *args[5:])          # warm-up (context, module load)
best = None
for rep in range(3):
    t = time.time()
    r, e, q, info = gpu.gen_asymrho(*args)
    wall = time.time() - t
    ms = gpu.gen_timing()
    if best is None or ms.sum() < best[1].sum():
        best = (wall, ms.copy())
wall, ms = best
maxj = 66
np_ = ((maxj + 1 + 3) // 4) * 4
K = 2 * np_
rows = 181 * 3 * 361
gemm_flop = 2.0 * 2 * rows * K * 384                  # both parity classes, padded N
useful_flop = 2.0 * rows * (2 * 67 + 2 * 66) * 361
wall_sym = []
for rep in range(3):
    t = time.time()
    rs, es, qs, _ = gpu.gen_symrho(0.37, 128, 1, 0, 180, 0.5, 0.3, 66)
    wall_sym.append(time.time() - t)
t = time.time()
out, _ = gpu.gen_linden(0.5, 128, 1.92253, 1500, -1)
wall_lin = time.time() - t
t = time.time()
gpu.write_e15_8("/tmp/tg_full.rho", r)
wall_write = time.time() - t
print(json.dumps({
    "asymrho_full_table": {"args": "0.37 128 -1 0 180 0.6666525 0.2306476 0.1769383 66", "values": int(3 * r.size),
                           "device_ms": {"eigen+coeff": ms[0], "phi": ms[1], "chi_gemm": ms[2], "combine": ms[3], "total": float(ms.sum())},
                           "wall_s_incl_d2h": wall, "chi_gemm_tflops_padded": gemm_flop / (ms[2] * 1e-3) / 1e12,
                           "chi_gemm_tflops_useful": useful_flop / (ms[2] * 1e-3) / 1e12,
                           "fp64_peak_tflops": gpu.fp64_peak_tflops(), "rho000": float(r[0, 0, 0]), "Ztau_over_8pi2": float(info[13] / (8 * np.pi ** 2))},
    "symrho_full_table_wall_s": wall_sym, "linden_1500pt_wall_s": wall_lin,
    "write_e15_8_one_table_s": wall_write, "write_bytes": os.path.getsize("/tmp/tg_full.rho")}))
