"""Oracle worm port vs the reference's own mc_qworm.cc objects (oracle/_ref), one configuration per process.

Run as ``python tests/ref_compare_worm.py <C2|C3|C5> [--golden OUT.npz]``.  Both sides are fed the same explicit
uniforms per SPRNG stream (mc_randg.cc:90-174) and must stay BIT-IDENTICAL through a long random sequence of
MCWormMove calls (open / close / advance / recede / swap), including the world-line masks of the PotEnergy variants
while the worm is open.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as op  # noqa: E402

SMALL = dict(C2=dict(P=32, Q=8, nsolv=5), C3=dict(P=32, Q=8), C5=dict(P=32, Q=8, nsolv=6))
WORM = dict(C2=("He4", 0.13, 8), C3=("H2", 0.35, 8), C5=("H2", 0.35, 6))
STREAMS = (1, 2, 3, 4, 5, 8, 9, 10, 11, 12)


def make(name):
    cfgs = op._configs()
    cfg = cfgs.make_config(name, **SMALL[name])
    cfg.system.worm = WORM[name]
    # a compact cluster so that exchange (swap) moves have partners within reach
    s = cfg.system
    nb = s.types[0].numb
    c = cfg.coords.reshape(3, s.N, s.P).copy()
    com = c[:, :nb, :].mean(axis=(1, 2), keepdims=True)
    c[:, :nb, :] = com + 0.8 * (c[:, :nb, :] - com)
    cfg.coords = np.ascontiguousarray(c.reshape(3, -1))
    return cfg


def run(name, with_ref=True, ncalls=400):
    cfg = make(name)
    s = cfg.system
    O = op.Oracle(cfg)
    O.worm_init(0, s.worm[1], s.worm[2])
    R = op.Ref(cfg) if with_ref else None
    if with_ref:
        assert R.lib.ref_worm_enabled()
        R.queue_mode(True)
    rng = np.random.default_rng(2024)
    nb = s.types[0].numb
    N, P = s.N, s.P
    mism = {"state": 0, "coords": 0.0, "perm": 0, "consumed": 0, "pot_mask": 0.0, "counters": 0.0, "world_line": 0}
    seen_open = seen_swap = 0
    for call in range(ncalls):
        u = {k: rng.random(400) for k in STREAMS}
        O.worm_clear()
        for k in STREAMS:
            O.worm_push(k, u[k])
        O.worm_op(7)
        if with_ref:
            R.lib.ref_rng_clear()
            for k in STREAMS:
                R.push(k, u[k])
            R.lib.ref_worm_op(7)
            st = (op.C.c_int * 5)(); R.lib.ref_worm_get(st)
            mism["state"] += int(list(st) != O.worm_get())
            mism["consumed"] += sum(int(R.lib.ref_rng_pending(k) != O.worm_pending(k)) for k in STREAMS)
        st = O.worm_get()
        seen_open += st[0]
        if call % 20 == 0 or call == ncalls - 1:
            co, ao, _ = O.get_state()
            po, ro = O.get_perm(nb)
            if with_ref:
                cr, ar, _ = R.get_state()
                mism["coords"] = max(mism["coords"], float(np.abs(co - cr).max()))
                pr, rr = np.zeros(nb, dtype=np.int32), np.zeros(nb, dtype=np.int32)
                R.lib.ref_get_perm(op._ip(pr), op._ip(rr), nb)
                mism["perm"] += int(not (np.array_equal(po, pr) and np.array_equal(ro, rr)))
            if st[0] and with_ref:
                # world-line mask and masked potential sums in the G sector
                for a in range(nb):
                    for it in range(P):
                        mism["world_line"] += int(O.world_line(a, it) != bool(R.lib.ref_world_line(a, it)))
                pe = np.array([[O.pot_energy_it(a, it) for it in range(P)] for a in range(N)])
                pr_ = np.array([[R.lib.ref_PotEnergy_it(a, it) for it in range(P)] for a in range(N)])
                mism["pot_mask"] = max(mism["pot_mask"], float(np.abs(pe - pr_).max()))
                pp = np.array([O.pot_energy_path(a) for a in range(N)])
                pq = np.array([R.lib.ref_PotEnergy_path(a) for a in range(N)])
                mism["pot_mask"] = max(mism["pot_mask"], float(np.abs(pp - pq).max()))
    t, a, cq = O.worm_counters()
    if with_ref:
        tr, ar_, cr_ = np.zeros(7), np.zeros(7), op.C.c_double()
        R.lib.ref_worm_counters(op._dp(tr), op._dp(ar_), op.C.byref(cr_))
        mism["counters"] = float(max(np.abs(t - tr).max(), np.abs(a - ar_).max(), abs(cq - cr_.value)))
    po, ro = O.get_perm(nb)
    co, _, _ = O.get_state()
    res = dict(mism)
    res["open_accepted"] = float(a[0]); res["close_accepted"] = float(a[1]); res["advance_accepted"] = float(a[4])
    res["recede_accepted"] = float(a[5]); res["swap_accepted"] = float(a[6])
    res["calls_in_G_sector_accepted"] = seen_open
    gold = dict(final_coords=co, final_perm=po, counters_total=t, counters_accepted=a, final_worm=np.array(O.worm_get()))
    return res, gold


def main():
    name = sys.argv[1]
    golden = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--golden" else None
    res, gold = run(name, with_ref=True)
    print("RESULT " + json.dumps(res))
    if golden:
        np.savez_compressed(golden, **gold)


if __name__ == "__main__":
    main()
