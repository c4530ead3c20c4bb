"""world_size-2 gloo test of the N>1 host logic: chains are sharded rank-major (chain = rank*C_local + i) with
disjoint MRG32k3a streams, every rank accumulates SUMS and COUNTS, one all-reduce(sum) forms the block totals and
ratios are taken afterwards -- the result must equal the single-process run over all chains."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SEED = (12345, 23456, 34567, 45678, 56789, 67890)
C_LOCAL, NSTEPS = 2, 40


def chain_accumulators(chain_global):
    """[count, kin, pot, rot, rotsq, mc totals(6), mc accepted(6)] of one chain after NSTEPS steps (oracle replay)."""
    import __graft_entry__ as ge
    from oracle import oracle_py as op
    pkg = ge.load_package()
    cfg = pkg.configs.make_config("C5", P=32, Q=8, nsolv=4)
    O = op.Oracle(cfg)
    O.sched_seed(SEED, chain_global)
    O.sched_run(0, NSTEPS)
    srot, esq, _ = O.get_rot_energy()
    tot, acc = O.counters()
    return np.r_[1.0, O.get_kin(), O.get_pot(0), srot, esq, tot.reshape(-1), acc.reshape(-1)]


def worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = sum(chain_accumulators(rank * C_LOCAL + i) for i in range(C_LOCAL))
    t = torch.from_numpy(local.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process(tmp_path):
    out = str(tmp_path / "acc.npy")
    mp.start_processes(worker, args=(2, 29531 + os.getpid() % 200, out), nprocs=2, join=True, start_method="spawn")
    got = np.load(out)
    want = sum(chain_accumulators(c) for c in range(2 * C_LOCAL))
    assert got[0] == 2 * C_LOCAL
    assert np.allclose(got, want, rtol=1e-13, atol=0)
    # different chains really are different trajectories (disjoint streams)
    a, b = chain_accumulators(0), chain_accumulators(1)
    assert not np.allclose(a[1:5], b[1:5])
    # ratios after the reduction: acceptance = accepted / total over all ranks
    tot, acc = got[5:11], got[11:17]
    ratio = acc[tot > 0] / tot[tot > 0]
    assert np.all((ratio >= 0) & (ratio <= 1))
