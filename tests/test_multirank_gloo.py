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


# ---- the product's own sharding logic (no GPU needed): libpimcgpu.so's host-side stream addressing -----------------------
def _product_stream_state(seed, stream):
    import ctypes as C
    import __graft_entry__ as ge
    pkg = ge.load_package()
    L = C.CDLL(pkg.gpu.LIB)
    st = (C.c_ulong * 6)()
    assert L.pimcgpu_host_stream_state((C.c_ulong * 6)(*seed), C.c_long(stream), st) == 0
    return [int(v) for v in st]


def _mrg_draws_from_state(state, n):
    """MRG32k3a (rngstream.cc:242-265) in exact integer arithmetic from a 6-word state"""
    m1, m2 = 4294967087, 4294944443
    s = list(state)
    out = []
    for _ in range(n):
        p1 = (1403580 * s[1] - 810728 * s[0]) % m1
        s[0], s[1], s[2] = s[1], s[2], p1
        p2 = (527612 * s[5] - 1370589 * s[3]) % m2
        s[3], s[4], s[5] = s[4], s[5], p2
        out.append(((p1 - p2) if p1 > p2 else (p1 - p2 + m1)) * (1.0 / (m1 + 1)))
    return np.array(out)


def shard_worker(rank, world, port, out, S):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what pimcgpu_seed does on every rank: chain_offset = rank * C_local, stream (chain_offset + i) * S + s
    mine = np.array([[_product_stream_state(SEED, (rank * C_LOCAL + i) * S + s) for s in (0, 1, S - 1)] for i in range(C_LOCAL)], dtype=np.int64)
    gathered = [torch.zeros_like(torch.from_numpy(mine)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))
    if rank == 0:
        np.save(out, np.stack([g.numpy() for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_chain_sharding_uses_disjoint_product_streams(tmp_path):
    """Row (e): global chain c = rank * C_local + i owns streams c*S .. c*S+S-1 (S = P + Q + 8).  The stream states every rank
    derives with the PRODUCT's host function (what pimcgpu_seed uploads, csrc/pimcgpu.cu:stream_state) are gathered over gloo:
    together they are exactly the states one process derives for chains 0 .. 2C-1, all distinct, and the uniforms they generate
    are those of the reference's RngStream objects (oracle restatement of rngstream.cc)."""
    from oracle import oracle_py as op
    S = 32 + 8 + 8
    out = str(tmp_path / "states.npy")
    mp.start_processes(shard_worker, args=(2, 29731 + os.getpid() % 200, out, S), nprocs=2, join=True, start_method="spawn")
    got = np.load(out)                                            # [rank][local chain][3 streams][6]
    assert got.shape == (2, C_LOCAL, 3, 6)
    seen = set()
    for rank in range(2):
        for i in range(C_LOCAL):
            c = rank * C_LOCAL + i
            for k, s in enumerate((0, 1, S - 1)):
                want = _product_stream_state(SEED, c * S + s)
                assert list(got[rank, i, k]) == want
                assert tuple(want) not in seen
                seen.add(tuple(want))
                ref = op.mrg_draws(SEED, c * S + s, 1, 16)[0]     # the (c*S+s)-th RngStream after SetPackageSeed
                assert np.array_equal(_mrg_draws_from_state(want, 16), ref)
