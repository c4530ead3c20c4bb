#!/usr/bin/env python
"""bench.py -- PIMC bead-updates/s (translational + rotational) of the B200 hot path.

  python bench.py --gpus N --steps K --warmup W [--workload C5] [--chains 8]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
  python bench.py --impl reference ...    the reference's own CPU hot path (oracle/_ref) on the host cores

One "step" = one Monte-Carlo PASS (P iterations of the reference's `time` loop, mc_main.cc:348-381)
of every chain resident on the GPU.  Bead-updates are counted with the formula of SURVEY.md 8(d):
bisection 2^L-1 per atom per reference call, whole-path move P per atom, rotational step 1.
Default workload: C5, synthetic N2O in (pH2)_100 at 0.5 K, P=1024, Q=128, 8 chains per GPU (BASELINE.json
configs[4], 64 chains over 8 GPUs); weak scaling, chains are independent, the only collective is the
NCCL all-reduce of the estimator accumulators at the end of the block.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

# algorithmic flops / operand bytes of the leaves (reference algorithm, every operand once)
LEAF_FLOPS = dict(spot1d=45, lpot2d=58, srot=33, gauss3=170, rotden=470, vcord=335, caleng=685)
LEAF_BYTES = dict(spot1d=72, lpot2d=88, vcord=56, caleng=48)


def algorithmic_work(s):
    """(flops, bytes) per PASS per chain of the reference algorithm, and the bead-updates per pass."""
    N, P, Q, R = s.N, s.P, s.Q, s.R
    mol = [t for t in s.types if t.molecule]
    atoms = [t for t in s.types if not t.molecule]
    na = atoms[0].numb if atoms else 0
    nm = mol[0].numb if mol else 0
    kind = mol[0].molecule if mol else 0
    cross = "lpot2d" if kind == 1 else "vcord"

    def partners(t):          # (flops, bytes) of the sum over partners of one bead of type t
        if t.molecule:
            f = na * LEAF_FLOPS[cross] + (nm - 1) * LEAF_FLOPS["caleng"]
            b = na * LEAF_BYTES[cross] + (nm - 1) * LEAF_BYTES["caleng"]
        else:
            f = (na - 1) * LEAF_FLOPS["spot1d"] + nm * LEAF_FLOPS[cross]
            b = (na - 1) * LEAF_BYTES["spot1d"] + nm * LEAF_BYTES[cross]
        return f, b

    flops = bytes_ = 0.0
    for t in s.types:
        L = t.levels
        pf, pb = partners(t)
        red = (2 ** (L + 1) - 2 - L) / (2 ** L - 1)                    # the reference's redundant level schedule
        nb = P * t.numb * (2 ** L - 1)
        flops += nb * (2 * pf * red + LEAF_FLOPS["gauss3"])
        bytes_ += nb * (72 + 2 * pb) * red
        flops += P * t.numb * 2 * pf                                    # whole-path move
        bytes_ += P * t.numb * (24 + 2 * pb)
        if t.molecule and Q:
            nrot = P * Q * t.numb
            dens = 4 * (LEAF_FLOPS["srot"] if kind == 1 else LEAF_FLOPS["rotden"])
            dens_b = 4 * (56 if kind == 1 else 32)
            flops += nrot * (dens + 2 * R * pf + 130)
            bytes_ += nrot * (72 + dens_b + 2 * R * pb + 24 * R + 48)
    return flops, bytes_, s.bead_updates_per_pass()


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def sample_size(s):
    """time steps per CPU-baseline sample: a whole pass would take minutes on one core for C5"""
    return max(1, min(s.P, 64))


def _reference_child(args):
    """one reference chain in this process (the reference keeps its state in globals): prints {"t_step", "t_mol"}"""
    from oracle import oracle_py as op
    pkg = ge.load_package()
    cfg = pkg.configs.make_config(args.workload)
    s = cfg.system
    threads = int(os.environ.get("OMP_NUM_THREADS", "1"))
    cpus = os.environ.get("PIMC_REF_CPUS")
    if cpus:
        try:
            os.sched_setaffinity(0, {int(c) for c in cpus.split(",")})
        except OSError:
            pass
    fast = op.ref_available(fast=True)
    sys.stdout.flush()
    saved = os.dup(1)                                              # the reference chats on stdout (cout); keep ours to one JSON line
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        R = op.Ref(cfg, nthreads=threads, fast=fast)
        R.lib.ref_run_steps.restype = __import__("ctypes").c_double
        nsamp = sample_size(s)
        t0 = 1                                                     # samples start after time 0: the whole-path sweep is
        for _ in range(args.warmup):                               # timed once, separately, and added pro rata below
            R.lib.ref_run_steps(t0, nsamp); t0 += nsamp
        t_mol = R.lib.ref_run_steps(0, 1)                          # one step at time == 0: molecular + 1 ordinary step
        tt = 0.0
        for _ in range(args.steps):
            tt += R.lib.ref_run_steps(t0, nsamp); t0 += nsamp
    finally:
        os.dup2(saved, 1)
    print(json.dumps({"t_step": tt / (args.steps * nsamp), "t_mol": t_mol, "fast": bool(fast)}))


def _spawn_reference(args, nproc, threads, cpus_of):
    """`nproc` concurrent reference chains, `threads` OpenMP threads each; returns the children's timing dicts"""
    import subprocess
    procs = []
    for k in range(nproc):
        env = {**os.environ, "PIMC_REF_CHILD": "1", "OMP_NUM_THREADS": str(threads), "OMP_PROC_BIND": "false", "RANK": "0", "WORLD_SIZE": "1"}
        if cpus_of:
            env["PIMC_REF_CPUS"] = ",".join(str(c) for c in cpus_of(k))
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", str(args.steps), "--warmup", str(args.warmup)]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env))
    out = []
    for pr in procs:
        o = pr.communicate(timeout=1500)[0]
        line = [ln for ln in o.splitlines() if ln.startswith("{")]
        if not line:
            return None
        out.append(json.loads(line[-1]))
    return out


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref: its unmodified C++ objects + the C++ restatement of
    its Fortran leaves) on the host cores, on a bounded sample of the SAME workload as the GPU arm: `chains` independent
    Markov chains run side by side, one process each (the reference keeps its state in globals), the cores divided evenly
    between them as OpenMP threads.  The single-chain figures with one thread and with every core are reported
    beside it."""
    if os.environ.get("PIMC_REF_CHILD"):
        return _reference_child(args)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as op
    pkg = ge.load_package()
    cfg = pkg.configs.make_config(args.workload)
    s = cfg.system
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if not (op.ref_available(fast=True) or op.ref_available()):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (needs /root/reference at build time)"}))
        return
    chains = args.chains or (8 if args.workload == "C5" else 148)
    nproc = min(chains, cores)
    threads = max(1, cores // nproc)
    allowed = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(cores))
    res = _spawn_reference(args, nproc, threads, lambda k: allowed[k * threads:(k + 1) * threads])
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "the reference processes produced no timing"}))
        return
    nsamp = sample_size(s)
    bu = s.bead_updates_per_pass()

    def pass_time(r):
        return r["t_step"] * s.P + max(0.0, r["t_mol"] - r["t_step"])        # a full pass = P ordinary steps + the time-0 extras
    t_pass = max(pass_time(r) for r in res)                                    # the slowest of the concurrent chains
    # `chains` chains on `nproc` concurrent processes: ceil(chains / nproc) rounds of the measured pass time
    rounds = (chains + nproc - 1) // nproc
    value = bu["total"] * chains / (t_pass * rounds)
    single = {}
    import copy
    short = copy.copy(args)
    short.steps, short.warmup = min(args.steps, 3), min(args.warmup, 1)       # side figures: a shorter sample
    for label, th in (("omp1", 1), ("omp_all", cores)):
        r1 = _spawn_reference(short, 1, th, None)
        single[label] = bu["total"] / pass_time(r1[0]) if r1 else None
    fast = res[0].get("fast")
    line = {"metric": "pimc_bead_updates_per_sec", "value": value, "unit": "bead-updates/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_pass * rounds, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, s, chains), "chains_per_gpu": chains, "sample_time_steps": nsamp,
                       "processes": nproc, "omp_threads_per_process": threads},
            "cpu_baseline": {"value": value, "unit": "bead-updates/s", "cores": cores, "kind": "reference",
                             "single_chain_omp1": single["omp1"], "single_chain_omp_all": single["omp_all"],
                             "sample": f"{chains} independent chains as {nproc} concurrent processes x {threads} OpenMP threads; each: {args.steps} x {nsamp} iterations of "
                                       f"the time loop (mc_main.cc:349-381) + one time-0 step, scaled to a pass; slowest process counts; "
                                       f"{'-Ofast' if fast else '-O2'} build of the reference objects"},
            "e2e": {"value": value, "unit": "bead-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(w, s, chains):
    wm = f" (WORM {s.worm[0]} {s.worm[1]} {s.worm[2]})" if getattr(s, "worm", None) else " (worm off)"
    names = {"C1": "examples/MF_1He_0.37K_512_128", "C2": "examples/MF_8He_0.37K_512_128" + wm,
             "C3": "examples/SO2_4pH2_0.37K_1024_256" + wm, "C4": "examples/H2Odimer_0.74K_4096_2048",
             "C5": "synthetic N2O in (pH2)_100 at 0.5 K"}
    return f"{w}: {names[w]}, N={s.N}, P={s.P}, Q={s.Q}, {chains} chains/GPU"


def cpu_baseline(pkg, args, cfg):
    """Bounded CPU sample in a subprocess (the reference keeps global state and must not share our CUDA process)."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "3", "--warmup", "1"]
    if args.chains:
        cmd += ["--chains", str(args.chains)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
        for ln in reversed(out.stdout.splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                return d.get("cpu_baseline", {"unavailable": d.get("unavailable")})
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}
    return {"unavailable": "no output"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="C5")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: 8 for C5, 148 otherwise)")
    ap.add_argument("--cpc", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--team", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--worm", action="store_true", help="keep the deck's WORM line (C2, C3): exchange sampled with the worm algorithm")
    ap.add_argument("--tables", default="synthetic", choices=["synthetic", "generated"],
                    help="C1-C4: 'generated' replaces the analytic stand-in of the rho/E/E2 tables (SURVEY 8d fallback) by the tables the "
                         "device generator makes from the molecule's rotational constants (asymrho, 8 printed digits)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the first communicator is created; the contract is ONE JSON line on stdout
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    pkg = ge.load_package()
    cfg = pkg.configs.make_config(args.workload, worm=args.worm)
    s = cfg.system
    tables_note = "shipped .rot file" if "rotlin" in cfg.tables else "analytic stand-in (SURVEY 8d)"
    if args.tables == "generated" and "rot3d" in cfg.tables:
        mol = [t.name for t in s.types if t.molecule == 2][0]
        A, B, C = pkg.configs.ROT_CONSTANTS[mol]
        maxj = pkg.gpu.asym_auto_maxj(s.temperature, s.Q, A, B, C)
        r3, e3, q3, _ = pkg.gpu.gen_asymrho(s.temperature, s.Q, -1, 0, 180, A, B, C, maxj)
        cfg.tables["rot3d"] = tuple(x.reshape(-1) for x in (r3, e3, q3))
        tables_note = f"generated on the device: asymrho {s.temperature} {s.Q} -1 0 180 {A} {B} {C} {maxj} ({pkg.gpu.gen_timing().sum():.1f} ms)"
    chains = args.chains or (8 if args.workload == "C5" else 148)
    G = pkg.gpu.PimcGpu(cfg, nchains=chains, chain_offset=rank * chains, device=local, ctas_per_chain=args.cpc,
                        threads_per_cta=args.threads, team=args.team)
    G.seed((12345,) * 6)
    stream = torch.cuda.ExternalStream(G.L.pimcgpu_stream(), device=local)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")          # 256 MB > 126 MB L2

    class Acc:          # zero-copy torch view of the library's accumulator buffer for the NCCL all-reduce
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    lay = G.accum_layout()
    acc_t = torch.as_tensor(Acc(G.L.pimcgpu_accum_device_ptr(), lay["n_total"]), device="cuda")

    P = s.P
    flops_pass, bytes_pass, bu = algorithmic_work(s)
    per_step_units = bu["total"] * chains                     # bead-updates per step on this GPU

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K passes, CUDA events on the library's stream, L2 flushed between passes ----
    for _ in range(args.warmup):
        G.steps(P)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t_kernel = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        G.steps(P, sync=False)
        e1.record(stream)
        e1.synchronize()
        G.sync()
        t_kernel += e0.elapsed_time(e1) * 1e-3
    barrier()
    if dist:
        t = torch.tensor([t_kernel], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_kernel = float(t.item())
    value = per_step_units * world * args.steps / t_kernel

    # ---- end to end through the C ABI with host buffers: upload -> pass -> estimators -> (all-reduce) -> download, every step ----
    # Host state lives in pinned buffers allocated once, as the reference's MCCoords / MCAngles do (mc_setup.cc:135-163).  Two sets
    # of chains (A, B) take turns on the device, the way a driver time-shares one GPU between two jobs: while set A's pass runs,
    # set B's previous result travels to the host and its next input to the device on the library's copy stream (split-phase
    # pimcgpu_upload_states_begin/_commit, pimcgpu_download_states_begin/_end).  Every step still uploads its own input from host
    # memory and downloads its own result; nothing is skipped, the copies just do not stall the move kernel.
    n_beads = s.N * P
    pins = [(torch.empty((chains, 3, n_beads), dtype=torch.float64, pin_memory=True), torch.empty((chains, 3, n_beads), dtype=torch.float64, pin_memory=True)) for _ in range(2)]
    host = [(c.numpy(), a.numpy()) for c, a in pins]
    for hc, ha in host:
        G.download_all_into(hc, ha)
    perm = None
    if cfg.perm is not None:
        perm = np.ascontiguousarray(np.tile(np.ascontiguousarray(cfg.perm, dtype=np.int32), (chains, 1)))
    pin_acc = torch.empty(lay["n_total"], dtype=torch.float64, pin_memory=True)
    host_acc = pin_acc.numpy()
    h2d = chains * ((P * 3 * ((s.N + 3) // 4 * 4) + 2 * max(1, s.Q) * 3 * max(1, sum(t.numb for t in s.types if t.molecule))) * 8 + (3 * s.N + 3) * 4)
    d2h = h2d - chains * (3 * s.N + 3) * 4 + lay["n_total"] * 8        # beads, rotor angle / axis rows, accumulators

    trace = {} if os.environ.get("PIMC_E2E_TRACE") else None       # host wall time per call of the loop (rank 0, stderr)

    def timed(name, fn, *a):
        if trace is None:
            return fn(*a)
        t0 = time.perf_counter()
        r = fn(*a)
        trace[name] = trace.get(name, 0.0) + time.perf_counter() - t0
        return r

    def e2e_steps(nsteps):
        G.upload_begin(host[0][0], host[0][1], perm)
        for k in range(nsteps):
            cur, oth = k % 2, 1 - k % 2
            timed("upload_commit", G.upload_commit)                 # this step's input (set `cur`) into the state
            timed("accum_reset", G.accum_reset)
            timed("steps", G.steps, P, False)
            timed("measure", G.measure)
            G.L.pimcgpu_accum_device_ptr()
            if dist:
                def reduce_acc():
                    torch.cuda.current_stream().wait_stream(stream)
                    dist.all_reduce(acc_t)
                    stream.wait_stream(torch.cuda.current_stream())
                timed("all_reduce", reduce_acc)
            timed("accum_begin", G.accum_download_begin, host_acc)  # the step's sums travel first, ahead of its configuration
            # the device is busy with the pass from here on: finish the previous step's download, send the next step's input
            if k > 0:
                timed("download_end", G.download_end)               # set `oth` (result of the previous step) is on the host
            timed("upload_begin", G.upload_begin, host[oth][0], host[oth][1], perm)     # next step's input starts travelling now
            timed("download_begin", G.download_begin, host[cur][0], host[cur][1])       # this step's result: snapshot, then D2H on the copy stream
            timed("accum_end", G.accum_download_end)                                    # this step's estimator sums are on the host (synchronises the pass)
        G.download_end()
        G.upload_commit()                                           # the input that was sent ahead for a step that will not run

    e2e_steps(2)                                                    # warm the split-phase path (buffers, streams)
    if trace is not None:
        trace.clear()
    barrier()
    w0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - w0
    if trace is not None:
        print(f"[e2e trace rank {rank}] ms per step: " + json.dumps({k: round(1e3 * v / args.steps, 3) for k, v in trace.items()}) + f" total {1e3 * t_e2e / args.steps:.3f}", file=sys.stderr)
    if dist:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_value = per_step_units * world * args.steps / t_e2e

    if rank == 0:
        peaks, how = measured_peaks()
        fp64_peak = pkg.gpu.fp64_peak_tflops()
        ach_tf = flops_pass * chains * args.steps / t_kernel / 1e12          # per GPU (dominant kernel = the whole step)
        ach_gb = bytes_pass * chains * args.steps / t_kernel / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        line = {
            "metric": "pimc_bead_updates_per_sec", "value": value, "unit": "bead-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_kernel / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, s, chains), "rotor_tables": tables_note, "chains_per_gpu": chains, "step": "one MC pass = P iterations of the time loop",
                       "l2": "flushed (256 MB write) between timed passes; the state is L2/SMEM-resident by design inside a pass",
                       "parallelism": f"independent chains x{world} GPUs, NCCL all-reduce of accumulators per block",
                       "geometry": G.geometry()},
            "e2e": {"value": e2e_value, "unit": "bead-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps,
            "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                         "traffic": traffic, "kernel": "pimc_steps_kernel (one launch = one pass of all chains)",
                         "peak_source": "pimcgpu_fp64_peak DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                         "hbm": {"achieved": ach_gb, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_gb / peaks["hbm_gbs"], "of": how}},
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(pkg, args, cfg)
        print(json.dumps(line))
    G.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
