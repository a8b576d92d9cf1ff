#!/usr/bin/env python
"""bench.py -- ensemble adjoint trajectories/sec on BASELINE.json's headline workload (config C2):
Lorenz d=3 P=3, N=65536 members per GPU, GaussAdjoint, Tsit5 fixed dt=0.01, T=10 (S=1000 steps), saveat 0.1 (K=101),
cotangent dgdu = u - 2, fp64, shared p, synthetic u0 = [1,0,0] + 0.1 z.

One "step" = one full gradient evaluation of the ensemble: forward solve + fused reverse adjoint pass + dG/dp
reduction (+ one all-reduce of dp over ranks for N>1; members are sharded, weak scaling: 65536 per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--members M]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

`value`   : members/s with inputs resident in HBM (device pointers through the C ABI), CUDA-event timed, max over ranks.
`e2e`     : same metric through the public API (solve + adjoint_sensitivities) with HOST buffers: pinned u0 H2D and
            du0/dp D2H inside the timed region every step.
`roofline`: the reverse kernel (dominant) against the measured HBM peak, algorithmic bytes = 122.4 B per member-step
             x N x S per launch.
`cpu_baseline`: the C oracle (a PORT of the reference algorithm; Julia cannot run here) on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(family="lorenz", sensealg="gauss", stepper="tsit5_fixed", T=10.0, dt=0.01, nsave=101,
                members_per_gpu=65536, cost=(1.0, -2.0), seed=20260923)
ALG_BYTES_PER_MEMBER_STEP = 8 * (3 + 2 * 6 + 101.0 / 1000.0 * 3)   # 122.4 B


def make_inputs(N, offset=0):
    rng = np.random.Generator(np.random.Philox(key=WORKLOAD["seed"] + offset))
    u0 = np.array([1.0, 0.0, 0.0])[:, None] + 0.1 * rng.standard_normal((3, N))
    p = np.array([10.0, 28.0, 8.0 / 3.0])
    return np.ascontiguousarray(u0), p


def host_cores():
    """cores this process may actually run on (the box reports 128 CPUs but the cgroup/affinity mask is smaller)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:                                                     # cgroup v2 CPU quota (the GPU boxes show 128 CPUs but grant fewer)
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(np.ceil(int(quota) / int(period)))))
    except Exception:
        try:                                                 # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = max(1, min(n, int(np.ceil(q / per))))
        except Exception:
            pass
    return n


_BEST_THREADS = None


def best_threads():
    """OpenMP thread count that actually maximises the oracle's throughput on this host (affinity masks and CPU quotas
    differ between boxes: probe 8 .. host_cores() on a small sample, keep the best)."""
    global _BEST_THREADS
    if _BEST_THREADS is None:
        cap = host_cores()
        cands = sorted({c for c in (4, 8, 16, 32, 64, 128, cap) if c <= cap}) or [1]
        cpu_oracle_rate(256, cands[0])                        # load the library
        rates = {c: cpu_oracle_rate(max(1024, 16 * c), c, repeats=2)[0] for c in cands}
        _BEST_THREADS = max(rates, key=rates.get)
        sys.stderr.write(f"[bench] oracle threads probe: {rates} -> {_BEST_THREADS}\n")
    return _BEST_THREADS


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the reverse kernel from the committed ncu summary (per launch)"""
    path = os.path.join(ROOT, "profiles", "r1_reverse_ncu_summary.json")
    try:
        with open(path) as f:
            m = json.load(f)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        r, w = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        return r["value"] * scale[r["unit"]] + w["value"] * scale[w["unit"]]
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, f"/tmp/b200adj_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """t_begin/t_end: wall-clock (time.time()) bounds of the timed region; samples outside are dropped."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                if t_begin is not None:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t_begin - 0.02 or ts > t_end + 0.02:
                        continue
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def cpu_oracle_rate(sample_members, threads, repeats=1):
    """members/s of the C oracle (port of the reference algorithm) on `threads` host cores, same workload."""
    from oracle import oracle as O
    W = WORKLOAD
    saveat = np.linspace(0.0, W["T"], W["nsave"])
    u0, p = make_inputs(sample_members)
    cfg = O.make_cfg(W["family"], W["sensealg"], W["stepper"], sample_members, saveat, 0.0, W["T"], dt=W["dt"],
                     cost=("affine",) + W["cost"])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.gradient(cfg, saveat, u0, p, want_saved=False, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample_members / best, best


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; Julia is not installed, so the
    reference itself cannot run) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = best_threads()
    sample = args.members or 8192
    times = []
    for i in range(args.warmup + args.steps):
        rate, dt = cpu_oracle_rate(sample, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    W = WORKLOAD
    line = {
        "impl": "reference", "metric": "ensemble adjoint trajectories/sec", "value": value, "unit": "trajectories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 Lorenz d=3 P=3 GaussAdjoint Tsit5 fixed dt=0.01 T=10 saveat=0.1 dgdu=u-2 (bounded sample)",
                   "members_per_step": sample, "S": int(round(W["T"] / W["dt"])), "K": W["nsave"]},
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} members of the C2 workload per step, OpenMP over members"},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import scimlsensitivity_jl_b200 as b

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    W = WORKLOAD
    N = args.members or W["members_per_gpu"]       # per GPU (weak scaling)
    S = int(round(W["T"] / W["dt"]))
    saveat = np.linspace(0.0, W["T"], W["nsave"])
    u0_h, p_h = make_inputs(N, offset=rank)
    cost = b.AffineCost(*W["cost"])
    dev = f"cuda:{local}"

    # ---------------- device-resident arm (`value`) ----------------
    eng = b.DeviceEnsemble(W["family"], W["sensealg"], W["stepper"], N, saveat, (0.0, W["T"]), W["dt"], on_device=True,
                           device=local, cost=cost, traj_offset=rank * N, block_threads=args.block)
    eng.use_current_torch_stream()
    u0_d = torch.tensor(u0_h, device=dev); p_d = torch.tensor(p_h, device=dev)
    du0_d = torch.empty((3, N), dtype=torch.float64, device=dev); dp_d = torch.empty(3, dtype=torch.float64, device=dev)

    def step_device():
        eng.handle.forward(u0_d, p_d, None, None, None)
        eng.handle.reverse(None, du0_d, dp_d)
        if world > 1:
            dist.all_reduce(dp_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = eng.handle.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    barrier()
    wall0 = time.time()
    ev[0].record()
    for i in range(args.steps):
        eng.handle.forward(u0_d, p_d, None, None, None)
        ev[3 * i + 1].record()
        eng.handle.reverse(None, du0_d, dp_d)
        ev[3 * i + 2].record()
        if world > 1:
            dist.all_reduce(dp_d)
        ev[3 * i + 3].record()
    barrier()
    wall1 = time.time()
    total_ms = ev[0].elapsed_time(ev[-1])
    fwd_ms = float(np.mean([ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps)]))
    rev_ms = float(np.mean([ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps)]))
    launches = eng.handle.launch_count - launches0 + (args.steps if world > 1 else 0)
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    t = torch.tensor([total_ms, fwd_ms, rev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms, rev_ms = (float(x) for x in t.cpu())
    ms_per_step = total_ms / args.steps
    value = N * world / (ms_per_step * 1e-3)
    dp_check = dp_d.cpu().numpy().tolist()

    # ---------------- end-to-end arm (`e2e`): public API, host buffers ----------------
    u0_pin = torch.tensor(u0_h).pin_memory(); p_pin = torch.tensor(p_h).pin_memory()
    prob = b.EnsembleProblem(b.ODEProblem(W["family"], u0_h[:, 0], (0.0, W["T"]), p_pin.numpy()), u0s=u0_pin.numpy())
    ealg = b.EnsembleB200(device=local, buffers_on_device=False, reuse_handle=True, presharded=True, pin_outputs=True)   # each rank owns its members
    alg = b.Tsit5(dt=W["dt"])

    def step_e2e():
        sol = b.solve(prob, alg, ealg, saveat=saveat, sensealg=b.B200Adjoint(b.GaussAdjoint(), block_threads=args.block),
                      save_on=False)
        return b.adjoint_sensitivities(sol, alg, t=saveat, dgdu_discrete=cost, sensealg=b.GaussAdjoint())

    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(max(1, min(args.warmup, 3))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        du0_e, dp_e = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.cpu()[0])
    e2e_value = N * world / e2e_s
    h2d = u0_h.nbytes + p_h.nbytes
    d2h = 3 * N * 8 + 3 * 8
    e2e_ok = bool(np.allclose(np.asarray(dp_e).ravel(), np.asarray(dp_check), rtol=1e-12))

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = ALG_BYTES_PER_MEMBER_STEP * N * S
        achieved = alg_bytes / (rev_ms * 1e-3) / 1e9
        compulsory = 8.0 * (S * 3 + 3 + W["nsave"] * 0) * N     # checkpoint read + du0 write (affine cost: no cotangent read)
        threads = best_threads()
        cpu_sample = 8192
        cpu_oracle_rate(256, threads)                      # warm the library / thread pool
        cpu_rate, cpu_s = cpu_oracle_rate(cpu_sample, threads)
        line = {
            "metric": "ensemble adjoint trajectories/sec", "value": value, "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 Lorenz d=3 P=3 N=65536/GPU GaussAdjoint Tsit5 fixed dt=0.01 T=10 saveat=0.1 dgdu=u-2 shared p",
                       "members_per_gpu": N, "S": S, "K": W["nsave"], "parallelism": f"ensemble-shard x{world}",
                       "l2": "per-step working set = 1.57 GB of checkpoints per GPU (>> 126 MB L2), no explicit flush",
                       "block_threads": args.block or "auto: ceil32(N / (n_SM * waves)) = 448, one block per SM"},
            "phases_ms": {"forward": fwd_ms, "reverse": rev_ms, "allreduce": max(0.0, ms_per_step - fwd_ms - rev_ms)},
            "roofline": {"bound": "hbm", "kernel": "tsit5_reverse_kernel<Lorenz,GAUSS>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic() if (N == W["members_per_gpu"]) else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_member_step": ALG_BYTES_PER_MEMBER_STEP,
                         "compulsory_bytes_per_launch": compulsory,
                         "note": "per-step accounting of SURVEY 8d (state counted as if it lived in HBM between steps); the time "
                                 "loop is in-kernel so real DRAM traffic is ~ the compulsory bytes; the kernel is fp64-FMA bound"},
            "cpu_baseline": {"value": cpu_rate, "unit": "trajectories/s", "cores": threads, "kind": "port",
                             "sample": f"{cpu_sample} members of the same C2 workload, one gradient, {cpu_s:.2f} s wall"},
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "api": "solve(EnsembleProblem, Tsit5, EnsembleB200) + adjoint_sensitivities(AffineCost)",
                    "matches_device_arm": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "dp": dp_check,
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_secondary(args):
    """--workload c2f32|c3|c4|c5: the other BASELINE configs (parity-test cases; reported for context, same JSON shape)."""
    import torch
    import scimlsensitivity_jl_b200 as b
    from oracle import oracle as O
    torch.cuda.set_device(0)
    w = args.workload
    rng = np.random.default_rng(WORKLOAD["seed"])
    threads = best_threads()
    if w == "c3":
        N = args.members or 16384
        T = 100.0
        saveat = np.logspace(-2, 2, 10); saveat[-1] = T
        u0 = np.repeat(np.array([[1.0], [0.0], [0.0]]), N, 1)
        p = np.array([0.04, 3e7, 1e4])[:, None] * np.exp(0.05 * rng.standard_normal((3, N)))
        kw = dict(abstol=1e-8, reltol=1e-8, quad_abstol=1e-10, quad_reltol=1e-10)
        eng = b.DeviceEnsemble("robertson", "quadrature", "rosenbrock23", N, saveat, (0.0, T), 0.0, shared_p=False, on_device=True,
                               cost=b.AffineCost(1.0, 0.0), max_steps=8192, **kw)
        ocfg = lambda n: O.make_cfg("robertson", "quadrature", "rosenbrock23", n, saveat, 0.0, T, cost=("affine", 1.0, 0.0), shared_p=False, **kw)
        name, dtype, sample = "C3 Robertson d=3 P=3 per-member k, QuadratureAdjoint(1e-10), Rosenbrock23 adaptive tol 1e-8, T=100, 10 log-spaced saves", "f64", 256
    elif w == "c2f32":
        # the fp32 throughput variant of C2: same ensemble, T = 1 (S = 100, K = 11), fp32 state and tables
        N = args.members or 65536
        T, dt = 1.0, 0.01
        saveat = np.linspace(0.0, T, 11)
        u0, p = make_inputs(N)
        dtype = "f32"
        eng = b.DeviceEnsemble("lorenz", "gauss", "tsit5_fixed", N, saveat, (0.0, T), dt, on_device=True, dtype=dtype, cost=b.AffineCost(1.0, -2.0))
        ocfg = lambda n: O.make_cfg("lorenz", "gauss", "tsit5_fixed", n, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -2.0))
        name, sample = "C2-fp32 Lorenz d=3 P=3 GaussAdjoint Tsit5 fixed dt=0.01 T=1 saveat=0.1 dgdu=u-2 shared p, fp32 state (fp64 oracle on the CPU side)", 16384
    elif w == "c4":
        N = args.members or 4096
        T, dt = 1.5, 0.05
        saveat = np.linspace(0.05, T, 30)
        u0 = rng.uniform(-2, 2, (2, N))
        H = 64
        p = np.concatenate([(rng.standard_normal((H, 2)) / np.sqrt(2)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                            (rng.standard_normal((H, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(H),
                            (rng.standard_normal((2, H)) / np.sqrt(H)).ravel(order="F"), 0.1 * rng.standard_normal(2)])
        dtype = args.dtype or "bf16_f32acc"       # bf16_f32acc (the named config: tensor-core VJP) | f32 | f64
        eng = b.DeviceEnsemble("mlp", "interpolating", "tsit5_fixed", N, saveat, (0.0, T), dt, on_device=True, dtype=dtype, cost=b.AffineCost(1.0, -0.5))
        ocfg = lambda n: O.make_cfg("mlp", "interpolating", "tsit5_fixed", n, saveat, 0.0, T, dt=dt, cost=("affine", 1.0, -0.5), mlp_hidden=H)
        name, sample = "C4 MLP 2->64->64->2 shared weights P=4482, InterpolatingAdjoint, Tsit5 fixed dt=0.05, T=1.5, 30 saves", 512
    else:
        N = args.members or 131072
        T, dt = 1.0, 0.01
        saveat = np.linspace(0.0, T, 101)
        u0 = np.ones((2, N)); p = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])
        eng = b.DeviceEnsemble("sde_lv", "backsolve", "em", N, saveat, (0.0, T), dt, on_device=True, cost=b.AffineCost(0.0, 1.0), seed=20260923)
        ocfg = lambda n: O.make_cfg("sde_lv", "backsolve", "em", n, saveat, 0.0, T, dt=dt, cost=("affine", 0.0, 1.0))
        name, dtype, sample = "C5 SDE-LV diag noise d=2 P=6, BacksolveAdjoint (Ito transformed drift), EM dt=0.01, T=1, saveat 0.01, Philox noise regenerated", "f64", 8192
    td = torch.float64 if dtype == "f64" else torch.float32
    u0_d = torch.tensor(u0, device="cuda", dtype=td); p_d = torch.tensor(p, device="cuda", dtype=td)
    du0_d = torch.empty(u0.shape, dtype=td, device="cuda"); dp_d = torch.empty(p.shape, dtype=td, device="cuda")

    def step():
        eng.handle.forward(u0_d, p_d, None, None, None)
        eng.handle.reverse(None, du0_d, dp_d)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    l0 = eng.handle.launch_count
    ev[0].record()
    for i in range(args.steps):
        eng.handle.forward(u0_d, p_d, None, None, None); ev[2 * i + 1].record()
        eng.handle.reverse(None, du0_d, dp_d); ev[2 * i + 2].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[-1]) / args.steps
    fwd = float(np.mean([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)]))
    rev = float(np.mean([ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)]))
    # CPU oracle on a bounded sample of the same workload
    cfgs = ocfg(sample)
    dW = None
    if w == "c5":
        dW = np.sqrt(dt) * rng.standard_normal((100, 2, sample))
    pc = p if p.ndim == 1 else p[:, :sample]
    O.gradient(ocfg(min(sample, 16)), saveat, u0[:, :min(sample, 16)], p if p.ndim == 1 else p[:, :min(sample, 16)], dW=None if dW is None else dW[:, :, :min(sample, 16)], want_saved=False, nthreads=threads)
    t0 = time.perf_counter()
    ref = O.gradient(cfgs, saveat, u0[:, :sample], pc, dW=dW, want_saved=False, nthreads=threads)
    cpu_s = time.perf_counter() - t0
    line = {"metric": "ensemble adjoint trajectories/sec", "value": N / (ms * 1e-3), "unit": "trajectories/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": {"workload": name, "members_per_gpu": N},
            "phases_ms": {"forward": fwd, "reverse": rev},
            "cpu_baseline": {"value": sample / cpu_s, "unit": "trajectories/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} members of the same workload, one gradient, {cpu_s:.2f} s wall"},
            "gpu_launches": int(eng.handle.launch_count - l0)}
    emit(line)
    eng.close()


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything libraries print (NCCL's version banner, torchrun notices) goes to stderr; stdout carries exactly the
    one JSON line, written through the saved descriptor by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=0, help="override members per GPU (default 65536) / reference sample")
    ap.add_argument("--block", type=int, default=0, help="CUDA block size override (multiple of 32, <= 512)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c2f32", "c3", "c4", "c5"], help="c2 = BASELINE headline (default)")
    ap.add_argument("--dtype", default="", help="c4 only: bf16_f32acc (default), f32 or f64")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "c2":
        run_secondary(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
