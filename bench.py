#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched physics step (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model humanoid] [--batch 65536] [--dtype f64]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the CPU arm: the oracle port timed on the host cores

Workload (config.workload): Humanoid (the reference's humanoid.xml: Newton, 1 iteration / 4 line-search iterations,
Euler, eulerdamp disabled -- the configuration the reference's README numbers are quoted on), batch 65 536 per GPU,
float64 like every published reference number.  A step = one `step(mx, d)` over the whole batch.

  value               device-resident stepping, `--outputs state` (default): the step returns the state fields only; the same
                      steps with `outputs="full"` (every field of Data, like the reference) are timed too -> other_output_mode
  e2e                 the same through mujoco_torch_b200.HostPipeline with the per-step input (ctrl) copied from pinned host memory
                      and the observation (qpos, qvel) read back, every step; e2e_host_state = round 1's protocol (whole state
                      travels both ways)
  roofline            the binding roof of the dominant kernel (FMA pipe), the whole-step HBM figure beside it, per-kernel split
                      measured on the SAME steps (replayed from the timed region's initial state)
  variants            BASELINE configs[3]: f32 and ImplicitFast variants of the workload (single-GPU runs)
  obs_gather          multi-rank runs: the optional NCCL all-gather of (qpos, qvel, reward, done) on a side stream
  cpu_baseline        the oracle port timed on this box's host cores; cpu_baseline_ref quotes the unmodified reference's eager
                      step timed in the build container (profiles/reference_eager_cpu.json)
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# reference README numbers, single H200, float64, torch.compile(vmap) tuned, XML-default options
L2_BYTES = 126 * 1000 * 1000   # B200 L2; batches whose per-step working set is smaller get a flush between timed steps
PUBLISHED = {("humanoid", "f64", 65536): 2333178.0, ("humanoid", "f64", 131072): 2490457.0, ("humanoid", "f64", 32768): 2054194.0,
             ("humanoid", "f64", 4096): 750288.0, ("humanoid", "f64", 1024): 202922.0, ("humanoid", "f64", 128): 26560.0,
             ("humanoid", "f64", 1): 232.0,
             ("halfcheetah", "f64", 4096): 745732.0, ("halfcheetah", "f64", 32768): 3584003.0, ("halfcheetah", "f64", 65536): 4122131.0,
             ("halfcheetah", "f64", 131072): 4479824.0, ("ant", "f64", 32768): 981050.0, ("ant", "f64", 65536): 1341747.0,
             ("ant", "f64", 131072): 1638131.0, ("walker2d", "f64", 32768): 537248.0, ("hopper", "f64", 131072): 4893648.0}
# minimal state I/O words per env-step: read qpos,qvel,ctrl,warmstart,time(+act); write qpos,qvel,qacc,warmstart,time(+act)


def algorithmic_bytes(mj, tsize):
    rd = mj.nq + mj.nv + mj.nu + mj.nv + 1 + mj.na
    wr = mj.nq + mj.nv + mj.nv + mj.nv + 1 + mj.na + mj.nsensordata
    return (rd + wr) * tsize


def algorithmic_flops(mj, mx):
    """Algorithmic flops per env-step from the model dims, SURVEY.md 8(d) formulae (FMA = 2 flops; the symmetric count of
    J^T D J, so that exploiting symmetry / sparsity is not rewarded twice).  Solver work is data dependent: the solver body
    is counted ONCE (exact for Humanoid, iterations = 1; a lower bound for the models that iterate to convergence) and the
    line search at its configured maximum of ls_iterations.  Returned per kernel (k_solve also factorises M and solves for
    qacc_smooth, and runs the integrator's re-factorisation) and in total."""
    nv, nbody, ncon, nefc = mj.nv, mj.nbody, mx.ncon, mx.nefc
    integ, newton = int(mj.opt.integrator), int(mj.opt.solver) == 2
    ls_it = int(mj.opt.ls_iterations)
    chol, tri = nv ** 3 / 3.0, 2.0 * nv * nv
    smooth = 700.0 * nbody + 6.0 * nv * nv                          # kinematics..rne, make_m (tree-sparse half)
    factor = chol + tri                                              # factor_m, qacc_smooth
    constraint = ncon * 2 * nv * 18 + 18 * nv * ncon + 4 * nefc * nv   # contact Jacobians, frame.Jdiff, rows + aref
    solve = 0.0
    if nefc:
        ctx = 4 * nefc * nv + 2 * nv * nv                            # J.x, J^T f, M.x
        grad = (nv * (nv + 1) * nefc + chol + tri) if newton else tri
        linesearch = 2 * nefc * nv + 2 * nv * nv + (2 + 3 * ls_it) * 10 * nefc
        solve = 3 * ctx + 2 * grad + linesearch + ctx               # warm/smooth contexts, 2 gradients, 1 line search, post-LS update
    refactor = chol + tri if integ == 3 or (integ == 0 and not (int(mj.opt.disableflags) & (1 << 15)) and float(np.abs(mj.dof_damping).sum()) > 0) else 0.0
    nfwd = 4 if integ == 1 else 1
    by_kernel = {"k_smooth": nfwd * smooth, "k_constraint": nfwd * constraint, "k_solve": nfwd * (factor + solve) + (0.0 if integ == 1 else refactor)}
    return {"by_kernel": by_kernel, "total": sum(by_kernel.values())}


def stock_options(args):
    """True when the run uses the XML's own integrator / solver / cone (what the committed ncu captures were taken on)."""
    return args.integrator is None and args.solver is None and args.cone is None


def load_model(name, args):
    from mujoco_torch_b200.mjmodel import MjModelLite

    mj = MjModelLite.load_npz(os.path.join(ROOT, "assets", "models", name + ".npz"))
    if args.integrator is not None:
        mj.opt.integrator = {"euler": 0, "rk4": 1, "implicitfast": 3}[args.integrator]
    if args.solver is not None:
        mj.opt.solver = {"cg": 1, "newton": 2}[args.solver]
    if args.cone is not None:
        mj.opt.cone = {"pyramidal": 0, "elliptic": 1}[args.cone]
    return mj


def initial_state(mj, name, B, rank=0):
    """Synthetic randomised initial states of this rank's env shard (mujoco_torch_b200/sharding.py)."""
    from mujoco_torch_b200 import sharding

    return sharding.initial_state(mj, name, rank * B, (rank + 1) * B, seed=42)


def measured_traffic(kernel, model, dtype, B, outputs="state"):
    """DRAM bytes per launch of `kernel` (ncu dram__bytes_read.sum + dram__bytes_write.sum of the committed capture,
    profiles/traffic.json), or None when no capture matches the workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(f"{kernel}:{model}:{dtype}:{B}:{outputs}")
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_rate(mj, name, B, nsteps, threads):
    """CPU arm: the oracle port (float64 batched torch on the host cores)."""
    from oracle import forward as of
    from oracle import model as om

    torch.set_num_threads(threads)
    m = om.OModel(mj, torch.float64)
    qpos, qvel, ctrl = initial_state(mj, name, B)
    d = of.make_data(m, B)
    d.qpos, d.qvel = torch.as_tensor(qpos), torch.as_tensor(qvel)
    d.ctrl = torch.as_tensor(ctrl[0])
    d = of.step(m, d)  # warm-up
    t0 = time.perf_counter()
    for i in range(nsteps):
        d.ctrl = torch.as_tensor(ctrl[i % 4])
        d = of.step(m, d)
    dt = time.perf_counter() - t0
    return B * nsteps / dt, dt / nsteps


def run_reference(args):
    """`--impl reference`: the reference's own implementation cannot be imported here (mujoco / tensordict are not
    installed), so this arm times the oracle port with all host threads on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mj = load_model(args.model, args)
    cores = os.cpu_count() or 1
    B = args.cpu_batch
    rates, per = [], []
    for _ in range(max(1, args.warmup // 3)):
        oracle_rate(mj, args.model, B, 1, cores)
    for _ in range(max(1, min(args.steps, 3))):
        r, ms = oracle_rate(mj, args.model, B, 2, cores)
        rates.append(r)
        per.append(ms)
    v = float(np.median(rates))
    sample = f"oracle port, B={B} envs x 2 steps per sample, {len(rates)} samples, {cores} torch threads, float64"
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.median(per)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, mj), "batch": B},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args, mj):
    integ = {0: "Euler", 1: "RK4", 3: "ImplicitFast"}[int(mj.opt.integrator)]
    solver = {1: "CG", 2: "Newton"}[int(mj.opt.solver)]
    cone = {0: "pyramidal", 1: "elliptic"}[int(mj.opt.cone)]
    return (f"{args.model} (nv={mj.nv}), {solver} it={int(mj.opt.iterations)}/ls={int(mj.opt.ls_iterations)}, {cone} cone, {integ}, "
            f"dt={float(mj.opt.timestep)}, batch {args.batch}/GPU")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="humanoid")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--integrator", default=None, choices=["euler", "rk4", "implicitfast"])
    ap.add_argument("--solver", default=None, choices=["cg", "newton"])
    ap.add_argument("--cone", default=None, choices=["pyramidal", "elliptic"])
    ap.add_argument("--outputs", default="state", choices=["state", "full"],
                    help="what a step materialises: 'state' = the state fields only (what a rollout feeds back in), 'full' = every "
                         "field of Data like the reference; the physics is identical (tests/test_gpu_api.py)")
    ap.add_argument("--cpu-batch", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather-obs", action="store_true",
                    help="also all-gather (qpos, qvel), reward, done over the job every step on a side stream (sharding.ObsGather); "
                         "on by default when the job has more than one rank")
    ap.add_argument("--no-gather-obs", action="store_true")
    ap.add_argument("--no-other-mode", action="store_true", help="skip the second timed loop with the other --outputs mode")
    ap.add_argument("--graph", action="store_true",
                    help="time the rollout through mujoco_torch_b200.GraphedStep (one CUDA-graph replay per step) instead of step() calls")
    ap.add_argument("--no-variants", action="store_true", help="skip the f32 / ImplicitFast variants of the Humanoid workload")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import mujoco_torch_b200 as mt
    from mujoco_torch_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if dist and os.environ.get("MJT_NO_CPU_BIND", "0") != "1":
        from mujoco_torch_b200 import sharding as _sh0

        numa = _sh0.bind_to_local_cpus(local)   # before any pinned allocation: first touch places the staging buffers
    if dist:
        import torch.distributed as td

        # NCCL may log its banner to stdout when NCCL_DEBUG is set in the environment: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            td.init_process_group("nccl", device_id=dev)
            td.barrier()   # creates the communicator now, while stdout is parked on stderr
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    tsize = 8 if args.dtype == "f64" else 4
    mj = load_model(args.model, args)
    B = args.batch
    mx = mt.device_put(mj, dtype=dtype).to(dev)
    qpos, qvel, ctrl = initial_state(mj, args.model, B, rank)  # rank r owns global envs [r*B, (r+1)*B)
    d0 = mt.make_data(mx).expand(B).clone()
    d0 = d0.replace(qpos=torch.as_tensor(qpos, device=dev).to(dtype), qvel=torch.as_tensor(qvel, device=dev).to(dtype))
    ctrls = [torch.as_tensor(c, device=dev).to(dtype).t().contiguous().t() for c in ctrl]  # env-minor resident ctrl

    def barrier():
        if dist:
            td.barrier()
        torch.cuda.synchronize()

    from mujoco_torch_b200 import sharding

    ws_bytes = ((mx.state_layout["pub_words"] + mx.state_layout["ws_words"]) if args.outputs == "state" else mx.data_words) * B * tsize
    l2_flush = torch.zeros(L2_BYTES * 2 // 4, dtype=torch.int32, device=dev) if ws_bytes < L2_BYTES else None

    def rollout(d, steps, outputs, gather=None, zeros_b=None, fixed=False):
        """`steps` steps from Data `d` with the resident ctrl sequence; returns (device ms between the two events, last Data)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.graph and outputs == args.outputs and gather is None and not fixed:
            graphed.reset(d)
            barrier()
            if l2_flush is not None:
                total = 0.0
                for i in range(steps):
                    l2_flush.add_(1)
                    ev0.record()
                    d = graphed(ctrls[i % 4])
                    ev1.record()
                    ev1.synchronize()
                    total += ev0.elapsed_time(ev1)
                barrier()
                return total, d
            ev0.record()
            for i in range(steps):
                d = graphed(ctrls[i % 4])
            ev1.record()
            barrier()
            return ev0.elapsed_time(ev1), d
        barrier()
        if l2_flush is not None:
            # working set smaller than L2: every step is timed on its own, with a write larger than L2 between the brackets
            total = 0.0
            for i in range(steps):
                l2_flush.add_(1)
                ev0.record()
                d = mt.step(mx, d.replace(ctrl=ctrls[i % 4]), fixed_iterations=fixed, outputs=outputs)
                if gather is not None:
                    gather.launch([d.qpos, d.qvel], zeros_b, zeros_b)
                    gather.wait()
                ev1.record()
                ev1.synchronize()
                total += ev0.elapsed_time(ev1)
            barrier()
            return total, d
        ev0.record()
        for i in range(steps):
            d = mt.step(mx, d.replace(ctrl=ctrls[i % 4]), fixed_iterations=fixed, outputs=outputs)
            if gather is not None:
                if i:
                    gather.wait()
                gather.launch([d.qpos, d.qvel], zeros_b, zeros_b)
        if gather is not None:
            gather.wait()
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1), d

    graphed = mt.GraphedStep(mx, d0, 1, outputs=args.outputs) if args.graph else None
    # ---- device-resident throughput (the headline): W untimed warm-up steps, then exactly K timed steps
    d = d0
    for i in range(args.warmup):
        d = mt.step(mx, d.replace(ctrl=ctrls[i % 4]), outputs=args.outputs)
    barrier()
    d_start = d   # the timed region's initial state, replayed below for the kernel split and the other output mode
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, d = rollout(d_start, args.steps, args.outputs)
    clocks = sampler.stop() if rank == 0 else None
    finite = bool(torch.isfinite(d.qpos).all())
    ms_max = sharding.max_over_ranks(ms, dev)
    value = world * B * args.steps / (ms_max * 1e-3)
    step_ms = ms_max / args.steps

    # ---- per-kernel device timing: the SAME steps replayed from the same state with every launch bracketed by CUDA events on
    # its stream (so that the split adds up to the step also for solvers whose iteration count depends on the state)
    ops.profile_begin()
    dd = d_start
    for i in range(args.steps):
        dd = mt.step(mx, dd.replace(ctrl=ctrls[i % 4]), outputs=args.outputs)
    prof = ops.profile_end()
    per_step = {k: v[0] / args.steps for k, v in prof.items()}
    top = max(per_step, key=per_step.get)
    # live-tile models: share of envs the last step handed to the full-tile pass (hidden per-env flag; full layout only)
    tile_overflow = None
    if args.outputs == "full" and getattr(mx, "live_caps", (0, 0))[0] > 0 and os.environ.get("MJT_LIVE_TILE", "1") != "0":
        try:
            # the full layout is one (data_words, B) env-minor buffer that starts with qpos; RK4 runs its forward passes in
            # stage buffers, so the flag of the last pass is only in the step's output for the one-pass integrators
            if int(mj.opt.integrator) == 1 or dd.qpos.stride() != (1, B):
                raise ValueError("no flag in this output")
            off = mx.field_off["scr_ovf"]
            flags = torch.as_strided(dd.qpos, (B,), (1,), dd.qpos.storage_offset() + off * B)
            tile_overflow = float((flags != 0).double().mean())
        except Exception:  # noqa: BLE001
            tile_overflow = None
    del dd

    # ---- the other output mode on the same steps (both numbers in one line: outputs=state is the rollout form, outputs=full
    # materialises every field of Data like the reference does)
    other = "full" if args.outputs == "state" else "state"
    other_mode = None
    if not args.no_other_mode:
        for i in range(3):
            mt.step(mx, d_start.replace(ctrl=ctrls[i % 4]), outputs=other)
        ms_o, _ = rollout(d_start, args.steps, other)
        ms_o = sharding.max_over_ranks(ms_o, dev)
        other_mode = {"outputs": other, "value": world * B * args.steps / (ms_o * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_o / args.steps}

    # ---- rooflines.  The dominant kernel (k_solve) is bound by FMA-pipe issue, not by HBM: the top-level roofline is the compute
    # one -- algorithmic flops of that kernel per launch over its measured launch time against the FMA peak measured in this run --
    # and the HBM figure (minimal state I/O of the whole step over the step time) sits beside it.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fl = algorithmic_flops(mj, mx)
    pipe_name = "fp64" if args.dtype == "f64" else "fp32"
    try:
        fma_peak = ops.fma_peak(args.dtype == "f64")
    except Exception:  # noqa: BLE001
        fma_peak = None
    kflops = fl["by_kernel"].get(top, fl["total"])
    ach = kflops * B / (per_step[top] * 1e-3) / 1e12
    traffic = {k: measured_traffic(k, args.model, args.dtype, B, args.outputs) for k in per_step}
    abytes = algorithmic_bytes(mj, tsize)
    roofline = {
        "bound": pipe_name, "kernel": top, "achieved": ach, "peak": fma_peak, "unit": "TFLOP/s", "frac": (ach / fma_peak) if fma_peak else None,
        "traffic": traffic.get(top), "algorithmic_flops_per_env_launch": kflops,
        "peak_source": "mjt_fma_peak: register-resident FMA chains timed in this run (MEASURED_PEAKS.json holds no FP64/FP32 vector peak)",
        "kernel_ms_per_step": per_step, "kernel_share": {k: v / sum(per_step.values()) for k, v in per_step.items()},
        "kernel_sum_ms": sum(per_step.values()),
        "whole_step": {"bound": pipe_name, "achieved": fl["total"] * B / (step_ms * 1e-3) / 1e12, "peak": fma_peak, "unit": "TFLOP/s",
                       "frac": (fl["total"] * B / (step_ms * 1e-3) / 1e12 / fma_peak) if fma_peak else None,
                       "algorithmic_flops_per_env_step": fl["total"],
                       "scope": "solver body counted once, line search at ls_iterations (SURVEY.md 8d formulae)"},
        "hbm": {"bound": "hbm", "achieved": abytes * B / (step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": abytes * B / (step_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_env_step": abytes,
                "traffic": (sum(v for v in traffic.values() if v) if all(traffic.get(k) for k in per_step if per_step[k] > 0.05 * step_ms) else None),
                "traffic_by_kernel": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                "scope": "whole step: minimal state I/O (SURVEY.md 8d) over ms_per_step; traffic = ncu dram bytes per step summed over the kernels (profiles/traffic.json)"},
    }

    # ---- end to end through the public API from pinned host buffers (mujoco_torch_b200.HostPipeline: H2D -> step -> D2H per chunk
    # on its own stream, one copy per direction).  Two protocols, both timed every run:
    #   e2e            : the simulation state stays on the GPU (where step() returns it); every step the step's input -- ctrl -- comes
    #                    from pinned host memory and its result -- the observation qpos, qvel -- is read back to the host.  This is
    #                    the traffic of a host-side policy driving the simulator, the call pattern of the reference's own users.
    #   e2e_host_state : round 1's protocol, kept for continuity: the WHOLE state (qpos, qvel, act, time, qacc_warmstart) + ctrl
    #                    travels to the device and the whole new state travels back, every step (2.4x the bytes).
    def measure_e2e(resident):
        pipe = mt.HostPipeline(mx, d, nchunk=int(os.environ["MJT_E2E_CHUNKS"]) if "MJT_E2E_CHUNKS" in os.environ else None,
                               outputs=args.outputs, resident=resident)
        for _ in range(3):
            pipe.step(synchronize=False)   # per-chunk streams keep their own H2D -> step -> D2H order across steps
        pipe.synchronize()
        barrier()
        n_e2e = max(5, min(args.steps, 30))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_e2e):
            pipe.step(synchronize=False)
        pipe.synchronize()
        e1.record()
        barrier()
        el = sharding.max_over_ranks(e0.elapsed_time(e1), dev) * 1e-3
        ok = all(bool(torch.isfinite(b["qpos"]).all()) for b in pipe.host)
        return {"value": world * B * n_e2e / el, "unit": "env-steps/s", "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                "steps": n_e2e, "chunks": len(pipe.slices), "finite": ok,
                "protocol": ("state resident on the device; per step H2D ctrl from pinned host memory, D2H qpos + qvel" if resident == "device"
                             else "state resident on the host; per step H2D qpos, qvel, act, time, qacc_warmstart, ctrl and D2H of the new state"),
                **({"cpu_binding": numa} if numa is not None else {})}

    e2e = e2e_host = None
    if not args.no_e2e:
        e2e = measure_e2e("device")
        e2e_host = measure_e2e("host")

    # ---- the one optional collective: all-gather of (qpos, qvel) + reward + done over the job on a side stream,
    # overlapped with the next physics step.  Measured whenever the job has more than one rank (or on request).
    obs_gather = None
    if args.gather_obs or (dist and not args.no_gather_obs):
        gather = sharding.ObsGather(mj.nq + mj.nv, B, dtype, dev)
        zeros_b = torch.zeros(B, dtype=dtype, device=dev)
        rollout(d_start, 3, args.outputs, gather, zeros_b)
        ms_g, _ = rollout(d_start, args.steps, args.outputs, gather, zeros_b)
        ms_g = sharding.max_over_ranks(ms_g, dev)
        obs_gather = {"value": world * B * args.steps / (ms_g * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_g / args.steps,
                      "added_ms_per_step": ms_g / args.steps - step_ms, "bytes_gathered_per_rank_per_step": gather.bytes_per_gather,
                      "message": "per env: qpos, qvel, reward, done", "backend": "nccl all_gather_into_tensor" if dist else "single rank (pack + copy only)"}

    # ---- BASELINE configs[3] in full on one line: the other precision / integrator variants of the Humanoid workload, same batch,
    # same protocol (3 warm-up + up to 20 timed steps each).  Single-GPU runs only.
    variants = None
    if not dist and not args.no_variants and args.model == "humanoid" and stock_options(args):
        variants = []
        for vd, vi in (("f64", "implicitfast"), ("f32", None), ("f32", "implicitfast")):
            if (vd, vi) == (args.dtype, None):
                continue
            vmj = load_model(args.model, argparse.Namespace(integrator=vi, solver=None, cone=None))
            vdt = torch.float64 if vd == "f64" else torch.float32
            vmx = mt.device_put(vmj, dtype=vdt).to(dev)
            vdata = mt.make_data(vmx).expand(B).clone().replace(qpos=torch.as_tensor(qpos, device=dev).to(vdt), qvel=torch.as_tensor(qvel, device=dev).to(vdt))
            vctrl = [c.to(vdt).t().contiguous().t() for c in ctrls]
            for i in range(3):
                vdata = mt.step(vmx, vdata.replace(ctrl=vctrl[i % 4]), outputs=args.outputs)
            nv_steps = max(5, min(args.steps, 20))
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(nv_steps):
                vdata = mt.step(vmx, vdata.replace(ctrl=vctrl[i % 4]), outputs=args.outputs)
            a1.record()
            torch.cuda.synchronize()
            vms = a0.elapsed_time(a1) / nv_steps
            bad = float((~torch.isfinite(vdata.qpos).all(1)).double().mean())
            variants.append({"dtype": vd, "integrator": vi or "euler (XML)", "outputs": args.outputs, "value": B / (vms * 1e-3), "unit": "env-steps/s",
                             "ms_per_step": vms, "steps": nv_steps, "nonfinite_env_frac": bad,
                             **({"note": "Humanoid's XML stops the solver after 1 iteration; ImplicitFast re-solves qacc from that unconverged constraint "
                                         "force, and a few % of the randomly actuated envs diverge within ~10 steps -- in the CPU oracle as well "
                                         "(DESIGN.md 4); _check_state resets a coordinate once it passes 1e10"} if vi == "implicitfast" else {})})
            del vdata, vmx

    if rank != 0:
        if dist:
            td.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and not dist:
        cores = os.cpu_count() or 1
        v, sec = oracle_rate(mj, args.model, args.cpu_batch, 3, cores)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"oracle port (float64 batched torch), B={args.cpu_batch} envs x 3 steps, {cores} threads"}
    # the UNMODIFIED reference's eager CPU step cannot run on the GPU box (its Python sources live under /root/reference, which does
    # not travel); it was timed in the build container through oracle/refshim (scripts/time_reference_eager.py) and is quoted as context
    cpu_ref = None
    try:
        cpu_ref = json.load(open(os.path.join(ROOT, "profiles", "reference_eager_cpu.json"))).get(args.model)
    except Exception:  # noqa: BLE001
        pass

    stock = stock_options(args)   # published numbers use the XML's own options
    pub = PUBLISHED.get((args.model, args.dtype, B)) if stock else None
    launches = ops.lib().mjt_step_launches(ops.model_handle(mx, args.outputs)) * args.steps
    import ctypes

    _e, _s = ctypes.c_int(), ctypes.c_int()
    ops.lib().mjt_solve_config(ops.model_handle(mx, args.outputs), ctypes.byref(_e), ctypes.byref(_s))
    words = (mx.state_layout["pub_words"] + mx.state_layout["ws_words"]) if args.outputs == "state" else mx.data_words
    line = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": (value / pub) if pub else None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args, mj), "batch_per_gpu": B,
                   "call": "GraphedStep: one CUDA-graph replay per step (same kernels as step())" if args.graph else "step() per step", "parallelism": f"dp{world} (envs sharded, no collective)",
                   "outputs": ("state: every step returns qpos, qvel, act, time, qacc, qacc_warmstart, act_dot, sensordata (bit-identical to the "
                               "full-Data step, tests/test_gpu_api.py); the full-Data number is in other_output_mode") if args.outputs == "state"
                   else "full: every field of Data materialised each step, like the reference",
                   "l2": (f"per-step working set {words * B * tsize / 1e6:.0f} MB > 126 MB L2 (inputs larger than L2)" if l2_flush is None else
                          f"per-step working set {words * B * tsize / 1e6:.0f} MB < 126 MB L2: L2 flushed (252 MB write) between timed steps, each "
                          "step bracketed by its own events"),
                   "nv": mj.nv, "ncon": mx.ncon, "nefc": mx.nefc, "solve_envs_per_cta": _e.value, "solve_smem_bytes": _s.value,
                   **({"solve_live_tile": {"rows": mx.live_caps[0], "j_words": mx.live_caps[1], "overflow_frac_last_step": tile_overflow}}
                      if tile_overflow is not None else {})},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "e2e_host_state": e2e_host, "other_output_mode": other_mode, "variants": variants, "obs_gather": obs_gather, "cpu_baseline_ref": cpu_ref, "finite": finite,
        "published_baseline": {"value": pub, "hardware": "H200 f64 torch.compile(vmap) tuned, reference README.md:178-247"} if pub else None,
    }
    print(json.dumps(line))
    if dist:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
