#!/usr/bin/env python
"""bench.py -- ts-PWS stacking hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one full stack (linear stack + ts-PWS) of the workload's traces.
Metric (BASELINE.json): stacked trace-samples/s = traces x samples / time, whole job over all ranks.

  value     inputs resident in HBM when the timed region starts (device plan API)
  e2e       the same stack through the reference-facing C-ABI call tspws_main() with HOST
            (pinned) buffers: H2D of the traces and D2H of the two stacks inside the timed region
  roofline  dominant kernel of the workload (colsum2_kernel for two-stage, the forward transform
            for single-stage), CUDA-event duration measured live in the timed region,
            algorithmic bytes = 4 B per trace-sample / measured HBM peak;
            frac_whole_step is the same arithmetic over the whole step
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, compiled from /root/reference/src by
            oracle/Makefile) on this box's host cores, bounded sample, rank 0, N=1 only
  parity    N=1: this run's stacks against the reference's on the same traces;
            N>1: every rank's stack against rank 0 (bit-wise), peer exchange against the NCCL all-reduce
            path, and the strong-scaling stack (8192 traces sharded) against the reference on rank 0
  extra_workloads (N=1): short legs of the other BASELINE configs (single-stage w0=6, Mexican hat two-stage,
            bootstrap / jackknife at 2000 x 32768, batch of station pairs 500 x 8192)

Scaling.  `value` is WEAK scaling: every rank holds `mtr` traces of ONE stack of N*mtr traces; the ranks exchange
the small group-sum block through peer memory (NVLink) inside the first tail kernel.  The same JSON line
carries `strong`: the BASELINE configs[2] reading -- the 8192-trace stack sharded 8192/N per rank -- for the
headline frame and for the Mexican-hat frame.  Independent units shard with no exchange at all:
--workload pairs_c5 (station pairs) and --workload resample_c4 (bootstrap realisations).
With --impl reference the reference's CPU implementation is timed instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # north-star headline: 8192 x 16384, Morlet default w0 (44 scales), two-stage Kmax=10 + unbiased, nu=2
    "two_stage_morlet": dict(kind="stack", mtr=8192, n=16384, params=dict(type=-1, Kmax=10, unbiased=1, wu=2.0, lrm=1),
                             desc="8192x16384 Morlet (44 scales) two-stage Kmax=10 unbiased nu=2 rm"),
    # BASELINE configs[2]: Mexican hat, two-stage, unbiased
    "two_stage_mexhat": dict(kind="stack", mtr=8192, n=16384, params=dict(type=-3, Kmax=10, unbiased=1, wu=2.0, lrm=1),
                             desc="8192x16384 Mexican hat (20 scales) two-stage Kmax=10 unbiased nu=2 rm"),
    # BASELINE configs[1]: Morlet w0=6, single stage (linear + ts-PWS)
    "single_stage_w6": dict(kind="stack", mtr=8192, n=16384, params=dict(type=-1, w0=6.0, w0set=4, wu=2.0, lrm=1),
                            desc="8192x16384 Morlet w0=6 (55 scales) single-stage nu=2 rm"),
    # BASELINE configs[3]: bootstrap (sub-sampling) realisations over 2000 x 32768; realisations shard over ranks
    "resample_c4": dict(kind="resample", mtr=2000, n=32768, M=1000, params=dict(type=-1, wu=2.0, lrm=1),
                        desc="2000x32768 default Morlet (48 scales), 1000 sub-sampling realisations p=0.5 "
                             "(single-stage and two-stage Kmax=10), realisations sharded over ranks"),
    # BASELINE configs[4]: 4096 independent station pairs; pairs shard over ranks
    "pairs_c5": dict(kind="pairs", mtr=500, n=8192, npairs=4096, params=dict(type=-1, wu=2.0, lrm=1),
                     desc="4096 station pairs x 500 traces x 8192 samples, default Morlet (40 scales) nu=2 rm, "
                          "pairs sharded over ranks"),
}
DEFAULT_WORKLOAD = "two_stage_morlet"
METRIC, UNIT = "stacked trace-samples/s", "trace-samples/s"

_JSON_FD = None


def protect_stdout():
    """Libraries (NCCL's version banner, torch elastic) write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the rest of the run and keep the real stdout for emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    fd = _JSON_FD if _JSON_FD is not None else 1
    while data:
        n = os.write(fd, data)
        data = data[n:]


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def load_pkg():
    spec = importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(ROOT, "ts-pws_b200", "__init__.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["ts_pws_b200"] = m
    spec.loader.exec_module(m)
    return m


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def nerr(a, b):
    """max|a-b| / max|b|"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b)) / d) if d > 0 else float(np.max(np.abs(a - b)))


def synthetic_signal(nmax):
    """Windowed exponential chirp in the style of the reference's src/FWTa/FWT_Test.m:38-47: f 0.005 -> 0.03
    cycles/sample over the central 80 %, Hann-tapered 40 % / 40 %.  Kept here so that the
    measured arm imports nothing from oracle/."""
    n0, n1 = int(0.1 * nmax), int(0.9 * nmax)
    L = n1 - n0
    t = np.arange(L, dtype=np.float64)
    f0, f1 = 0.005, 0.03
    k = (f1 / f0) ** (1.0 / L)
    phase = 2 * np.pi * f0 * (k ** t - 1) / np.log(k)
    win = np.ones(L)
    e = int(0.4 * L)
    h = 0.5 * (1 - np.cos(np.pi * np.arange(e) / e))
    win[:e] = h
    win[L - e:] = h[::-1]
    s = np.zeros(nmax)
    s[n0:n1] = np.sin(phase) * win
    return s


def make_traces(mtr, n, first=0, out=None):
    """x = 0.5*chirp + N(0,1), counter-based per first global trace index, fp32."""
    sig = (synthetic_signal(n) * 0.5).astype(np.float32)
    if out is None:
        out = np.empty((mtr, n), dtype=np.float32)
    rng = np.random.Generator(np.random.Philox(key=20171 + first))
    blk = 256
    for i in range(0, mtr, blk):
        j = min(mtr, i + blk)
        rng.standard_normal((j - i, n), dtype=np.float32, out=out[i:j])
        out[i:j] += sig
    return out


def bind_to_gpu_numa_node(local):
    """Run this rank (and first-touch its pinned buffers) on the CPUs local to its GPU: at 8 ranks the host-to-device
    copies otherwise cross the socket interconnect (R1: 184 GB/s aggregate = 42 % of 8 x the 1-GPU rate)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:00.0/local_cpulist"
        with open(path) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to {dom:04x}:{bus:02x}:00.0"
    except Exception as e:                              # noqa: BLE001
        return f"not bound ({type(e).__name__})"
    return "not bound"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons, pw = [], 0.0, set(), 0.0
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2])); pw = max(pw, float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": pw or None}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------
def reference_time(B, wl, sample_traces, steps=1, warmup=0, threads=None, x=None, params=None, selection=None):
    """Times the reference's own tspws_main on `sample_traces` traces of the workload (host).  Returns the
    outputs of the last call too (the parity checks use them)."""
    from oracle import ref_lib as R
    if not R.available():
        return None
    ref = B.declare_reference(R.load())
    ncores = os.cpu_count() or 1
    threads = threads or ncores
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    except Exception:
        pass
    if x is None:
        x = make_traces(sample_traces, wl["n"])
    p = B.default_params(**(params or wl["params"]))
    if p.Kmax and p.Kmax > sample_traces:
        p.Kmax = sample_traces
    ts, r = [], None
    for i in range(warmup + steps):
        xx = x.copy()
        t0 = time.perf_counter()
        r = B.call_tspws_main(ref, xx, p, dt=1.0, copy=False)
        t1 = time.perf_counter()
        assert r["rc"] == 0
        if i >= warmup:
            ts.append(t1 - t0)
    t = float(np.mean(ts))
    return dict(value=sample_traces * wl["n"] / t, seconds=t, threads=int(threads), sample_traces=sample_traces,
                ls=r["ls"], tsPWS=r["tsPWS"], out=r)


def cpu_sample_size(wl):
    # two-stage runs at ~1e8..7e8 samples/s on the CPU -> the full workload is ~0.2-2 s; single-stage at
    # ~1.5e6 samples/s -> 256 traces are ~3 s
    return wl["mtr"] if wl["params"].get("Kmax") else min(wl["mtr"], 256)


def run_reference_arm(args, wl, name):
    """The reference's own CPU implementation (oracle/_ref = /root/reference/src compiled unmodified) through its
    stock tspws_main on this box's host cores, --steps / --warmup honoured as given.  Each step is a bounded
    sample of the workload (all 8192 traces for the two-stage stacks: 0.2 s; 256 traces for single-stage: ~3 s)."""
    B = load_pkg()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if wl["kind"] != "stack":
        wl = dict(wl, params=dict(wl["params"]))
    sample = cpu_sample_size(wl)
    # keep the whole run within a few minutes: a step of the single-stage sample is seconds
    est = 0.25 if wl["params"].get("Kmax") else 3.5 * sample / 256 * wl["n"] / 16384
    steps = args.steps if est * (args.steps + args.warmup) <= 240 else max(1, int(240 / est) - args.warmup)
    r = reference_time(B, wl, sample, steps=steps, warmup=args.warmup)
    if r is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libtspws_ref.so missing (build with make -C oracle)"})
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "desc": wl["desc"],
                   "sample": f"{sample} of {wl['mtr']} traces x {wl['n']} samples per step (the reference's time is linear "
                             f"in the trace count, so trace-samples/s does not depend on the sample)"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                         "cpu": cpu_model(),
                         "sample": f"reference tspws_main (gcc -O3 -fopenmp, x86-64-v3) on {sample} traces x {wl['n']} samples"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm: one stack, traces sharded over the ranks
# ------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.B = load_pkg()
        self.lib = self.B.load_library(os.environ.get("TSPWS_LIB") or None)     # TSPWS_LIB: diagnostics builds (tools/pipe_timeline_mg.py)
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        self.numa = bind_to_gpu_numa_node(self.local)
        self.lib.tspws_b200_set_option(self.B.OPT_DEVICE, self.local)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line only
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.side = torch.cuda.Stream()      # a real stream: the legacy default stream cannot be graph-captured
        torch.cuda.set_stream(self.side)
        self.stream = self.side.cuda_stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def all_ok(self, ok):
        t = self.torch.tensor([1 if ok else 0], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t[0]) == 1

    def shutdown(self):
        if self.world > 1:
            t = threading.Timer(30.0, lambda: os._exit(0))    # never let teardown outlive the measurement
            t.daemon = True
            t.start()
            try:
                self.dist.barrier()
                self.dist.destroy_process_group()
            except Exception:                           # noqa: BLE001
                pass
            t.cancel()


class ShardedStack:
    """One stack of mtr_total traces, this rank holding [offset, offset + mtr_local): plan + exchange set-up."""

    def __init__(self, ctx, params, n, mtr_local, mtr_total, offset, want_peer=True):
        self.ctx, self.n = ctx, n
        B, lib, torch, dist = ctx.B, ctx.lib, ctx.torch, ctx.dist
        self.p = B.default_params(**params)
        lib.tspws_b200_derive_params(C.byref(self.p), n, 1.0)
        self.two_stage = bool(self.p.Kmax)
        self.mk = lambda: lib.tspws_b200_plan_create(C.byref(self.p), n, mtr_local, mtr_total, offset)
        self.plan = self.mk()
        if not self.plan:
            raise SystemExit("tspws_b200_plan_create failed")
        self.peer, self.exch, self.graph = False, None, None
        world = ctx.world
        if world > 1 and want_peer and not os.environ.get("TSPWS_BENCH_NCCL"):
            # peer exchange: every rank maps the others' exchange regions (CUDA IPC); plan_reduce pushes the rank's
            # block into every region over NVLink, plan_finish sums the blocks inside its first kernel.  The
            # decision is collective: if any rank cannot take part, all fall back to the all-reduce path.
            h = (C.c_ubyte * 64)()
            ok = lib.tspws_b200_plan_peer_export(self.plan, world, h) == 0
            mine = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            if ctx.all_ok(ok):
                buf = (C.c_ubyte * (64 * world))(*torch.cat(allh).cpu().tolist())
                ok = lib.tspws_b200_plan_peer_import(self.plan, ctx.rank, world, buf) == 0
            self.peer = ctx.all_ok(ok)
            if not self.peer:
                if ctx.rank == 0:
                    log("peer exchange set-up failed on some rank: using the NCCL all-reduce path")
                lib.tspws_b200_plan_peer_close(self.plan)       # unmap on every rank, then free (CUDA IPC rule)
                dist.barrier()
                lib.tspws_b200_plan_destroy(self.plan)          # a fresh plan without peer state
                self.plan = self.mk()
        if world > 1 and not self.peer:
            nbytes = lib.tspws_b200_plan_exchange_bytes(self.plan)
            is64 = lib.tspws_b200_plan_exchange_is_f64(self.plan)
            dt = torch.float64 if is64 else torch.float32
            self.exch = torch.zeros(nbytes // (8 if is64 else 4), device="cuda", dtype=dt)
            assert lib.tspws_b200_plan_bind_exchange(self.plan, self.exch.data_ptr()) == 0

    def step(self, d_x, d_ls, d_ts):
        ctx, lib = self.ctx, self.ctx.lib
        if ctx.world == 1 or self.peer:
            rc = lib.tspws_b200_plan_stack(self.plan, d_x.data_ptr(), d_ls.data_ptr(), d_ts.data_ptr(), ctx.stream)
        else:
            rc = lib.tspws_b200_plan_reduce(self.plan, d_x.data_ptr(), ctx.stream)
            ctx.dist.all_reduce(self.exch)
            lib.tspws_b200_plan_mark_exchanged(self.plan)
            rc |= lib.tspws_b200_plan_finish(self.plan, d_ls.data_ptr(), d_ts.data_ptr(), ctx.stream)
        if rc:
            raise SystemExit(f"stack failed rc={rc}")

    def capture(self, d_x, d_ls, d_ts):
        """N > 1 without the peer exchange: reduce kernels + NCCL all-reduce + tail kernels as one CUDA graph,
        like tspws_b200_plan_stack does by itself otherwise; all ranks replay the graph, or none does."""
        ctx, torch = self.ctx, self.ctx.torch
        if ctx.world == 1 or self.peer or os.environ.get("TSPWS_BENCH_NO_GRAPH"):
            return
        g = None
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=ctx.side, capture_error_mode="thread_local"):
                self.step(d_x, d_ls, d_ts)
        except Exception as e:                          # noqa: BLE001
            log(f"rank {ctx.rank}: graph capture of the multi-GPU step failed ({e}); direct launches")
            g = None
            torch.cuda.synchronize()
        self.graph = g if ctx.all_ok(g is not None) else None

    def fast_step(self, d_x, d_ls, d_ts):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.step(d_x, d_ls, d_ts)

    def timed(self, d_x, d_ls, d_ts, steps, fast=True):
        """K steps between two CUDA events on the launching stream, barrier + synchronize on both sides,
        max over ranks; returns ms per step."""
        ctx, torch = self.ctx, self.ctx.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn = self.fast_step if fast else self.step
        ctx.barrier()
        ev0.record()
        for _ in range(steps):
            fn(d_x, d_ls, d_ts)
        ev1.record()
        ctx.barrier()
        return ctx.max_over_ranks(ev0.elapsed_time(ev1))[0] / steps

    def close(self):
        ctx, lib = self.ctx, self.ctx.lib
        self.graph = None                               # the captured step holds NCCL work: drop it first
        ctx.torch.cuda.synchronize()
        if ctx.world > 1:
            try:
                ctx.dist.barrier()                      # every rank is done with the peers' blocks
                if self.peer:
                    lib.tspws_b200_plan_peer_close(self.plan)   # unmap first ...
                    ctx.dist.barrier()                          # ... free only when nobody maps the region any more
            except Exception:                           # noqa: BLE001
                pass
        lib.tspws_b200_plan_destroy(self.plan)
        self.plan = None


class PipedStacks:
    """A stream of independent stacks of one geometry through tspws_b200_pipe_* : the group sums of stack i+1 run
    while the tail of stack i finishes (depth plans as buffers).  At N > 1 every plan of the pipe gets the peer
    exchange of ShardedStack (one exchange region per plan)."""

    def __init__(self, ctx, params, n, mtr_local, mtr_total, offset, depth=4):
        self.ctx, self.n, self.depth = ctx, n, (depth or 4)          # depth 0 = AUTO (the library measures both ways)
        B, lib, torch, dist = ctx.B, ctx.lib, ctx.torch, ctx.dist
        p = B.default_params(**params)
        lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
        self.pipe = lib.tspws_b200_pipe_create(C.byref(p), n, mtr_local, mtr_total, offset, depth)
        self.ok = bool(self.pipe)
        world = ctx.world
        self.peer = False
        if world > 1:
            ok = self.ok
            for k in range(self.depth):
                plan = lib.tspws_b200_pipe_plan(self.pipe, k) if self.pipe else None
                h = (C.c_ubyte * 64)()
                ok = ok and lib.tspws_b200_plan_peer_export(plan, world, h) == 0
                mine = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
                allh = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(allh, mine)
                if ctx.all_ok(ok):
                    buf = (C.c_ubyte * (64 * world))(*torch.cat(allh).cpu().tolist())
                    ok = lib.tspws_b200_plan_peer_import(plan, ctx.rank, world, buf) == 0
                ok = ctx.all_ok(ok)
                if not ok:
                    break
            self.ok = self.peer = ok
        if self.ok:
            self.H = torch.cuda.ExternalStream(lib.tspws_b200_pipe_stream(self.pipe, 0))
            self.T = torch.cuda.ExternalStream(lib.tspws_b200_pipe_stream(self.pipe, 1))
            self.outs = [(torch.empty(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(self.depth)]
            self.i = 0

    def submit(self, d_x):
        k = self.i % self.depth
        rc = self.ctx.lib.tspws_b200_pipe_submit(self.pipe, d_x.data_ptr(), self.outs[k][0].data_ptr(), self.outs[k][1].data_ptr())
        if rc:
            raise SystemExit(f"pipe submit failed rc={rc}")
        self.i += 1
        return self.outs[k]

    def sync(self):
        if self.ctx.lib.tspws_b200_pipe_sync(self.pipe):
            raise SystemExit("pipe sync failed")

    def timed(self, xs, steps):
        """K stacks (inputs alternating over xs) between an event on the sums stream before the first submit and an
        event on the tail stream after the last, barrier + synchronize on both sides, max over ranks; ms per stack."""
        ctx, torch = self.ctx, self.ctx.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync()
        ctx.barrier()
        ev0.record(self.H)
        for j in range(steps):
            self.submit(xs[j % len(xs)])
        ev1.record(self.T)
        self.sync()
        ctx.barrier()
        return ctx.max_over_ranks(ev0.elapsed_time(ev1))[0] / steps

    def close(self):
        ctx, lib = self.ctx, self.ctx.lib
        if not self.pipe:
            return
        ctx.torch.cuda.synchronize()
        if ctx.world > 1:
            try:
                ctx.dist.barrier()
                for k in range(self.depth):
                    lib.tspws_b200_plan_peer_close(lib.tspws_b200_pipe_plan(self.pipe, k))
                ctx.dist.barrier()
            except Exception:                           # noqa: BLE001
                pass
        lib.tspws_b200_pipe_destroy(self.pipe)
        self.pipe = None


def peer_err_check(lib, ps):
    return ps.peer and any(lib.tspws_b200_plan_peer_error(lib.tspws_b200_pipe_plan(ps.pipe, k)) for k in range(ps.depth))


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of colsum2_kernel at 8192 x 16384 (432 CTAs), from the committed
# ncu --set full capture profiles/r02_colsum2_ncu_raw.csv -- not measured inside a bench run (ncu replays kernels)
COLSUM2_TRAFFIC = (537.754e6, "NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed "
                   "ncu --set full capture of this kernel at this shape and grid (profiles/r02_colsum2_ncu_raw.csv: 536.76 MB + 0.99 MB)")


def whole_step_frac(mtr, n, ms_step, hbm_peak):
    return (4.0 * mtr * n / (ms_step * 1e-3) / 1e9) / hbm_peak


def strong_leg(ctx, params, n, mtr_total, d_x, steps, check_reference):
    """BASELINE configs[2] as written: ONE stack of mtr_total traces, traces sharded mtr_total / N per rank (the
    first rows of every rank's resident traces).  Returns the line's `strong` entry."""
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    m = mtr_total // world
    d_ls = torch.empty(n, device="cuda", dtype=torch.float32)
    d_ts = torch.empty(n, device="cuda", dtype=torch.float32)
    st = ShardedStack(ctx, params, n, m, m * world, rank * m)
    xs = d_x[:m]
    for _ in range(3):
        st.step(xs, d_ls, d_ts)
    ctx.barrier()
    st.capture(xs, d_ls, d_ts)
    for _ in range(2):
        st.fast_step(xs, d_ls, d_ts)
    ms = st.timed(xs, d_ls, d_ts, steps)
    out = {"traces_total": m * world, "traces_per_gpu": m, "ms_per_step": ms,
           "value": float(m * world) * n / (ms * 1e-3), "unit": UNIT,
           "exchange": "peer memory over NVLink" if st.peer else ("NCCL all-reduce" if world > 1 else "none")}
    if st.peer and ctx.lib.tspws_b200_plan_peer_error(st.plan):
        out["error"] = "a peer wait timed out"
    if check_reference:
        # all shards to rank 0, the reference's tspws_main on the assembled stack (host), then compare
        if world > 1:
            parts = [torch.empty_like(xs) for _ in range(world)] if rank == 0 else None
            dist.gather(xs.contiguous(), parts, dst=0)
        else:
            parts = [xs]
        if rank == 0:
            try:
                full = torch.cat(parts).cpu().numpy()
                r = reference_time(ctx.B, dict(n=n, params=params), full.shape[0], x=full, params=params)
                if r:
                    out["parity_vs_reference"] = {"tsPWS": nerr(d_ts.cpu().numpy(), r["tsPWS"]), "ls": nerr(d_ls.cpu().numpy(), r["ls"]),
                                                  "tolerance": 1e-4, "reference_seconds": r["seconds"]}
            except Exception as e:                      # noqa: BLE001
                out["parity_vs_reference"] = f"not run ({type(e).__name__}: {e})"
    st.close()
    return out


def run_stack(args, wl, name):
    ctx = Ctx(args)
    torch, dist, B, lib = ctx.torch, ctx.dist, ctx.B, ctx.lib
    world, rank, local = ctx.world, ctx.rank, ctx.local
    mtr, n = wl["mtr"], wl["n"]

    host = torch.empty((mtr, n), dtype=torch.float32).pin_memory()
    make_traces(mtr, n, first=rank * mtr, out=host.numpy())
    d_x = host.cuda(non_blocking=True)
    d_ls = torch.empty(n, device="cuda", dtype=torch.float32)
    d_ts = torch.empty(n, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()

    st = ShardedStack(ctx, wl["params"], n, mtr, mtr * world, rank * mtr)
    p, two_stage, peer = st.p, st.two_stage, st.peer
    steps = args.steps

    for _ in range(max(3, args.warmup)):
        st.step(d_x, d_ls, d_ts)
    ctx.barrier()
    st.capture(d_x, d_ls, d_ts)
    for _ in range(2):
        st.fast_step(d_x, d_ls, d_ts)
    ctx.barrier()

    # ---- timed region: K steps, device-resident inputs (512 MiB per rank > 126 MB L2) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    # (1) with an event after every kernel: the per-kernel durations the roofline uses
    lib.tspws_b200_profile(1)
    lib.tspws_b200_profile_reset()
    n0 = lib.tspws_b200_kernel_launches()
    ms_prof = st.timed(d_x, d_ls, d_ts, steps, fast=False)
    launches = lib.tspws_b200_kernel_launches() - n0
    prof = B.read_profile(lib)
    lib.tspws_b200_profile(0)
    # (2) the same K steps without those events (they cost a few us per step): this is `value`
    ms_step = st.timed(d_x, d_ls, d_ts, steps)
    # (3) sustained: the same step back to back for ~1.5 s (the K-step region above lasts a few ms)
    sus_steps = int(min(20000, max(steps, 1.5e3 / max(ms_step, 1e-3))))
    ms_sustained = st.timed(d_x, d_ls, d_ts, sus_steps)

    # the dominant two-stage kernel alone, back to back (no per-kernel events in between): on an
    # unsharded plan plan_reduce() launches exactly the group-sum kernel
    k_alone_ms = None
    if two_stage and world == 1:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            lib.tspws_b200_plan_reduce(st.plan, d_x.data_ptr(), ctx.stream)
        ctx.barrier()
        ev0.record()
        for _ in range(steps):
            lib.tspws_b200_plan_reduce(st.plan, d_x.data_ptr(), ctx.stream)
        ev1.record()
        ctx.barrier()
        k_alone_ms = ev0.elapsed_time(ev1) / steps
        st.step(d_x, d_ls, d_ts)                        # leave complete outputs behind
        torch.cuda.synchronize()

    # (4) two-stage: a stream of independent stacks, software-pipelined (tspws_b200_pipe_*): the HBM-bound group sums of
    # stack i+1 overlap the latency-bound tail of stack i.  This is the throughput `value` reports; the single-stack
    # figures above stay in config.  Two input buffers alternate (each 512 MiB > L2).
    pipelined = None
    ms_single = ms_step
    if two_stage and not args.no_pipe:
        ps = PipedStacks(ctx, wl["params"], n, mtr, mtr * world, rank * mtr, depth=args.pipe_depth)
        if ps.ok:
            d_x2 = torch.roll(d_x, 1, 0)
            xs = [d_x, d_x2]
            for j in range(2 * ps.depth):
                o_ls, o_ts = ps.submit(xs[j % 2])
            ps.sync()
            ps.i = 0
            o_ls, o_ts = ps.submit(d_x)                    # parity of the pipelined path against the single-stack step
            ps.sync()
            pipe_diff = nerr(o_ts.cpu().numpy(), d_ts.cpu().numpy())
            ps.i = 0
            n1 = lib.tspws_b200_kernel_launches()
            ms_pipe = ps.timed(xs, steps)
            pipe_launches = lib.tspws_b200_kernel_launches() - n1
            ms_pipe_sus = ps.timed(xs, sus_steps)
            pipelined = {"stacks_in_flight": ps.depth, "api": "tspws_b200_pipe_submit / tspws_b200_pipe_sync",
                         "ms_per_step": ms_pipe, "ms_per_step_sustained": ms_pipe_sus, "sustained_steps": sus_steps,
                         "tsPWS_vs_single_stack_step": ctx.max_over_ranks(pipe_diff)[0],
                         "input_buffers": 2, "exchange": ("peer memory over NVLink" if ps.peer else "none")}
            if peer_err_check(lib, ps):
                raise SystemExit(f"rank {rank}: a peer wait timed out in the pipelined leg -- the measurement is invalid")
            # `value` is the better of the two ways to run K stacks; the pipeline only counts when it also wins over the
            # long run (beside the peer exchange at N > 1 it is erratic: a lucky K-step figure must not become the headline)
            if ms_pipe < ms_step and ms_pipe_sus < ms_sustained:
                ms_step, launches = ms_pipe, pipe_launches
            else:
                pipelined["note"] = "not faster than one stack at a time here (K steps and sustained): value / ms_per_step are the single-stack figures"
            del d_x2
        ps.close()

    total_samples = float(mtr) * n * world
    value = total_samples / (ms_step * 1e-3)

    # ---- parity ----
    parity = {}
    if world > 1:
        mine = d_ts.clone()
        allt = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        parity["all_ranks_bit_identical"] = bool(all(torch.equal(allt[0], t) for t in allt))
        if peer:
            # the same stack through the NCCL all-reduce path (a second plan without peer state)
            ref_ls, ref_ts = torch.empty_like(d_ls), torch.empty_like(d_ts)
            st2 = ShardedStack(ctx, wl["params"], n, mtr, mtr * world, rank * mtr, want_peer=False)
            st2.step(d_x, ref_ls, ref_ts)
            torch.cuda.synchronize()
            parity["peer_vs_nccl_allreduce_tsPWS"] = ctx.max_over_ranks(nerr(d_ts.cpu().numpy(), ref_ts.cpu().numpy()))[0]
            st2.close()

    # ---- e2e: host (pinned) buffers, H2D + D2H inside the timed region ----
    e2e = None
    pe = B.default_params(**wl["params"])
    ls_h = torch.empty(n, dtype=torch.float32).pin_memory()
    ts_h = torch.empty(n, dtype=torch.float32).pin_memory()
    e_steps = max(1, min(steps, 10))
    if world > 1:
        # N > 1: pinned host shard -> H2D -> stack (peer exchange or all-reduce) -> D2H of the stacks,
        # all inside the timed region, through the public device-plan API
        def e2e_step():
            d_x.copy_(host, non_blocking=True)
            st.step(d_x, d_ls, d_ts)
            ls_h.copy_(d_ls, non_blocking=True)
            ts_h.copy_(d_ts, non_blocking=True)
            torch.cuda.synchronize()
        e2e_step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        ctx.barrier()
        t1 = time.perf_counter()
        te = ctx.max_over_ranks((t1 - t0) / e_steps)[0]
        e2e = {"value": total_samples / te, "unit": UNIT,
               "h2d_bytes_per_step": int(mtr) * n * 4 * world, "d2h_bytes_per_step": 2 * n * 4 * world,
               "ms_per_step": te * 1e3, "steps": e_steps, "h2d_gbs_aggregate": int(mtr) * n * 4 * world / te / 1e9,
               "host_binding": ctx.numa,
               "api": "tspws_b200_plan_stack (peer exchange) with pinned host shards" if peer else
                      "tspws_b200_plan_reduce / all-reduce / tspws_b200_plan_finish with pinned host shards"}
    else:
        out, dat = B.TsPWSOut(), B.Data()
        dat.sigall = C.cast(host.data_ptr(), C.POINTER(C.c_float))
        dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = n, mtr, 1.0, -(n - 1) / 2.0
        out.ls = C.cast(ls_h.data_ptr(), C.POINTER(C.c_float))
        out.tsPWS = C.cast(ts_h.data_ptr(), C.POINTER(C.c_float))
        out.N, out.mtr = n, mtr
        for _ in range(2):
            pp = B.TsPWS.from_buffer_copy(pe)
            assert lib.tspws_main(C.byref(pp), C.byref(out), C.byref(dat)) == 0
        t0 = time.perf_counter()
        for _ in range(e_steps):
            pp = B.TsPWS.from_buffer_copy(pe)
            assert lib.tspws_main(C.byref(pp), C.byref(out), C.byref(dat)) == 0
        t1 = time.perf_counter()
        te = (t1 - t0) / e_steps
        # the bare copy of the same bytes, for scale
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(3):
            d_x.copy_(host, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        h2d_ms = ev0.elapsed_time(ev1) / 3
        e2e = {"value": total_samples / te, "unit": UNIT,
               "h2d_bytes_per_step": int(mtr) * n * 4, "d2h_bytes_per_step": 2 * n * 4,
               "ms_per_step": te * 1e3, "steps": e_steps, "h2d_only_ms": h2d_ms, "host_binding": ctx.numa,
               "api": "tspws_main(t_tsPWS*, t_tsPWS_out*, t_data*) with pinned host buffers"}
        # parity spot check of the e2e outputs against the device-plan outputs of the same traces
        e2e["max_abs_diff_vs_device_path"] = float(np.max(np.abs(ts_h.numpy() - d_ts.cpu().numpy())))

    clocks = sampler.finish() if sampler else None      # sampled over the timed region and the e2e loop
    if peer and lib.tspws_b200_plan_peer_error(st.plan):
        raise SystemExit(f"rank {rank}: a peer wait timed out -- the measurement is invalid")

    # ---- strong scaling of ONE 8192-trace stack (BASELINE configs[2]); N = 1 gives the reference point ----
    strong = None
    if two_stage and not args.no_strong and (world > 1):
        strong = {"headline_frame": strong_leg(ctx, wl["params"], n, mtr, d_x, max(steps, 20), True)}
        if name != "two_stage_mexhat":
            strong["mexhat_frame"] = strong_leg(ctx, WORKLOADS["two_stage_mexhat"]["params"], n, mtr, d_x, max(steps, 20), True)
        strong["note"] = ("one stack of 8192 traces, 8192/N per rank; the N = 1 line's ms_per_step is the reference point; "
                          "the tail (transforms of the Kmax group sums + inverse) runs replicated on every rank")

    st_close = st.close
    if rank != 0:
        st_close()
        ctx.shutdown()
        return

    # ---- roofline of the dominant kernel ----
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_src = None, None
    alg_bytes = 4.0 * mtr * n                            # every fp32 sample read once
    if two_stage:
        kname = "colsum2_kernel" if "colsum2_kernel" in prof else "colsum_kernel"
        kcnt, kms = prof.get(kname, (0, 0.0))
        if kname == "colsum2_kernel" and (mtr, n) == (8192, 16384):
            traffic, traffic_src = COLSUM2_TRAFFIC
    else:
        parts = {k: v for k, v in prof.items() if k.startswith("fwd2_acc_kernel") or k == "fwd_kernel" or k.startswith("fwd_dense_mma")
                 or k in ("trace_image_kernel", "tile_sum_kernel", "dense_finish_kernel")}
        kname = "forward transform kernels"
        kcnt = max([v[0] for v in parts.values()] or [0])
        kms = sum(v[1] for v in parts.values())
    k_avg_ms = kms / max(kcnt, 1)
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "kernel_ms": k_avg_ms, "kernel_timing": "CUDA events recorded around the kernel inside the timed steps",
                "kernel_ms_back_to_back": k_alone_ms,
                "traffic": traffic, "traffic_source": traffic_src,
                "frac_whole_step": whole_step_frac(mtr, n, ms_step, hbm_peak),
                "frac_whole_step_sustained": whole_step_frac(mtr, n, min(pipelined["ms_per_step_sustained"], ms_sustained) if pipelined else ms_sustained, hbm_peak),
                "frac_single_stack_step": whole_step_frac(mtr, n, ms_single, hbm_peak),
                "kernels_ms_per_step": {k: v[1] / steps for k, v in prof.items() if v[0]}}
    roofline["whole_step_frac"] = roofline["frac_whole_step"]
    if not two_stage:
        # single-stage is compute bound: also report the FP32 rate of the forward transform
        fr = lib.tspws_b200_frame_create(p.type, p.J, p.V, n, p.s0, p.b0, p.w0, p.uni)
        f = fr.contents
        taps = sum(int(f.L[s]) * int(f.Ns[s]) for s in range(f.S))
        lib.tspws_b200_frame_destroy(fr)
        roofline["fp32_tflops_direct_form"] = 4.0 * taps * mtr / (k_avg_ms * 1e-3) / 1e12 if k_avg_ms > 0 else 0.0

    cpu = None
    if world == 1 and not args.no_cpu:
        sample = cpu_sample_size(wl)
        r = reference_time(B, wl, sample, steps=1, warmup=0, x=host.numpy()[:sample])
        if r:
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference", "cpu": cpu_model(),
                   "sample": f"reference tspws_main (oracle/_ref) on {r['sample_traces']} of {mtr} traces x {n} samples, "
                             f"{r['seconds']:.2f} s on {os.cpu_count()} host cores"}
            if sample == mtr:                           # same traces, same parameters: the parity figure of this run
                parity["vs_reference"] = {"tsPWS": nerr(d_ts.cpu().numpy(), r["tsPWS"]), "ls": nerr(d_ls.cpu().numpy(), r["ls"]),
                                          "tolerance": 1e-4}

    extra = None
    if world == 1 and not args.no_extra and name == DEFAULT_WORKLOAD:
        st_close()
        st_close = lambda: None                          # noqa: E731
        del d_x
        torch.cuda.empty_cache()
        extra = extra_workloads(ctx, args, hbm_peak)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "desc": wl["desc"], "traces_per_gpu": mtr, "samples": n,
                   "scales": int(p.J * p.V), "parallelism": f"traces sharded x{world}, {'peer-memory sum' if peer else 'NCCL all-reduce'} of the "
                   f"{'group sums' if two_stage else 'phasor/linear sums'}" if world > 1 else "single GPU",
                   "l2_note": "per-step input 512 MiB per GPU > 126 MB L2 (no flush needed)",
                   "ms_per_step_with_kernel_events": ms_prof,
                   "step": ("K independent stacks submitted to a 2-stream software pipeline (pipelined); ms_per_step = elapsed / K"
                            if pipelined and ms_step == pipelined["ms_per_step"] else "K stacks one after the other"),
                   "pipelined": pipelined,
                   "single_stack": {"ms_per_step": ms_single, "ms_per_step_sustained": ms_sustained, "sustained_steps": sus_steps,
                                    "note": "one stack at a time through tspws_b200_plan_stack (latency of a stack)"},
                   "ms_per_step_sustained": min(pipelined["ms_per_step_sustained"], ms_sustained) if pipelined else ms_sustained, "sustained_steps": sus_steps,
                   "step_replayed_as_cuda_graph": bool(world == 1 or peer or st.graph is not None),
                   "exchange": ("none" if world == 1 else "peer memory over NVLink, fused into the stack's own kernels" if peer
                                else "NCCL all-reduce")},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "parity": parity or None, "strong": strong, "extra_workloads": extra,
    }
    emit(line)
    st_close()
    ctx.shutdown()


# ------------------------------------------------------------------------------------------
# the other BASELINE configs as short legs of the default run (N = 1)
# ------------------------------------------------------------------------------------------
def short_stack_leg(ctx, wlname, steps, hbm_peak, cpu_traces):
    torch, B, lib = ctx.torch, ctx.B, ctx.lib
    wl = WORKLOADS[wlname]
    mtr, n = wl["mtr"], wl["n"]
    host = make_traces(mtr, n)
    d_x = torch.from_numpy(host).cuda()
    d_ls = torch.empty(n, device="cuda", dtype=torch.float32)
    d_ts = torch.empty(n, device="cuda", dtype=torch.float32)
    st = ShardedStack(ctx, wl["params"], n, mtr, mtr, 0)
    for _ in range(3):
        st.step(d_x, d_ls, d_ts)
    lib.tspws_b200_profile(1)
    lib.tspws_b200_profile_reset()
    st.timed(d_x, d_ls, d_ts, steps, fast=False)
    prof = B.read_profile(lib)
    lib.tspws_b200_profile(0)
    ms = st.timed(d_x, d_ls, d_ts, steps)
    out = {"desc": wl["desc"], "value": float(mtr) * n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "frac_whole_step": whole_step_frac(mtr, n, ms, hbm_peak),
           "kernels_ms_per_step": {k: v[1] / steps for k, v in prof.items() if v[0]}}
    if st.two_stage:
        # the stream-of-stacks API in AUTO mode: it measures pipelined against one-at-a-time on its first 16 stacks
        ps = PipedStacks(ctx, wl["params"], n, mtr, mtr, 0, depth=0)
        if ps.ok:
            for _ in range(24):
                ps.submit(d_x)
            ps.sync()
            ps.submit(d_x)
            cal = (C.c_float * 2)()
            mode = lib.tspws_b200_pipe_mode(ps.pipe, cal)
            msp = ps.timed([d_x], steps)
            out["stream_of_stacks_auto"] = {"api": "tspws_b200_pipe_create(depth = 0)", "chosen": {0: "undecided", 1: "pipelined", 2: "one at a time"}.get(mode, mode),
                                            "calibration_ms_per_stack": {"pipelined": cal[0], "one_at_a_time": cal[1]},
                                            "ms_per_step": msp, "value": float(mtr) * n / (msp * 1e-3),
                                            "frac_whole_step": whole_step_frac(mtr, n, msp, hbm_peak)}
        ps.close()
    if not st.two_stage:
        fr = lib.tspws_b200_frame_create(st.p.type, st.p.J, st.p.V, n, st.p.s0, st.p.b0, st.p.w0, st.p.uni)
        f = fr.contents
        taps = sum(int(f.L[s]) * int(f.Ns[s]) for s in range(f.S))
        lib.tspws_b200_frame_destroy(fr)
        out["tflops_direct_form"] = 4.0 * taps * mtr / (ms * 1e-3) / 1e12
    if cpu_traces:
        r = reference_time(B, wl, cpu_traces, x=host[:cpu_traces])
        if r:
            out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference", "cpu": cpu_model(),
                                   "sample": f"{cpu_traces} of {mtr} traces, {r['seconds']:.2f} s"}
            if cpu_traces == mtr:
                out["parity_vs_reference"] = {"tsPWS": nerr(d_ts.cpu().numpy(), r["tsPWS"]), "ls": nerr(d_ls.cpu().numpy(), r["ls"])}
    st.close()
    return out


def resample_leg(ctx, M, two_stage, cpu_M=0, rows=None):
    """configs[3]: M sub-sampling realisations (p = 0.5) of 2000 x 32768 through tspws_main, host buffers."""
    B, lib = ctx.B, ctx.lib
    wl = WORKLOADS["resample_c4"]
    mtr, n = wl["mtr"], wl["n"]
    x = make_traces(mtr, n)
    K = mtr // 2
    rng = np.random.default_rng(11)
    sel = np.zeros((wl["M"], mtr), np.int8)
    for m in range(wl["M"]):
        sel[m, rng.choice(mtr, K, replace=False)] = 1
    if rows is None:
        rows = (0, M)
    mysel = np.ascontiguousarray(sel[rows[0]:rows[1]])
    Mloc = mysel.shape[0]
    kw = dict(wl["params"], subsmpl_N=Mloc, subsmpl_p=0.5)
    if two_stage:
        kw.update(Kmax=10, unbiased=1)
    lib.tspws_b200_set_selection(mysel.ctypes.data_as(C.c_char_p), Mloc, mtr)
    try:
        # warm-up: the same call once (plan, device buffers, pinned result staging), then the better of two timed calls --
        # the call is host-heavy (pageable H2D of 262 MB, scatter of 2 x M rows) and its time moves with the host's state
        B.call_tspws_main(lib, x, B.default_params(**kw))
        r, dt = None, None
        for _ in range(2):
            ri = B.call_tspws_main(lib, x, B.default_params(**kw))
            if dt is None or ri["seconds"] < dt:
                r, dt = ri, ri["seconds"]               # the tspws_main call itself (outputs allocated and touched before)
    finally:
        lib.tspws_b200_set_selection(None, 0, 0)
    assert r["rc"] == 0, r["rc"]
    out = {"realisations": Mloc, "seconds": dt, "realisations_per_s": Mloc / dt,
           "value": float(Mloc) * K * n / dt, "unit": UNIT,
           "note": "value counts the K = 1000 traces every realisation stacks; call includes the H2D of the 262 MB of traces "
                   "and the D2H of 2 x M x 32768 floats"}
    if cpu_M:
        from oracle import ref_lib as R
        if R.available():
            ref = B.declare_reference(R.load())
            kwc = dict(kw, subsmpl_N=cpu_M)
            # the reference draws its own random plans (rand()); its time does not depend on which traces are drawn
            t0 = time.perf_counter()
            rr = B.call_tspws_main(ref, x, B.default_params(**kwc))
            dtc = time.perf_counter() - t0
            if rr["rc"] == 0:
                out["cpu_baseline"] = {"value": float(cpu_M) * K * n / dtc, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                       "cpu": cpu_model(), "sample": f"M = {cpu_M} realisations (+ the main stack), {dtc:.1f} s"}
    return out, r


def pairs_leg(ctx, npairs, reps=2):
    """configs[4]: independent station pairs through tspws_b200_main_batch from pinned host memory."""
    torch, B, lib = ctx.torch, ctx.B, ctx.lib
    wl = WORKLOADS["pairs_c5"]
    mtr, n = wl["mtr"], wl["n"]
    host = torch.empty((npairs, mtr, n), dtype=torch.float32).pin_memory()
    make_traces(npairs * mtr, n, first=ctx.rank * npairs * mtr, out=host.numpy().reshape(npairs * mtr, n))
    ls = torch.zeros(npairs, n).pin_memory()
    ts = torch.zeros(npairs, n).pin_memory()
    xs = [host[i].numpy() for i in range(npairs)]
    lsn = [ls[i].numpy() for i in range(npairs)]
    tsn = [ts[i].numpy() for i in range(npairs)]
    p = B.default_params(**wl["params"])
    r = B.call_main_batch(lib, xs, p, out_ls=lsn, out_ts=tsn)      # warm-up
    assert r["rc"] == 0
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = B.call_main_batch(lib, xs, p, out_ls=lsn, out_ts=tsn)
        assert r["rc"] == 0
    ctx.barrier()
    dt = ctx.max_over_ranks((time.perf_counter() - t0) / reps)[0]
    return {"pairs": npairs * ctx.world, "seconds": dt, "pairs_per_s": npairs * ctx.world / dt,
            "value": float(npairs) * ctx.world * mtr * n / dt, "unit": UNIT,
            "h2d_gbs_aggregate": float(npairs) * ctx.world * mtr * n * 4 / dt / 1e9,
            "api": "tspws_b200_main_batch, pinned host buffers (H2D of every pair inside the timed region)"}, host, ts


def extra_workloads(ctx, args, hbm_peak):
    out = {}
    t_start = time.perf_counter()
    try:
        out["two_stage_mexhat"] = short_stack_leg(ctx, "two_stage_mexhat", 20, hbm_peak, 8192)
        out["single_stage_w6"] = short_stack_leg(ctx, "single_stage_w6", 3, hbm_peak, 0 if args.no_cpu else 128)
        ctx.torch.cuda.empty_cache()
        o1, _ = resample_leg(ctx, 1000, two_stage=False, cpu_M=0)      # the reference needs ~1 min for the single-stage main stack alone
        o2, _ = resample_leg(ctx, 1000, two_stage=True, cpu_M=0 if args.no_cpu else 16)
        out["resample_c4"] = {"desc": WORKLOADS["resample_c4"]["desc"], "single_stage": o1, "two_stage": o2}
        o3, host, ts = pairs_leg(ctx, 64)
        o3["desc"] = WORKLOADS["pairs_c5"]["desc"] + " (64 of the 4096 pairs per batch call)"
        if not args.no_cpu:
            wl = WORKLOADS["pairs_c5"]
            r = reference_time(ctx.B, wl, wl["mtr"], x=host[0].numpy())
            if r:
                o3["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference", "cpu": cpu_model(),
                                      "sample": f"1 pair (500 x 8192), {r['seconds']:.2f} s"}
                o3["parity_vs_reference"] = {"tsPWS": nerr(ts[0].numpy(), r["tsPWS"])}
        out["pairs_c5"] = o3
    except Exception as e:                              # noqa: BLE001
        out["error"] = f"{type(e).__name__}: {e}"
    out["seconds_total"] = time.perf_counter() - t_start
    return out


# ------------------------------------------------------------------------------------------
# independent units sharded over the ranks, no exchange: station pairs, realisations
# ------------------------------------------------------------------------------------------
def run_pairs(args, wl, name):
    ctx = Ctx(args)
    world, rank = ctx.world, ctx.rank
    per_rank_total = wl["npairs"] // world
    batch = 64                                           # pairs per batch call (1 GiB of pinned host memory)
    reps = max(1, min(args.steps, per_rank_total // batch))
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    if sampler:
        sampler.start()
    o, _, _ = pairs_leg(ctx, batch, reps=reps)
    clocks = sampler.finish() if sampler else None
    if rank == 0:
        emit({"metric": METRIC, "value": o["value"], "unit": UNIT, "n_gpus": world, "steps": reps, "warmup": 1,
              "ms_per_step": o["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic",
              "config": {"workload": name, "desc": wl["desc"], "pairs_per_step": batch * world,
                         "note": f"a step = one tspws_b200_main_batch call of {batch} pairs per rank from pinned host memory; "
                                 f"{wl['npairs']} pairs = {wl['npairs'] // (batch * world)} such steps per rank; no communication",
                         "host_binding": ctx.numa},
              "pairs_per_s": o["pairs_per_s"], "e2e": {"value": o["value"], "unit": UNIT,
                                                        "h2d_bytes_per_step": batch * world * wl["mtr"] * wl["n"] * 4,
                                                        "d2h_bytes_per_step": batch * world * 2 * wl["n"] * 4,
                                                        "h2d_gbs_aggregate": o["h2d_gbs_aggregate"]},
              "clocks": clocks, "gpu_launches": int(ctx.lib.tspws_b200_kernel_launches())})
    ctx.shutdown()


def run_resample(args, wl, name):
    ctx = Ctx(args)
    world, rank = ctx.world, ctx.rank
    M = wl["M"]
    rows = (M * rank // world, M * (rank + 1) // world)
    res = {}
    for two_stage in (False, True):
        ctx.barrier()
        o, _ = resample_leg(ctx, M, two_stage, rows=rows)
        dt = ctx.max_over_ranks(o["seconds"])[0]
        res["two_stage" if two_stage else "single_stage"] = {
            "seconds": dt, "realisations_per_s": M / dt, "value": float(M) * (wl["mtr"] // 2) * wl["n"] / dt, "unit": UNIT}
    if rank == 0:
        v = res["single_stage"]
        emit({"metric": METRIC, "value": v["value"], "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
              "ms_per_step": v["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic",
              "config": {"workload": name, "desc": wl["desc"], "realisations_per_gpu": rows[1] - rows[0],
                         "note": "a step = all M = 1000 realisations (M / N per rank, traces replicated, no communication) through "
                                 "tspws_main with host buffers; value = single-stage; two_stage beside it"},
              "single_stage": res["single_stage"], "two_stage": res["two_stage"],
              "e2e": {"value": v["value"], "unit": UNIT, "h2d_bytes_per_step": world * wl["mtr"] * wl["n"] * 4,
                      "d2h_bytes_per_step": 2 * M * wl["n"] * 4},
              "gpu_launches": int(ctx.lib.tspws_b200_kernel_launches())})
    ctx.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the short legs of the other BASELINE configs (N = 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling legs (N > 1)")
    ap.add_argument("--no-pipe", action="store_true", help="two-stage: time one stack at a time instead of the stack pipeline")
    ap.add_argument("--pipe-depth", type=int, default=4, help="stacks in flight in the pipelined leg")
    args = ap.parse_args()
    protect_stdout()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl if wl["kind"] == "stack" else WORKLOADS[DEFAULT_WORKLOAD], args.workload)
    elif wl["kind"] == "pairs":
        run_pairs(args, wl, args.workload)
    elif wl["kind"] == "resample":
        run_resample(args, wl, args.workload)
    else:
        run_stack(args, wl, args.workload)


if __name__ == "__main__":
    main()
