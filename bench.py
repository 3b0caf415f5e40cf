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
            algorithmic bytes = 4 B per trace-sample / measured HBM peak
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, compiled from /root/reference/src by
            oracle/Makefile) on this box's host cores, bounded sample, rank 0, N=1 only

Weak scaling: every rank holds `mtr` traces of ONE stack of N*mtr traces; ranks exchange the small
group-sum (two-stage) or phasor-sum (single-stage) block with an NCCL all-reduce between the two
phases.  With --impl reference the reference's CPU implementation is timed instead (rank 0 only).
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
    "two_stage_morlet": dict(mtr=8192, n=16384, params=dict(type=-1, Kmax=10, unbiased=1, wu=2.0, lrm=1),
                             desc="8192x16384 Morlet (44 scales) two-stage Kmax=10 unbiased nu=2 rm"),
    # BASELINE configs[2]: Mexican hat, two-stage, unbiased
    "two_stage_mexhat": dict(mtr=8192, n=16384, params=dict(type=-3, Kmax=10, unbiased=1, wu=2.0, lrm=1),
                             desc="8192x16384 Mexican hat (20 scales) two-stage Kmax=10 unbiased nu=2 rm"),
    # BASELINE configs[1]: Morlet w0=6, single stage (linear + ts-PWS)
    "single_stage_w6": dict(mtr=8192, n=16384, params=dict(type=-1, w0=6.0, w0set=4, wu=2.0, lrm=1),
                            desc="8192x16384 Morlet w0=6 (55 scales) single-stage nu=2 rm"),
}
DEFAULT_WORKLOAD = "two_stage_morlet"


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


def make_traces(mtr, n, first=0):
    """x = 0.5*chirp + N(0,1), counter-based per global trace index, fp32."""
    sig = (synthetic_signal(n) * 0.5).astype(np.float32)
    out = np.empty((mtr, n), dtype=np.float32)
    rng = np.random.Generator(np.random.Philox(key=20171 + first))
    blk = 256
    for i in range(0, mtr, blk):
        j = min(mtr, i + blk)
        out[i:j] = rng.standard_normal((j - i, n), dtype=np.float32)
        out[i:j] += sig
    return out


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
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------
def reference_time(B, wl, sample_traces, steps=1, warmup=0, threads=None):
    """Times the reference's own tspws_main on `sample_traces` traces of the workload (host)."""
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
    x = make_traces(sample_traces, wl["n"])
    p = B.default_params(**wl["params"])
    if p.Kmax and p.Kmax > sample_traces:
        p.Kmax = sample_traces
    ts = []
    for i in range(warmup + steps):
        xx = x.copy()
        t0 = time.perf_counter()
        r = B.call_tspws_main(ref, xx, p, dt=1.0, copy=False)
        t1 = time.perf_counter()
        assert r["rc"] == 0
        if i >= warmup:
            ts.append(t1 - t0)
    t = float(np.mean(ts))
    return dict(value=sample_traces * wl["n"] / t, seconds=t, threads=int(threads), sample_traces=sample_traces)


def cpu_sample_size(wl):
    # two-stage runs at ~1e8 samples/s on the CPU -> the full workload is ~1-2 s; single-stage at
    # ~1.5e6 samples/s -> 1024 traces are ~10-15 s
    return wl["mtr"] if wl["params"].get("Kmax") else min(wl["mtr"], 1024)


def run_reference_arm(args, wl, name):
    B = load_pkg()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = cpu_sample_size(wl)
    ref_steps = max(1, min(args.steps, 5))       # each step is seconds of CPU work: keep the run within minutes
    r = reference_time(B, wl, sample, steps=ref_steps, warmup=min(args.warmup, 1))
    if r is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libtspws_ref.so missing (build with make -C oracle)"})
        return
    line = {
        "impl": "reference", "metric": "stacked trace-samples/s", "value": r["value"], "unit": "trace-samples/s",
        "n_gpus": args.gpus, "steps": ref_steps, "warmup": min(args.warmup, 1),
        "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "desc": wl["desc"], "sample": f"{sample} of {wl['mtr']} traces x {wl['n']} samples"},
        "cpu_baseline": {"value": r["value"], "unit": "trace-samples/s", "cores": r["threads"], "kind": "reference",
                         "sample": f"reference tspws_main (gcc -O3 -fopenmp, x86-64-v3) on {sample} traces x {wl['n']} samples"},
        "e2e": {"value": r["value"], "unit": "trace-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, wl, name):
    import torch
    import torch.distributed as dist

    B = load_pkg()
    lib = B.load_library()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    lib.tspws_b200_set_option(B.OPT_DEVICE, local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    mtr, n = wl["mtr"], wl["n"]
    p = B.default_params(**wl["params"])
    lib.tspws_b200_derive_params(C.byref(p), n, 1.0)
    two_stage = bool(p.Kmax)

    host = torch.from_numpy(make_traces(mtr, n, first=rank * mtr)).pin_memory()
    d_x = host.cuda(non_blocking=True)
    d_ls = torch.empty(n, device="cuda", dtype=torch.float32)
    d_ts = torch.empty(n, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()

    plan = lib.tspws_b200_plan_create(C.byref(p), n, mtr, mtr * world, rank * mtr)
    if not plan:
        raise SystemExit("tspws_b200_plan_create failed")
    side = torch.cuda.Stream()                      # a real stream: the legacy default stream cannot be graph-captured
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    exch = None
    peer = False
    if world > 1 and two_stage and not os.environ.get("TSPWS_BENCH_NCCL"):
        # peer exchange: every rank maps the others' exchange blocks (CUDA IPC) and the sum over ranks is
        # fused into the first tail kernel (NVLink loads) -- no collective call in the step
        h = (C.c_ubyte * 64)()
        ok = lib.tspws_b200_plan_peer_export(plan, world, h) == 0
        mine = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        buf = (C.c_ubyte * (64 * world))(*torch.cat(allh).cpu().tolist())
        ok = ok and lib.tspws_b200_plan_peer_import(plan, rank, world, buf) == 0
        okt = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        peer = int(okt[0]) == 1
        if not peer:
            if rank == 0:
                print("[bench] peer exchange set-up failed on some rank: using the NCCL all-reduce path", file=sys.stderr)
            lib.tspws_b200_plan_peer_close(plan)        # unmap on every rank, then free (CUDA IPC rule)
            dist.barrier()
            lib.tspws_b200_plan_destroy(plan)           # a fresh plan without peer state
            plan = lib.tspws_b200_plan_create(C.byref(p), n, mtr, mtr * world, rank * mtr)
            if not plan:
                raise SystemExit("tspws_b200_plan_create failed")
    if world > 1 and not peer:
        nbytes = lib.tspws_b200_plan_exchange_bytes(plan)
        is64 = lib.tspws_b200_plan_exchange_is_f64(plan)
        dt = torch.float64 if is64 else torch.float32
        exch = torch.zeros(nbytes // (8 if is64 else 4), device="cuda", dtype=dt)
        assert lib.tspws_b200_plan_bind_exchange(plan, exch.data_ptr()) == 0

    def step():
        if world == 1 or peer:
            rc = lib.tspws_b200_plan_stack(plan, d_x.data_ptr(), d_ls.data_ptr(), d_ts.data_ptr(), stream)
        else:
            rc = lib.tspws_b200_plan_reduce(plan, d_x.data_ptr(), stream)
            dist.all_reduce(exch)
            lib.tspws_b200_plan_mark_exchanged(plan)
            rc |= lib.tspws_b200_plan_finish(plan, d_ls.data_ptr(), d_ts.data_ptr(), stream)
        if rc:
            raise SystemExit(f"stack failed rc={rc}")

    # N > 1: the whole step (reduce kernels, NCCL all-reduce, tail kernels) replayed as one CUDA graph,
    # like tspws_b200_plan_stack does on one GPU; falls back to direct launches if capture fails
    step_graph = None
    def capture_step():
        nonlocal step_graph
        if world == 1 or peer or os.environ.get("TSPWS_BENCH_NO_GRAPH"):
            return                                      # plan_stack replays its own graph
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                step()
            step_graph = g
        except Exception as e:                      # noqa: BLE001
            print(f"[bench] rank {rank}: graph capture of the multi-GPU step failed ({e}); direct launches", file=sys.stderr)
            step_graph = None
            torch.cuda.synchronize()

    def fast_step():
        if step_graph is not None:
            step_graph.replay()
        else:
            step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    capture_step()
    ok = torch.tensor([1 if (world == 1 or peer or step_graph is not None) else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks replay the graph, or none does
        if int(ok[0]) == 0:
            step_graph = None
    for _ in range(2):
        fast_step()
    barrier()

    # ---- timed region: K steps, device-resident inputs (512 MiB per rank > 126 MB L2) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    lib.tspws_b200_profile(1)
    lib.tspws_b200_profile_reset()
    n0 = lib.tspws_b200_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.tspws_b200_kernel_launches() - n0
    prof = B.read_profile(lib)
    lib.tspws_b200_profile(0)

    # same loop without the per-kernel events (they cost a few us per step): this is `value`
    barrier()
    ev0.record()
    for _ in range(args.steps):
        fast_step()
    ev1.record()
    barrier()
    ms_clean = ev0.elapsed_time(ev1)

    # the dominant two-stage kernel alone, back to back (no per-kernel events in between): on an
    # unsharded plan plan_reduce() launches exactly the group-sum kernel
    k_alone_ms = None
    if two_stage and world == 1:
        for _ in range(3):
            lib.tspws_b200_plan_reduce(plan, d_x.data_ptr(), stream)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            lib.tspws_b200_plan_reduce(plan, d_x.data_ptr(), stream)
        ev1.record()
        barrier()
        k_alone_ms = ev0.elapsed_time(ev1) / args.steps

    t = torch.tensor([ms_clean, ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_clean, ms = float(t[0]), float(t[1])
    ms_step = ms_clean / args.steps
    total_samples = float(mtr) * n * world
    value = total_samples / (ms_step * 1e-3)

    # ---- e2e: tspws_main with pinned host buffers, H2D + D2H inside the timed region ----
    e2e = None
    pe = B.default_params(**wl["params"])
    if world > 1:
        # N > 1: pinned host shard -> H2D -> reduce -> NCCL all-reduce -> finish -> D2H of the stacks,
        # all inside the timed region, through the public device-plan API
        ls_h = torch.empty(n, dtype=torch.float32).pin_memory()
        ts_h = torch.empty(n, dtype=torch.float32).pin_memory()
        e_steps = max(1, min(args.steps, 10))

        def e2e_step():
            d_x.copy_(host, non_blocking=True)
            step()
            ls_h.copy_(d_ls, non_blocking=True)
            ts_h.copy_(d_ts, non_blocking=True)
            torch.cuda.synchronize()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        t1 = time.perf_counter()
        te = torch.tensor([(t1 - t0) / e_steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_samples / float(te[0]), "unit": "trace-samples/s",
               "h2d_bytes_per_step": int(mtr) * n * 4 * world, "d2h_bytes_per_step": 2 * n * 4 * world,
               "ms_per_step": float(te[0]) * 1e3, "steps": e_steps,
               "api": "tspws_b200_plan_stack (peer exchange) with pinned host shards" if peer else
                      "tspws_b200_plan_reduce / all-reduce / tspws_b200_plan_finish with pinned host shards"}
    else:
        out, dat = B.TsPWSOut(), B.Data()
        ls_h = torch.empty(n, dtype=torch.float32).pin_memory()
        ts_h = torch.empty(n, dtype=torch.float32).pin_memory()
        dat.sigall = C.cast(host.data_ptr(), C.POINTER(C.c_float))
        dat.hdr.max, dat.hdr.mtr, dat.hdr.dt, dat.hdr.beg = n, mtr, 1.0, -(n - 1) / 2.0
        out.ls = C.cast(ls_h.data_ptr(), C.POINTER(C.c_float))
        out.tsPWS = C.cast(ts_h.data_ptr(), C.POINTER(C.c_float))
        out.N, out.mtr = n, mtr
        e_steps = max(1, min(args.steps, 10))
        for _ in range(2):
            pp = B.TsPWS.from_buffer_copy(pe)
            assert lib.tspws_main(C.byref(pp), C.byref(out), C.byref(dat)) == 0
        t0 = time.perf_counter()
        for _ in range(e_steps):
            pp = B.TsPWS.from_buffer_copy(pe)
            assert lib.tspws_main(C.byref(pp), C.byref(out), C.byref(dat)) == 0
        t1 = time.perf_counter()
        e2e = {"value": total_samples / ((t1 - t0) / e_steps), "unit": "trace-samples/s",
               "h2d_bytes_per_step": int(mtr) * n * 4, "d2h_bytes_per_step": 2 * n * 4,
               "ms_per_step": (t1 - t0) / e_steps * 1e3, "steps": e_steps,
               "api": "tspws_main(t_tsPWS*, t_tsPWS_out*, t_data*) with pinned host buffers"}
        # parity spot check of the e2e outputs against the device-plan outputs of the same traces
        e2e["max_abs_diff_vs_device_path"] = float(np.max(np.abs(ts_h.numpy() - d_ts.cpu().numpy())))

    clocks = sampler.finish() if sampler else None      # sampled over the timed region and the e2e loop
    if peer and lib.tspws_b200_plan_peer_error(plan):
        raise SystemExit(f"rank {rank}: a peer wait timed out -- the measurement is invalid")

    def teardown():
        nonlocal step_graph
        step_graph = None                               # the captured step holds NCCL work: drop it first
        torch.cuda.synchronize()
        t = None
        if world > 1:
            # never let teardown outlive the measurement: the JSON line is already out
            t = threading.Timer(30.0, lambda: os._exit(0))
            t.daemon = True
            t.start()
            try:
                dist.barrier()                          # every rank is done with the peers' blocks
                if peer:
                    lib.tspws_b200_plan_peer_close(plan)    # unmap first ...
                    dist.barrier()                          # ... free only when nobody maps the region any more
            except Exception:                           # noqa: BLE001
                pass
        lib.tspws_b200_plan_destroy(plan)
        if world > 1:
            try:
                dist.destroy_process_group()
            except Exception:                           # noqa: BLE001
                pass
            t.cancel()

    if rank != 0:
        teardown()
        return

    # ---- roofline of the dominant kernel ----
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    if two_stage:
        kname = "colsum2_kernel" if "colsum2_kernel" in prof else "colsum_kernel"
        kcnt, kms = prof.get(kname, (0, 0.0))
        alg_bytes = 4.0 * mtr * n                        # every fp32 sample read once
        if kname == "colsum2_kernel" and (mtr, n) == (8192, 16384):
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
            # kernel at this shape: profiles/r01_colsum2_ncu_raw.csv (536.76 MB + 1.71 MB)
            traffic = 538.475e6
    else:
        parts = {k: v for k, v in prof.items() if k.startswith("fwd2_acc_kernel") or k == "fwd_kernel"}
        kname = "fwd2_acc_kernel<0..5>"
        kcnt = max([v[0] for v in parts.values()] or [0])
        kms = sum(v[1] for v in parts.values())
        alg_bytes = 4.0 * mtr * n
    k_in_step_ms = kms / max(kcnt, 1)
    # average launch duration = the in-step interval between the CUDA events recorded around the kernel
    # inside the timed region (it contains one event record, so it is a slight over-estimate); the
    # back-to-back figure (kernel launched alone, K launches between two events) is reported beside it
    k_avg_ms = k_in_step_ms
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "kernel_ms": k_avg_ms, "kernel_timing": "CUDA events recorded around the kernel inside the timed steps",
                "kernel_ms_back_to_back": k_alone_ms,
                "traffic": traffic,
                "whole_step_frac": (alg_bytes / (ms_step * 1e-3) / 1e9) / hbm_peak,
                "kernels_ms_per_step": {k: v[1] / args.steps for k, v in prof.items() if v[0]}}
    if not two_stage:
        # single-stage is compute bound: also report the FP32 rate of the forward transform
        fr = lib.tspws_b200_frame_create(p.type, p.J, p.V, n, p.s0, p.b0, p.w0, p.uni)
        f = fr.contents
        taps = sum(int(f.L[s]) * int(f.Ns[s]) for s in range(f.S))
        lib.tspws_b200_frame_destroy(fr)
        roofline["fp32_tflops_direct_form"] = 4.0 * taps * mtr / (k_avg_ms * 1e-3) / 1e12 if k_avg_ms > 0 else 0.0

    cpu = None
    if world == 1 and not args.no_cpu:
        r = reference_time(B, wl, cpu_sample_size(wl), steps=1, warmup=0)
        if r:
            cpu = {"value": r["value"], "unit": "trace-samples/s", "cores": r["threads"], "kind": "reference",
                   "sample": f"reference tspws_main (oracle/_ref) on {r['sample_traces']} of {mtr} traces x {n} samples, "
                             f"{r['seconds']:.2f} s on {os.cpu_count()} host cores"}

    line = {
        "metric": "stacked trace-samples/s", "value": value, "unit": "trace-samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "desc": wl["desc"], "traces_per_gpu": mtr, "samples": n,
                   "scales": int(p.J * p.V), "parallelism": f"traces sharded x{world}, {'peer-memory sum' if peer else 'NCCL all-reduce'} of the "
                   f"{'group sums' if two_stage else 'phasor/linear sums'}" if world > 1 else "single GPU",
                   "l2_note": "per-step input 512 MiB per GPU > 126 MB L2 (no flush needed)",
                   "ms_per_step_with_kernel_events": ms / args.steps,
                   "step_replayed_as_cuda_graph": bool(world == 1 or peer or step_graph is not None),
                   "exchange": ("none" if world == 1 else "peer memory over NVLink, fused into group_post2_kernel" if peer
                                else "NCCL all-reduce")},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    emit(line)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    protect_stdout()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl, args.workload)
    else:
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()
