#!/usr/bin/env python
"""bench.py — ALS ratings/sec per iteration (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mal|netflix|ml-1m|ml-100k]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's CPU path (multi-threaded BLAS port)

A step is one full ALS iteration of the reference train loop (EmfLord.js:892-902): byUser
half-step, byItem half-step, RMSE(validate), RMSE(test), RMSE(test, shift applied), over
synthetic ratings of the named shape.  `value` = dataset ratings / seconds per iteration with
every input resident in HBM (bulk row sets); `e2e` = the same iteration driven through the
worker's per-portion messages with HOST portion buffers and the solved factor rows read back
into the host factor segments every half-step.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "als_ratings_per_sec_per_iteration"
UNIT = "ratings/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mal", choices=["mal", "netflix", "ml-1m", "ml-100k"])
    ap.add_argument("--factors", type=int, default=0)
    ap.add_argument("--e2e-portion", type=int, default=10_000,
                    help="ratingsInPortionForAls / ForRmse of the headline e2e leg (reference default: 10 000, EmfBase.js:97-103)")
    ap.add_argument("--e2e-large-portion", type=int, default=8_000_000, help="portion size of the secondary e2e leg (0 = skip)")
    ap.add_argument("--e2e-python-steps", type=int, default=1,
                    help="timed iterations of the e2e leg driven by Python worker messages at --e2e-portion (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused-peers", dest="fused_peers", action="store_true", default=True,
                    help="N > 1: the solve kernels store rows into all peer replicas over NVLink (default)")
    ap.add_argument("--nccl-exchange", dest="fused_peers", action="store_false",
                    help="N > 1: refresh the replicas with an NCCL all-gather after every half-step instead")
    ap.add_argument("--gram", default="auto", choices=["auto", "ffma", "tc"])
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every few milliseconds
    from a thread (a multi-GPU step lasts milliseconds — `nvidia-smi -lms 100` would not get a single sample in);
    nvidia-smi as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.samples, self.stop_flag, self.nvml = index, [], None, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        try:
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)     # constant: asked once
        except Exception:
            mx = 0
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [x[0] for x in self.samples]
            reasons = sorted(k for k, b in bits.items() if any(x[2] & b for x in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None,
                    "sm_max_mhz": max(x[1] for x in self.samples) if self.samples else None,
                    "samples": len(sm), "reasons": reasons, "source": "nvml, 2 ms poll"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def pipe_peaks():
    """TF32 tensor-pipe and FP32 FFMA peaks measured on a B200 of this pool with scripts/micro/tf32_mma_rate.cu and
    ffma_rate.cu (profiles/pipe_peaks.json); the profiling recipe's nominal dense figure when the file is missing."""
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "pipe_peaks.json")))
        return {"tf32_tflops": float(p["tf32_mma_tflops_n256"]), "ffma_tflops": float(p.get("ffma_tflops", 56.0)),
                "source": "measured (profiles/pipe_peaks.json)"}
    except Exception:
        return {"tf32_tflops": 1100.0, "ffma_tflops": 56.0, "source": "fallback (B200_PROFILING.md nominal dense tf32)"}


def smem_term(k, ratings_per_launch, n_mma, per_launch_ms, sm_count=148, sm_mhz=1965.0):
    """Shared-memory bandwidth term of gram_tc: executed bytes through the SM's shared memory
    per launch — LSU wavefronts per rating measured by ncu (profiles/gram_tc_smem.json) plus the tensor core's own
    operand reads — at 128 B/clk/SM.  Not an algorithmic figure: it explains why the kernel stops at ~0.55 of HBM."""
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "gram_tc_smem.json")))
        if int(p["k"]) != k:
            return None
        lsu = float(p["lsu_wavefronts_per_rating"]) * float(p["wavefront_bytes"])
        mma = (128 + n_mma) * 4.0
        peak_gbs = float(p["smem_bytes_per_clk_per_sm"]) * sm_count * sm_mhz * 1e6 / 1e9
        t = ratings_per_launch * (lsu + mma) / (peak_gbs * 1e9) * 1e3
        return {"executed_bytes_per_rating": {"lsu_wavefronts": lsu, "mma_operand_reads": mma}, "peak_gbs": peak_gbs,
                "ms_at_peak": t, "frac": t / per_launch_ms,
                "source": "MODEL: profiles/gram_tc_smem.json (ncu LSU wavefront count) + tensor-core operand footprint, assumed to share one 128 B/clk/SM port x %d SMs x %.0f MHz; ncu itself reports the LSU data pipe 79-81 %% busy (DESIGN.md 3.5)" % (sm_count, sm_mhz)}
    except Exception:
        return None


def measured_traffic(workload, k, cls):
    """DRAM bytes per launch of kernel class `cls` from the committed ncu --set full capture
    (profiles/traffic_<workload>.json, written by scripts/gpu_round_evidence.sh + DESIGN.md §5); None when
    there is no capture for this workload / rank."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)))
        if k != 100:
            return None
        return float(t[cls]["dram_bytes_per_launch"])
    except Exception:
        return None


def build_table(args):
    from you_can_not_recommend_b200 import front_end as fe
    t = fe.synth_table(args.workload)
    k = args.factors or fe.SHAPES[args.workload]["factors"]
    return t, k


class CpuArm:
    """The reference's CPU implementation of the path: per-row gather -> sgemm(T,N) -> +lambda*n*I -> sgemv -> sgesv
    (oracle O32 over OpenBLAS) with ALL host threads in the reference's own layout: one worker per
    core, each working on its own portions with single-threaded BLAS (numThreadsForTrain.als = numCPUs,
    EmfBase.js:105-110) — measured here 5-10x faster than one worker with a multi-threaded BLAS, whose k x k
    problems are too small to thread.  Timed on an evenly spread sample of 10k-rating portions of both half-steps
    and extrapolated linearly to the dataset."""

    def __init__(self, table, k):
        from oracle import oracle
        from you_can_not_recommend_b200 import front_end as fe
        from you_can_not_recommend_b200.emf_master import EmfMaster
        self.oracle, self.fe, self.table, self.k = oracle, fe, table, k
        self.cores = os.cpu_count() or 1
        self.have_blas = oracle.set_blas(threads=1)
        self.m = EmfMaster(table, {"factorsCount": k})
        self.m.splitDataForTrain()
        self.U = fe.init_factors(table.users, k, 0)
        self.V = fe.init_factors(table.items, k, 1)
        self.nnz12 = self.m._csr("byUser").nnz
        self.m._csr("byItem")

    def sample(self, seconds, offset=0):
        from concurrent.futures import ThreadPoolExecutor
        m, fe, oracle = self.m, self.fe, self.oracle
        per_rating, desc = {}, []
        with ThreadPoolExecutor(self.cores) as pool:
            for step in ("byUser", "byItem"):
                csr = m._csr(step)
                pto = m.portionsRowIdTo[step]
                fixed, solved = (self.V, self.U) if step == "byUser" else (self.U, self.V)
                n_por = len(pto)
                stride = max(1, n_por // (128 * self.cores))     # evenly spread over the id range
                mr, mrow = m.maxRatingsInPortion[step], m.maxRowsInPortion[step] + 1
                done_r, t_used, used = 0, 0.0, 0
                p = offset % stride

                def solve(bufs):
                    return oracle.als_portion(bufs[0], bufs[1], bufs[2], fixed, solved, 0.05, use_blas=self.have_blas)

                while p < n_por and t_used < seconds / 2:
                    batch = []                                   # a few portions per worker, converted outside the clock
                    while p < n_por and len(batch) < 8 * self.cores:
                        batch.append(fe.build_portion(csr, 0 if p == 0 else int(pto[p - 1]), int(pto[p]), mrow, mr)[:3])
                        p += stride
                    t0 = time.perf_counter()
                    done_r += sum(pool.map(solve, batch))
                    t_used += time.perf_counter() - t0
                    used += len(batch)
                per_rating[step] = t_used / max(done_r, 1)
                desc.append("%s: %d of %d portions (%d ratings, %.1f s)" % (step, used, n_por, done_r, t_used))
        iter_s = self.nnz12 * (per_rating["byUser"] + per_rating["byItem"])
        return {
            "value": self.table.nnz / iter_s, "unit": UNIT, "cores": self.cores, "kind": "port",
            "sample": "; ".join(desc) + "; extrapolated linearly to %d train+validate ratings; RMSE passes excluded" % self.nnz12,
            "threads": "%d workers x 1 BLAS thread (the reference's numThreadsForTrain.als = numCPUs layout)" % self.cores,
            "blas": "openblas(scipy)" if self.have_blas else "portable C loops",
            "iter_seconds_extrapolated": iter_s,
        }


def run_reference(args):
    """--impl reference: rank 0 times the CPU port; a step is one bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    table, k = build_table(args)
    arm = CpuArm(table, k)
    vals, cb = [], None
    per_step = max(1.0, args.cpu_seconds / max(1, args.steps))
    for i in range(args.warmup + args.steps):
        cb = arm.sample(per_step if i >= args.warmup else min(per_step, 2.0), offset=i)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = statistics.mean(vals) if vals else cb["value"]
    cb = dict(cb, value=v)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * table.nnz / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "users": table.users, "items": table.items, "ratings": table.nnz,
                   "factors": k, "note": "CPU port of the reference path (oracle O32 over OpenBLAS), sample extrapolated"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


_REAL_STDOUT = None


def protect_stdout():
    """Libraries print to fd 1 (NCCL's version banner, for one): route fd 1 to stderr for the whole run
    and keep the real stdout for the single JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    protect_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from you_can_not_recommend_b200 import dist as ydist
    from you_can_not_recommend_b200 import native
    from you_can_not_recommend_b200.emf_master import EmfMaster

    rank, local, world = ydist.init_from_env("nccl")
    if world != args.gpus:
        if rank == 0 and world == 1 and args.gpus > 1:
            sys.stderr.write("bench.py: --gpus %d needs torchrun with %d ranks\n" % (args.gpus, args.gpus))
            sys.exit(2)
    if native.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device — the ALS hot path has no CPU fallback")
    torch.cuda.set_device(local)
    table, k = build_table(args)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---------------- device-resident iterations (value) ----------------
    opts = {"factorsCount": k, "gpu": {"bulk": True, "profile": True, "device": local, "gramPath": args.gram,
                                       "tcMinCols": int(os.environ.get("YCNR_TC_MIN", "0"))}}
    m = EmfMaster(table, opts, rank=rank, world=world)
    m.prepareToTrain()
    if world > 1 and args.fused_peers:
        m.connectPeers()
    stream = torch.cuda.ExternalStream(m.ctx.stream_ptr(), device=torch.device("cuda", local))
    for _ in range(args.warmup):
        m.trainIter()
    m.ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    m.ctx.profile_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    hist = []
    for _ in range(args.steps):
        hist.append(m.trainIter())
    e1.record(stream)
    m.ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof = m.ctx.profile_read()
    ms = max(dev_ms, 0.0)
    if world > 1:
        tmax = torch.tensor([ms, wall_ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, wall_ms = float(tmax[0]), float(tmax[1])
        launches = torch.tensor([prof["total_launches"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(launches)
        total_launches = int(launches[0])
    else:
        total_launches = int(prof["total_launches"])
    ms_per_step = ms / args.steps
    value = table.nnz / (ms_per_step / 1e3)
    nnz_by = {s: int(m.rowlists[s][1].sum(dtype=np.int64)) for s in ("byUser", "byItem")}

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peaks()
    gram_classes = ["primal_fused", "dual_fused", "gram_partial", "gram_tc"]
    # The dominant KERNEL: dual_fused is a class of up to 24 different kernels (one template instantiation per
    # tile-row count, each launched once per half-step), so it competes per instantiation; every other class is
    # one kernel launched once per half-step / pass.
    def kernel_ms(c):
        if c == "dual_fused" and prof[c]["launches"]:
            return prof[c]["ms"] / max(1.0, prof[c]["launches"] / (args.steps or 1))
        return prof[c]["ms"] / (args.steps or 1)
    # (classes that carry no ratings of their own — the k x k solve reads tile partials — cannot be put against a
    #  per-rating roofline; they stay in `kernels`)
    dom = max((c for c in native.KERNEL_CLASSES if prof[c]["ratings"] > 0), key=kernel_ms, default=max(native.KERNEL_CLASSES, key=kernel_ms))
    roof = None
    pk = pipe_peaks()
    if prof[dom]["launches"] > 0 and prof[dom]["ms"] > 0:
        per_launch_ms = prof[dom]["ms"] / prof[dom]["launches"]
        ratings_per_launch = prof[dom]["ratings"] / prof[dom]["launches"]
        alg_bytes = ratings_per_launch * k * 4                  # B_gather = nnz * k * 4
        alg_flops = ratings_per_launch * k * k                  # F_gram = nnz * k^2 (symmetric-aware)
        # what the kernel really executes per rating: the tensor-core Gram issues M = 128 x N = 2 NC MMAs (K = 8 ratings
        # each) whatever k is; the FFMA kernels execute the padded lower triangle of 4 x 4 tiles
        kp = 4 * ((k + 3) // 4)
        if dom == "gram_tc":
            kv = kp if k <= 124 else 2 * {3: 44}.get((k + 59) // 60, 52 if ((k + 59) // 60) * 52 >= k else 60)
            exec_flops = ratings_per_launch * 2.0 * 128 * 2 * ((kv + 4 + 7) // 8 * 8)
            pipe, pipe_peak = "tensor (tcgen05 kind::tf32)", pk["tf32_tflops"]
            n_mma = 2 * ((kv + 4 + 7) // 8 * 8)
        else:
            exec_flops = ratings_per_launch * 2.0 * (kp * (kp + 4) / 2 + kp)
            pipe, pipe_peak = "fp32 ffma", pk["ffma_tflops"]
        t_hbm = alg_bytes / (peak * 1e9) * 1e3
        t_pipe = exec_flops / (pipe_peak * 1e12) * 1e3
        bound = "hbm" if t_hbm >= t_pipe else "tensor"
        if bound == "hbm":
            achieved, rpeak, unit = alg_bytes / (per_launch_ms * 1e-3) / 1e9, peak, "GB/s"
        else:
            achieved, rpeak, unit = alg_flops / (per_launch_ms * 1e-3) / 1e12, pipe_peak * alg_flops / exec_flops, "TFLOP/s"
        step_bytes = sum(nnz_by.values()) * k * 4               # both half-steps of the iteration, algorithmic
        roof = {"bound": bound, "kernel": dom, "achieved": achieved, "peak": rpeak, "unit": unit,
                "frac": achieved / rpeak, "traffic": measured_traffic(args.workload, k, dom) if world == 1 else None,
                "traffic_source": "profiles/traffic_%s.json (ncu --set full, dram read+write per launch)" % args.workload,
                "peak_source": (peak_src + " (MEASURED_PEAKS.json hbm_gbs)") if bound == "hbm" else
                               pk["source"] + ", scaled by algorithmic / executed flops",
                "launch_ms": per_launch_ms, "launches_timed": prof[dom]["launches"],
                "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_flops_per_launch": alg_flops,
                "terms": {"hbm": {"ms_at_peak": t_hbm, "peak_gbs": peak, "frac": t_hbm / per_launch_ms},
                          "pipe": {"pipe": pipe, "executed_flops_per_launch": exec_flops, "peak_tflops": pipe_peak,
                                   "executed_over_algorithmic": exec_flops / alg_flops if alg_flops else None, "ms_at_peak": t_pipe,
                                   "frac": t_pipe / per_launch_ms, "peak_source": pk["source"]}},
                "roofline_ms": max(t_hbm, t_pipe), "frac_of_max_term": max(t_hbm, t_pipe) / per_launch_ms,
                "share_of_step": prof[dom]["ms"] / ms if ms > 0 else None,
                "whole_step": {"algorithmic_bytes": step_bytes, "ms": ms_per_step,
                               "achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                               "note": "gathered factor bytes of both half-steps on rank 0 / iteration time (RMSE passes included in the time) / HBM peak"}}
        if dom == "gram_tc":
            st = smem_term(k, ratings_per_launch, n_mma, per_launch_ms)
            if st:
                roof["terms"]["smem"] = st
    kernels = {c: {"ms_per_step": prof[c]["ms"] / args.steps, "launches_per_step": prof[c]["launches"] / args.steps,
                   "rows_per_step": prof[c]["rows"] / args.steps, "ratings_per_step": prof[c]["ratings"] / args.steps}
               for c in native.KERNEL_CLASSES if prof[c]["launches"]}
    last = hist[-1] if hist else {}
    m.endTrain()
    del m

    # ---------------- e2e: the worker's per-portion entry points with host portion buffers ----------------
    def run_e2e(portion, native_loop, n_steps, n_warm):
        opts2 = {"factorsCount": k,
                 "ratingsInPortionForAls": {"byUser": portion, "byItem": portion},
                 "ratingsInPortionForRmse": portion,
                 "gpu": {"bulk": False, "profile": False, "device": local, "gramPath": args.gram,
                         "cachePortions": True, "nativeLoop": native_loop}}
        m2 = EmfMaster(table, opts2, rank=rank, world=world)
        m2.prepareToTrain()
        if world > 1 and args.fused_peers:
            m2.connectPeers()
        for _ in range(n_warm):
            m2.trainIter()
        m2.ctx.synchronize()
        torch.cuda.synchronize()
        barrier()
        m2.h2d_bytes = m2.d2h_bytes = 0
        m2.phase_ms = {}
        t0 = time.perf_counter()
        for _ in range(n_steps):
            last2 = m2.trainIter()
        m2.ctx.synchronize()
        torch.cuda.synchronize()
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_steps
        h2d, d2h = m2.h2d_bytes / n_steps, m2.d2h_bytes / n_steps
        calls = sum(m2.my_portions[s_][1] - m2.my_portions[s_][0] for s_ in ("byUser", "byItem", "rmseValidate", "rmseTest", "rmseTest"))
        if world > 1:
            tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt[0])
            bb = torch.tensor([h2d, d2h, calls], device="cuda", dtype=torch.float64)
            dist.all_reduce(bb)
            h2d, d2h, calls = float(bb[0]), float(bb[1]), int(bb[2])
        out = {"value": table.nnz / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_s * 1e3, "steps": n_steps, "ratings_in_portion": portion,
               "portion_calls_per_step": calls,
               "phase_ms": {k_: v_ / n_steps for k_, v_ in m2.phase_ms.items()},
               "rmse": last2,
               "api": ("ycnr_als_portion / ycnr_rmse_portion_async, one call per portion, issued by a native loop "
                       "(ycnr_als_portions / ycnr_rmse_portions_async)") if native_loop else
                      "EmfWorker calcTrainAlsPortion / calcRmsePortion messages (Python mirror) -> ycnr_als_portion / ycnr_rmse_portion_async",
               "inputs": "converted portions cached in page-locked host memory (usePortionsCache taken to its limit: the "
                         "conversion of EmfMaster.js:571-614 is outside the timed region); H2D of every portion and D2H of "
                         "the solved rows into the host factor segments inside it"}
        m2.endTrain()
        del m2
        return out

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args.e2e_portion, True, max(1, args.steps), max(1, args.warmup))
        if args.e2e_large_portion and world == 1:      # (a few 8 M-rating portions cannot be balanced over ranks)
            e2e["large_portions"] = run_e2e(args.e2e_large_portion, False, max(1, min(args.steps, 3)), max(1, args.warmup))
        if args.e2e_python_steps:
            e2e["python_messages"] = run_e2e(args.e2e_portion, False, args.e2e_python_steps, max(1, min(args.warmup, 2)))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = CpuArm(table, k).sample(args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "users": table.users, "items": table.items, "ratings": table.nnz,
                       "factors": k, "train_ratings_per_half_step": nnz_by, "split": [85, 10, 5],
                       "ratings_in_portion": 10000, "parallelism": "rows nnz-balanced over %d GPU(s)" % world,
                       "replica_refresh": ("peer stores from the solve kernels" if args.fused_peers else "NCCL broadcast per rank slice") if world > 1 else "none",
                       "l2": "inputs_exceed_l2 (CSR + factors >> 126 MB)", "gram_path": args.gram,
                       "kernels_note": "per-class times are CUDA-event intervals; row sets under 400k rows per rank run their dual bins on three streams, so those intervals overlap and add up to more than the step",
                       "step": "byUser + byItem + 3 RMSE passes"},
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": total_launches,
            "clocks": clocks, "rmse": last,
        }
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
