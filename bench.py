#!/usr/bin/env python3
"""Benchmark of the B200-native Demons registration hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--field-mode fp64|fp64_strict|mixed|fp32]

A "step" is one pass of the hot path over one batch of synthetic input: one Demons registration of `--iters`
iterations (default 200) on the BASELINE config-2 workload -- synthetic CT pair 512x512x256, diffeomorphic,
sigma_d = sigma_u = 1.5 mm, convergence threshold 0 (exactly `iters` iterations run).  The metric is voxel-updates/s
(voxels x iterations executed / time).

  value     : loop-only throughput, inputs resident in HBM, device-timed (CUDA events on the engine's stream),
              max over ranks.
              N = 1 : BASELINE config 2 (512x512x256, 200 iterations per step).
              N > 1 : BASELINE config 4 -- ONE registration of 1024^3 sharded into z-slabs over the N ranks (halo
                      exchange over NVLink every iteration), "scaling": "strong".  The throughput of N independent
                      config-2 registrations (no data-path collective) is reported beside it as `replicas`, and
                      `--replicas` makes that the headline instead.
  e2e       : the same metric through the reference-facing call itself, the C++ drop-in AlignViaDemons(params, moving,
              fixed) behind host/dropin_harness.cc: image stacks in, deformation_field out; marshalling both ways, GPU
              resampling, H2D of fixed+moving, the loop and D2H of the fp64 field are all inside the timed region.  The
              C-ABI one-shot call `dcma_b200_demons_register` with pinned caller buffers is timed beside it (`c_abi`).
  roofline  : per-kernel table from the CUDA-event stage profile (algorithmic bytes per launch / mean launch time);
              the headline object is the kernel with the largest share of the step.
  cpu_baseline : the reference's own engine (oracle/_ref, built from /root/reference sources) on the host cores.

`--impl reference` times the reference CPU engine itself on a bounded sample of the same workload.
Only the cpu_baseline / --impl reference legs touch `oracle/`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"shape": (256, 512, 512), "spacing": (1.0, 1.0, 1.0), "sigma_d": 1.5, "sigma_u": 1.5,
            "diffeomorphic": True}
# algorithmic bytes per voxel-update, SURVEY.md section 8(d): diffeomorphic, C=1
# ("mixed": D and the gradient in double, U in float: force F4+W4+G24+U12 = 44, smooth(U) 24, compose D24+U12+D24 = 60,
#  smooth(D) 48, warp D24+M4+W4 = 32 -> 208)
B_ALG = {"fp32": 136.0, "fp64": 256.0, "fp64_strict": 256.0, "mixed": 208.0}
# storage bytes per element of (the deformation field D, the update field U)
FIELD_BYTES = {"fp32": (4, 4), "fp64": (8, 8), "fp64_strict": (8, 8), "mixed": (8, 4)}
DTYPE = {"fp32": "f32", "mixed": "f64 (update field stored as f32)", "fp64": "f64", "fp64_strict": "f64"}


def kernel_table(mode, nvox, pst, capi):
    """Per-kernel roofline rows from the per-stage CUDA-event profile of one step.  Algorithmic bytes per launch are the
    bytes the kernel has to move by design: force D+M4+F4+U, smooth 2 x field, compose D+U+D."""
    bd, bu = FIELD_BYTES[mode]
    smooth_name = {8: "k_smooth3d_tma<double> (TMA-staged separable 3-D Gaussian, one launch per field smooth)",
                   4: "k_smooth3d<float> (separable 3-D Gaussian, one launch per field smooth)"}
    spec = [(capi.STAGE_FORCE, "k_warp_force (warp of M through D + fixed gradient + force + MSE)", 3 * bd + 4 + 4 + 3 * bu),
            (capi.STAGE_SMOOTH_U, smooth_name[bu], 2 * 3 * bu),
            (capi.STAGE_COMPOSE, "k_compose (diffeomorphic composition)", 2 * 3 * bd + 3 * bu),
            (capi.STAGE_ADD, "k_add_update", 2 * 3 * bd + 3 * bu),
            (capi.STAGE_SMOOTH_D, smooth_name[bd], 2 * 3 * bd)]
    rows = {}
    for st, name, bpv in spec:
        n = int(pst.stage_launches[st])
        if n <= 0:
            continue
        r = rows.setdefault(name, {"kernel": name, "ms": 0.0, "launches": 0, "bytes": 0.0})
        r["ms"] += float(pst.stage_ms[st])
        r["launches"] += n
        r["bytes"] += float(bpv) * nvox * n
    return list(rows.values())


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_baseline(budget_s: float = 25.0, iters_timed: int = 3):
    """Times the reference engine (oracle/_ref) on the host cores on a bounded sample of the workload.

    Returns dict(value, unit, cores, kind, sample, ms_per_iter).  Test infrastructure used as a reported baseline."""
    import numpy as np
    import oracle
    from dicomautomaton_b200 import phantom
    kind = "reference" if oracle.available("reference") else "port"
    okind = "reference" if kind == "reference" else "port"
    cores = oracle.worker_threads(okind)
    par = oracle.make_params(max_iterations=1, convergence_threshold=0.0,
                             deformation_field_smoothing_sigma=WORKLOAD["sigma_d"],
                             update_field_smoothing_sigma=WORKLOAD["sigma_u"],
                             use_diffeomorphic=WORKLOAD["diffeomorphic"])
    # calibrate on 64x128x128 (1/64 of the workload), then pick the largest sample that fits the budget
    f, m = phantom.make_pair((64, 128, 128))
    eng = oracle.OracleEngine(f.numpy(), m.numpy(), par, kind=okind)
    eng.iterate()
    t = time.perf_counter()
    eng.iterate()
    per_vox = (time.perf_counter() - t) / (64 * 128 * 128)
    eng.close()
    full = WORKLOAD["shape"]
    shape = None
    for div in (1, 2, 4):
        cand = (full[0] // div, full[1] // div, full[2] // div)
        est = per_vox * cand[0] * cand[1] * cand[2] * (iters_timed + 1.5)
        mem_gb = cand[0] * cand[1] * cand[2] * 220 / 1e9
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 1e9
        except Exception:
            avail_gb = 16.0
        if est <= budget_s and mem_gb < 0.6 * avail_gb:
            shape = cand
            break
    if shape is None:
        shape = (full[0] // 4, full[1] // 4, full[2] // 4)
    dev = "cpu"
    try:
        import torch
        if torch.cuda.is_available():
            dev = "cuda"
    except Exception:
        pass
    f, m = phantom.make_pair(shape, device=dev)
    f, m = f.cpu().numpy(), m.cpu().numpy()
    eng = oracle.OracleEngine(f, m, par, kind=okind)
    eng.iterate()  # warm-up (first iteration also has W = M)
    t = time.perf_counter()
    for _ in range(iters_timed):
        eng.iterate()
    dt = time.perf_counter() - t
    eng.close()
    nvox = shape[0] * shape[1] * shape[2]
    return {"value": nvox * iters_timed / dt, "unit": "voxel-updates/s", "cores": cores, "kind": kind,
            "sample": f"{shape[2]}x{shape[1]}x{shape[0]} sub-volume of the workload phantom, same parameters, "
                      f"1 warm-up + {iters_timed} timed iterations of compute_single_iteration",
            "ms_per_iter": 1e3 * dt / iters_timed}


def run_reference(args):
    """--impl reference: the reference's own CPU engine on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, ms = [], []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_reference_baseline(budget_s=max(8.0, 150.0 / max(1, args.warmup + args.steps)), iters_timed=2)
        if i >= args.warmup:
            vals.append(base["value"])
            ms.append(base["ms_per_iter"])
    v = sum(vals) / len(vals)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg = workload_config(args, "fp64") if world <= 1 or args.replicas else slab_config(args, "fp64", world, SLAB_SHAPE)
    cfg["reference_sample"] = base["sample"]
    line = {"impl": "reference", "metric": "demons_voxel_updates_per_s", "value": v, "unit": "voxel-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / len(ms) * 2,
            "iterations_timed": 2,  # per step: 1 warm-up + 2 timed iterations of the reference engine on the sample
            "higher_is_better": True, "scaling": "weak" if world <= 1 or args.replicas else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "voxel-updates/s", "cores": base["cores"], "kind": base["kind"],
                             "sample": base["sample"]},
            "e2e": {"value": v, "unit": "voxel-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


SLAB_SHAPE = (1024, 1024, 1024)  # BASELINE config 4
SLAB_ITERS = 50                   # iterations per step of the sharded run (config 4 asks for >= 20)
CHECK_SHAPE = (256, 256, 256)     # slab-vs-whole bit-identity check (a size every rank can also run whole)


def slab_config(args, mode, world, shape):
    return {"workload": f"BASELINE config 4: one synthetic CT pair {shape[2]}x{shape[1]}x{shape[0]} (textured phantom + known "
                        f"sinusoidal warp) sharded into {world} z-slabs, diffeomorphic Demons, sigma_d=sigma_u="
                        f"{WORKLOAD['sigma_d']} mm (r=4, 9 taps), {args.slab_iters} iterations per step, threshold 0",
            "iterations_per_step": args.slab_iters, "field_mode": mode,
            "parallelism": f"z-slabs over {world} GPU(s), one rank per GPU, moving image replicated; per iteration the "
                           f"z-neighbours exchange r_z planes of U, halo planes of smoothed U and r_z planes of D over "
                           f"NVLink (IPC-mapped peer buffers, copy engines); MSE terms combined after the loop",
            "l2": "working set (tens of GB of fields per rank) is far larger than the 126 MB L2; no flush needed"}


def workload_config(args, mode):
    s = WORKLOAD["shape"]
    return {"workload": f"BASELINE config 2: synthetic CT pair {s[2]}x{s[1]}x{s[0]} (textured phantom + known sinusoidal "
                        f"warp), diffeomorphic Demons, sigma_d=sigma_u={WORKLOAD['sigma_d']} mm (r=4, 9 taps), "
                        f"{args.iters} iterations per step, threshold 0",
            "iterations_per_step": args.iters, "field_mode": mode,
            "parallelism": "single GPU" if int(os.environ.get("WORLD_SIZE", "1")) == 1 else
                           f"{os.environ.get('WORLD_SIZE')} independent registrations, one per GPU (no data-path collective)",
            "l2": "working set (>= 3 GB of fields) is far larger than the 126 MB L2; no flush needed"}


def run_ours(args):
    import numpy as np
    import torch
    from dicomautomaton_b200 import capi, phantom
    from dicomautomaton_b200.demons import AlignViaDemonsParams, DemonsEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    mode = args.field_mode
    shape, spacing = WORKLOAD["shape"], WORKLOAD["spacing"]
    nvox = shape[0] * shape[1] * shape[2]

    fixed, moving = phantom.make_pair(shape, spacing=spacing, case=rank, device=dev)
    torch.cuda.synchronize()
    par = AlignViaDemonsParams(max_iterations=args.iters, convergence_threshold=0.0,
                               deformation_field_smoothing_sigma=WORKLOAD["sigma_d"],
                               update_field_smoothing_sigma=WORKLOAD["sigma_u"],
                               use_diffeomorphic=WORKLOAD["diffeomorphic"], verbosity=0, field_mode=mode,
                               device=local)
    stream = torch.cuda.current_stream()
    # engine on torch's current stream so torch.cuda.Event brackets exactly its kernels
    eng = DemonsEngine.empty(shape, spacing, par, stream=stream.cuda_stream)

    def step():
        eng.load_device(fixed.data_ptr(), moving.data_ptr())
        hist, st = eng.run(args.iters, profile=False)
        return hist, st

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    iters_done = 0
    e0.record(stream)
    for _ in range(args.steps):
        hist, st = step()
        launches += st.kernel_launches
        iters_done += st.iterations_done
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    final_mse = float(hist[-1]) if len(hist) else float("nan")
    t = torch.tensor([ms, float(iters_done), float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        iters_all, launches_all = float(tsum[1]), int(tsum[2])
    else:
        iters_all, launches_all = float(iters_done), int(launches)
    value = nvox * iters_all / (ms * 1e-3)

    # ---- per-stage profile (one extra step, events around every kernel) -> roofline of the dominant kernel
    eng.load_device(fixed.data_ptr(), moving.data_ptr())
    _, pst = eng.run(args.iters, profile=True)
    stage = {capi.STAGE_NAMES[i]: {"ms": pst.stage_ms[i], "launches": int(pst.stage_launches[i])} for i in range(8)
             if pst.stage_launches[i] > 0}
    peak, peak_src = measured_peak_gbs()
    rows = kernel_table(mode, nvox, pst, capi)
    for r in rows:
        r["ms_per_launch"] = r["ms"] / r["launches"]
        r["algorithmic_bytes_per_launch"] = r["bytes"] / r["launches"]
        r["achieved_GBs"] = r["bytes"] / (r["ms"] * 1e-3) / 1e9
        r["frac"] = r["achieved_GBs"] / peak
        r["share_of_step"] = r["ms"] / max(pst.loop_ms, 1e-9)
        del r["bytes"]
    dom = max(rows, key=lambda r: r["ms"])  # the dominant kernel: largest share of the step
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        try:
            per = json.load(open(tp)).get(mode, {}).get("per_kernel_dram_bytes_per_launch", {})
            traffic = next((v for k, v in per.items() if dom["kernel"].startswith(k)), None)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom["kernel"],
                "achieved": dom["achieved_GBs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic,
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                "ms_per_launch": dom["ms_per_launch"], "share_of_step": dom["share_of_step"],
                "per_kernel": rows,
                "whole_iteration": {"algorithmic_bytes_per_voxel_update": B_ALG[mode],
                                    "achieved_GBs": value / max(world, 1) * B_ALG[mode] / 1e9,
                                    "frac_of_peak": value / max(world, 1) * B_ALG[mode] / 1e9 / peak},
                "stages": stage}
    eng.close()

    # ---- end to end through the C-ABI one-shot call with pinned HOST buffers
    lib = capi.lib()
    import ctypes as C
    hf = fixed.cpu().pin_memory()
    hm = moving.cpu().pin_memory()
    hD = torch.empty((nvox, 3), dtype=torch.float64).pin_memory()
    geom = capi.Geom(shape[0], shape[1], shape[2], 1, *spacing)
    cp = par.to_c()
    err = C.create_string_buffer(512)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    stats = capi.RunStats()

    def e2e_step():
        capi.check(lib.dcma_b200_demons_register(C.byref(geom), hf.data_ptr(), hm.data_ptr(), C.byref(cp),
                                                 hD.data_ptr(), None, C.byref(stats), err, len(err)), err)
    e2e_step()  # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(e2e_steps):
        e2e_step()
        e2e_iters += stats.iterations_done
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(e2e_iters)], dtype=torch.float64, device=dev)
    if dist is not None:
        a = te.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone()
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_iters = float(a[0]), float(b[1])
    e2e_capi = {"value": nvox * e2e_iters / e2e_s, "unit": "voxel-updates/s",
                "h2d_bytes_per_step": int(stats.h2d_bytes), "d2h_bytes_per_step": int(stats.d2h_bytes),
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "h2d_ms": stats.h2d_ms, "d2h_ms": stats.d2h_ms, "loop_ms": stats.loop_ms,
                "api": "dcma_b200_demons_register (C ABI, host pinned buffers, fp64 AoS deformation out)"}
    probe = (shape[0] // 2 * shape[1] + shape[1] // 3) * shape[2] + shape[2] // 3
    d_capi = hD[probe].clone()
    del hD

    # ---- end to end through the reference-facing call itself: the C++ drop-in AlignViaDemons(params, moving, fixed)
    # (host/Alignment_Demons.cc behind host/dropin_harness.cc): image stacks in, deformation_field out; both marshalling
    # steps, the GPU grid resampling, H2D/D2H and the construction of the deformation_field are inside the timed region
    e2e = dict(e2e_capi)
    other_e2e = {}
    harness = os.path.join(ROOT, "dicomautomaton_b200", "host", "_build", "libdcma_b200_dropin.so")
    if os.path.isfile(harness) and not args.no_dropin:
        dl = C.CDLL(harness)
        dl.dcma_dropin_create.restype = C.c_void_p
        dl.dcma_dropin_create.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]
        dl.dcma_dropin_run.restype = C.c_int
        dl.dcma_dropin_run.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int64,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]
        dl.dcma_dropin_destroy.argtypes = [C.c_void_p]
        hnd = dl.dcma_dropin_create(shape[0], shape[1], shape[2], hf.data_ptr(), hm.data_ptr(), *spacing)
        if hnd:
            secs, out3 = C.c_double(0), (C.c_double * 3)()

            def dropin_step():
                rc = dl.dcma_dropin_run(hnd, args.iters, 0.0, WORKLOAD["sigma_d"], WORKLOAD["sigma_u"],
                                        1 if WORKLOAD["diffeomorphic"] else 0, capi.FIELD_MODES[mode], local, probe,
                                        C.byref(secs), out3)
                if rc != 0:
                    raise SystemExit("AlignViaDemons (drop-in) failed")
                return secs.value
            dropin_step()  # warm-up: fills the pinned-buffer pool
            barrier()
            n_steps = max(1, args.dropin_steps)
            tot = sum(dropin_step() for _ in range(n_steps))
            td = torch.tensor([tot], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            tot = float(td[0])
            same = all(abs(out3[k] - float(d_capi[k])) <= 1e-12 for k in range(3))
            e2e = {"value": nvox * args.iters * n_steps * world / tot, "unit": "voxel-updates/s",
                   "h2d_bytes_per_step": int(stats.h2d_bytes) + nvox * 4,  # + the moving volume on its own grid (GPU resampling)
                   "d2h_bytes_per_step": int(stats.d2h_bytes) + nvox * 4, "steps": n_steps, "ms_per_step": 1e3 * tot / n_steps,
                   "api": "AlignViaDemons(params, moving, fixed) -- the C++ drop-in (host/Alignment_Demons.cc): image stacks in, "
                          "deformation_field out; marshalling both ways, GPU resampling onto the fixed grid, pooled pinned "
                          "staging, H2D/D2H all inside the timed region",
                   "agrees_with_c_abi_at_probe_voxel": bool(same),
                   "c_abi": e2e_capi, "ratio_to_c_abi": (nvox * args.iters * n_steps * world / tot) / e2e_capi["value"]}
            # the same call in the other in-tolerance mode(s), a few steps each (reported under other_modes)
            for om in [m for m in args.other_modes.split(",") if m and m != mode and m != "fp32"]:
                if args.no_other_modes:
                    break
                try:
                    def om_step(field_mode=capi.FIELD_MODES[om]):
                        rc = dl.dcma_dropin_run(hnd, args.iters, 0.0, WORKLOAD["sigma_d"], WORKLOAD["sigma_u"],
                                                1 if WORKLOAD["diffeomorphic"] else 0, field_mode, local, probe, C.byref(secs), out3)
                        if rc != 0:
                            raise RuntimeError("AlignViaDemons (drop-in) failed")
                        return secs.value
                    om_step()
                    barrier()
                    k = 3
                    tt = torch.tensor([sum(om_step() for _ in range(k))], dtype=torch.float64, device=dev)
                    if dist is not None:
                        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    other_e2e[om] = {"value": nvox * args.iters * k * world / float(tt[0]), "unit": "voxel-updates/s", "steps": k,
                                     "ms_per_step": 1e3 * float(tt[0]) / k, "api": "AlignViaDemons (C++ drop-in), as the headline e2e"}
                except Exception as ex:  # an extra figure: never takes the headline down with it
                    other_e2e[om] = {"value": None, "error": str(ex)}
            dl.dcma_dropin_destroy(hnd)

    other = {}
    if not args.no_other_modes:
        for om in [m for m in args.other_modes.split(",") if m]:
            if om == mode:
                continue
            opar = AlignViaDemonsParams(max_iterations=args.iters, convergence_threshold=0.0,
                                        deformation_field_smoothing_sigma=WORKLOAD["sigma_d"],
                                        update_field_smoothing_sigma=WORKLOAD["sigma_u"],
                                        use_diffeomorphic=WORKLOAD["diffeomorphic"], verbosity=0, field_mode=om,
                                        device=local)
            oeng = DemonsEngine.empty(shape, spacing, opar, stream=stream.cuda_stream)
            n_it = min(args.iters, 50)
            oeng.load_device(fixed.data_ptr(), moving.data_ptr())
            oeng.run(5)  # warm-up
            oeng.load_device(fixed.data_ptr(), moving.data_ptr())
            _, ost = oeng.run(n_it)
            oeng.close()
            ov = nvox * ost.iterations_done / (ost.loop_ms * 1e-3)
            other[om] = {"value_per_gpu": ov, "unit": "voxel-updates/s", "ms_per_iteration": ost.loop_ms / max(ost.iterations_done, 1),
                         "iterations": int(ost.iterations_done),
                         "whole_iteration_frac_of_peak": ov * B_ALG[om] / 1e9 / peak,
                         "algorithmic_bytes_per_voxel_update": B_ALG[om]}
            if om in other_e2e:
                other[om]["e2e"] = other_e2e[om]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_baseline()
        except Exception as ex:  # the baseline is reported context, never the product path
            cpu = {"value": None, "unit": "voxel-updates/s", "cores": None, "kind": "unavailable", "sample": str(ex)}

    if rank == 0:
        line = {"metric": "demons_voxel_updates_per_s", "value": value, "unit": "voxel-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPE[mode],
                "data": "synthetic", "config": workload_config(args, mode),
                "ms_per_iteration": ms / max(iters_done, 1), "final_mse": final_mse,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_all, "roofline": roofline,
                "cpu_baseline": cpu, "other_modes": other,
                "parity": "vs fixtures written by the reference engine compiled from its own sources, same inputs, "
                          "512x512x256, 200 iterations (tests/test_gpu_full_length.py, profiles/r2_full_length_parity.jsonl): "
                          "fp64_strict bit-identical (sha256 of the whole field); fp64 max |dD| 2.8e-7 mm, final MSE 3.8e-10 "
                          "relative; mixed 2.8e-7 mm / 7.5e-9; fp32 1.1e-3 mm (99.99th percentile 3.1e-6 mm) / 9.6e-7.  "
                          "Tolerance: 1e-3 mm / 1e-4."}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _replicas_throughput(args, dist, dev, rank, world, local, stream):
    """N independent config-2 registrations, one per GPU, no data-path collective (BASELINE config 5 style): a short
    device-timed run, max over ranks.  Reported beside the sharded headline."""
    import torch
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import AlignViaDemonsParams, DemonsEngine
    shape, spacing = WORKLOAD["shape"], WORKLOAD["spacing"]
    n_it = min(args.iters, 50)
    f, m = phantom.make_pair(shape, spacing=spacing, case=rank, device=dev)
    par = AlignViaDemonsParams(max_iterations=n_it, convergence_threshold=0.0,
                               deformation_field_smoothing_sigma=WORKLOAD["sigma_d"],
                               update_field_smoothing_sigma=WORKLOAD["sigma_u"],
                               use_diffeomorphic=WORKLOAD["diffeomorphic"], verbosity=0, field_mode=args.field_mode,
                               device=local)
    eng = DemonsEngine.empty(shape, spacing, par, stream=stream.cuda_stream)
    eng.load_device(f.data_ptr(), m.data_ptr())
    eng.run(5)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    eng.load_device(f.data_ptr(), m.data_ptr())
    _, st = eng.run(n_it)
    eng.close()
    t = torch.tensor([st.loop_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nvox = shape[0] * shape[1] * shape[2]
    del f, m
    return {"value": nvox * st.iterations_done * world / (float(t[0]) * 1e-3), "unit": "voxel-updates/s",
            "workload": f"{world} independent registrations of {shape[2]}x{shape[1]}x{shape[0]} (config-2 parameters), one per "
                        f"GPU, {int(st.iterations_done)} iterations, no data-path collective", "scaling": "weak"}


def _slab_bit_identity(args, dist, dev, rank, world, local):
    """The sharded run against the whole-volume engine at a size every rank can also run whole: the owned slices of the
    slab result must equal the whole-volume field BIT FOR BIT (halo exchange must not change any voxel's arithmetic)."""
    import numpy as np
    import torch
    from dicomautomaton_b200 import phantom
    from dicomautomaton_b200.demons import AlignViaDemonsParams, DemonsEngine, DemonsSlabGroup
    shape, iters = CHECK_SHAPE, 10
    f, m = phantom.make_pair(shape, device=dev)
    par = AlignViaDemonsParams(max_iterations=iters, convergence_threshold=0.0,
                               deformation_field_smoothing_sigma=WORKLOAD["sigma_d"],
                               update_field_smoothing_sigma=WORKLOAD["sigma_u"], use_diffeomorphic=True, verbosity=0,
                               field_mode=args.field_mode, device=local)
    eng = DemonsEngine.empty(shape, (1.0, 1.0, 1.0), par)
    eng.load_device(f.data_ptr(), m.data_ptr())
    hw, _ = eng.run(iters)
    whole = eng.get("deformation")
    eng.close()
    ident = [DemonsSlabGroup.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    grp = DemonsSlabGroup.nccl(shape, (1.0, 1.0, 1.0), par, ident[0], rank, world, local)
    grp.load_device(f.data_ptr(), m.data_ptr())
    hs, _ = grp.run(iters)
    z0, z1 = grp.owned_range()
    mine = grp.get_deformation()
    stats = grp.exchange_stats()
    grp.close()
    same = bool(np.array_equal(mine, whole[z0:z1]))
    mse_rel = float(np.max(np.abs(hs - hw) / np.abs(hw))) if len(hs) == len(hw) and len(hw) else float("nan")
    t = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    del f, m
    return {"shape": f"{shape[2]}x{shape[1]}x{shape[0]}", "iterations": iters, "bit_identical_on_all_ranks": bool(t[0] == 1.0),
            "mse_history_max_rel_diff": mse_rel, "overlapped": stats["overlapped"], "transport": stats["transport"]}


def run_slab(args):
    """N > 1 (default) or --slab: ONE registration sharded into z-slabs over the ranks (BASELINE config 4): halo exchange
    over NVLink inside the loop, strong scaling (the total work is fixed).  Device-timed per rank by the engine's own CUDA
    events around the loop, max over ranks."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from dicomautomaton_b200 import capi, phantom
    from dicomautomaton_b200.demons import AlignViaDemonsParams, DemonsSlabGroup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shape = tuple(int(v) for v in args.shape.split(",")) if args.shape else SLAB_SHAPE
    spacing = WORKLOAD["spacing"]
    nvox = shape[0] * shape[1] * shape[2]
    iters = args.slab_iters
    mode = args.field_mode
    fixed, moving = phantom.make_pair(shape, spacing=spacing, device=dev)
    torch.cuda.synchronize()
    par = AlignViaDemonsParams(max_iterations=iters, convergence_threshold=0.0,
                               deformation_field_smoothing_sigma=WORKLOAD["sigma_d"],
                               update_field_smoothing_sigma=WORKLOAD["sigma_u"],
                               use_diffeomorphic=WORKLOAD["diffeomorphic"], verbosity=0, field_mode=mode,
                               device=local)
    ident = [DemonsSlabGroup.unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(ident, src=0)
        grp = DemonsSlabGroup.nccl(shape, spacing, par, ident[0], rank, world, local)
    else:
        grp = DemonsSlabGroup.local(shape, spacing, par, devices=[local])

    def step():
        grp.load_device(fixed.data_ptr(), moving.data_ptr())
        return grp.run(iters)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ms, iters_done, launches, x_ms, x_n, halo_bytes = 0.0, 0, 0, 0.0, 0, 0
    for _ in range(args.steps):
        hist, st = step()
        ms += st.loop_ms
        iters_done += st.iterations_done
        launches += st.kernel_launches
        xs = grp.exchange_stats()
        x_ms += xs["exchange_ms"]
        x_n += xs["exchanges"]
        halo_bytes += grp.halo_bytes
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    t = torch.tensor([ms, float(launches), x_ms, float(halo_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, launches, x_ms_max, halo_bytes_all = float(tm[0]), int(ts[1]), float(tm[2]), float(ts[3])
    else:
        x_ms_max, halo_bytes_all = x_ms, float(halo_bytes)
    value = nvox * iters_done / (ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    final_mse = float(hist[-1]) if len(hist) else None

    # ---- end to end through the C ABI of the sharded path with pinned HOST buffers: every rank uploads its slab of the
    # fixed image and 1/N of the moving image (replicated over NVLink), runs the loop and downloads its slab of the fp64 field
    lib = capi.lib()
    z0, z1 = grp.owned_range()
    plane = shape[1] * shape[2]
    hf = fixed.cpu().pin_memory()
    hm = moving.cpu().pin_memory()
    hD = torch.empty(((z1 - z0) * plane, 3), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    stats = capi.RunStats()

    def e2e_step():
        capi.check(lib.dcma_b200_slab_group_load(grp._h, C.c_void_p(hf.data_ptr()), C.c_void_p(hm.data_ptr()), 0))
        capi.check(lib.dcma_b200_slab_group_run(grp._h, iters, None, C.byref(stats)))
        capi.check(lib.dcma_b200_slab_group_get_deformation(grp._h, C.c_void_p(hD.data_ptr())))
    e2e_step()
    barrier()
    te0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(e2e_steps):
        e2e_step()
        e2e_iters += stats.iterations_done
    barrier()
    e2e_s = time.perf_counter() - te0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    # slab of F (+halo) and this rank's 1/N of M (the ranks replicate M among themselves over NVLink: csrc/slab.cu, load)
    m_slices = shape[0] if world == 1 else -(-shape[0] // world)
    h2d = ((z1 - z0) + (grp.exchange_stats()["halo_planes"] * 2) + m_slices) * plane * 4
    e2e = {"value": nvox * e2e_iters / e2e_s, "unit": "voxel-updates/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int((z1 - z0) * plane * 24), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "bytes_are": "per rank (rank 0)",
           "nvlink_bytes_per_step": int((shape[0] - m_slices) * plane * 4) if world > 1 else 0,
           "api": "dcma_b200_slab_group_load / _run / _get_deformation (C ABI, pinned host buffers; each rank uploads its "
                  "slab of the fixed image and 1/N of the moving image -- the ranks replicate the moving image among "
                  "themselves over NVLink --, downloads its slab of the fp64 AoS field)"}
    xstats = grp.exchange_stats()
    grp.close()
    del hf, hm, hD, fixed, moving
    torch.cuda.empty_cache()

    check = _slab_bit_identity(args, dist, dev, rank, world, local) if world > 1 and not args.no_slab_check else None
    replicas = None
    if world > 1 and not args.no_replicas:
        replicas = _replicas_throughput(args, dist, dev, rank, world, local, torch.cuda.current_stream())

    if rank == 0:
        n_it = max(iters_done, 1)
        line = {"metric": "demons_voxel_updates_per_s", "value": value, "unit": "voxel-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": DTYPE[mode],
                "data": "synthetic", "config": slab_config(args, mode, world, shape),
                "ms_per_iteration": ms / n_it, "final_mse": final_mse, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "wall_ms_timed_region": wall_ms,
                "exchange_ms_per_iter": x_ms_max / n_it, "exchanges_per_iter": x_n / n_it,
                "halo_bytes_per_iter": halo_bytes_all / n_it, "halo_planes": xstats["halo_planes"],
                "exchange_overlapped": xstats["overlapped"], "transport": xstats["transport"],
                "slab_vs_whole_check": check, "replicas": replicas,
                "roofline": {"bound": "hbm", "kernel": "whole iteration (per-kernel split: see the N=1 line)",
                             "achieved": value / world * B_ALG[mode] / 1e9, "peak": peak, "peak_source": peak_src,
                             "unit": "GB/s", "frac": value / world * B_ALG[mode] / 1e9 / peak, "traffic": None,
                             "whole_iteration": {"algorithmic_bytes_per_voxel_update": B_ALG[mode],
                                                 "achieved_GBs_per_gpu": value / world * B_ALG[mode] / 1e9,
                                                 "frac_of_peak": value / world * B_ALG[mode] / 1e9 / peak}},
                "cpu_baseline": None, "nccl": {"version": ".".join(str(v) for v in torch.cuda.nccl.version()),
                                               "debug": os.environ.get("NCCL_DEBUG", "")},
                "parity": "slab result == whole-volume engine bit for bit (slab_vs_whole_check, run here); the whole-volume "
                          "engine vs the reference: tests/test_gpu_full_length.py"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=200, help="Demons iterations per step (BASELINE config 2: 200)")
    ap.add_argument("--field-mode", default=os.environ.get("DCMA_B200_FIELD_MODE", "fp64"),
                    choices=["fp64", "fp64_strict", "mixed", "fp32"],
                    help="field storage/arithmetic of the measured run.  Default fp64: the reference's own storage type and "
                         "arithmetic, 3e-7 mm from the reference engine after 200 iterations at full size.  mixed (D double, "
                         "U float; the same 3e-7 mm) is ~1.15x faster, fp32 (float fields, FFMA2 smoothing) ~1.55x; both are "
                         "reported beside the headline in `other_modes`.  fp32's worst voxel at full size is 1.1e-3 mm off "
                         "(99.99 %% of voxels within 3e-6 mm), i.e. outside the 1e-3 mm tolerance")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the short runs of the other field modes reported beside the headline")
    ap.add_argument("--other-modes", default="mixed,fp32", help="field modes reported beside the headline (50-iteration runs)")
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed steps of the pinned C-ABI end-to-end figure")
    ap.add_argument("--dropin-steps", type=int, default=10, help="timed steps of the AlignViaDemons (drop-in) end-to-end figure")
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in end-to-end figure (e2e = the C-ABI figure)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slab", action="store_true",
                    help="force the z-slab run (the default whenever WORLD_SIZE > 1) also on one GPU")
    ap.add_argument("--replicas", action="store_true",
                    help="N > 1: make N independent config-2 registrations (weak scaling, no collective) the headline "
                         "instead of the sharded 1024^3 registration")
    ap.add_argument("--shape", default="", help="sharded run only: slices,rows,cols (default 1024,1024,1024 = config 4)")
    ap.add_argument("--slab-iters", type=int, default=SLAB_ITERS, help="iterations per step of the sharded run")
    ap.add_argument("--no-slab-check", action="store_true", help="skip the slab-vs-whole bit-identity check")
    ap.add_argument("--no-replicas", action="store_true", help="skip the independent-registrations figure")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if os.environ.get("NCCL_DEBUG") is None and world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's own diagnostics in the log (INFO is very chatty at 8 ranks)
    if args.impl == "reference":
        run_reference(args)
    elif args.slab or (world > 1 and not args.replicas):
        run_slab(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
