#!/usr/bin/env python
"""bench.py — PSF planes/s on the BASELINE.json headline workload.

Workload (BASELINE.json configs[1]): polychromatic obscured-aperture PSF — 1024^2 annular pupil
(bbox 1001^2) with Zernike WFE, 100 wavelengths 500-900 nm, 512^2 detector at oversample 2
(1024^2 samples).  One *plane* = one dft2 (one Field at one wavelength); one *step* = one pass of
the hot path over the 100-wavelength batch: K1 pupil prep (fused into K2a's load) -> K2a matrix
Fourier transform -> K3 |E|^2 accumulate.  With N ranks every rank takes 100 wavelengths of a 100*N
wavelength PSF (weak scaling) and the local PSFs are summed with one NCCL all-reduce per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).
  value           inputs resident in HBM, CUDA-event time, max over ranks
  e2e             the public API with host (pinned) numpy arrays in and a numpy PSF out, copies inside the timed region
  roofline        the dominant kernel (K2a as executed): EXECUTED flops / measured FP64 pipe rate (frac <= 1); the
                  algorithmic flops of SURVEY.md 8(d) and the saving over them are separate keys
  roofline_tensor the north star's FP64 tensor-core (DMMA) execution of the same workload, forced, against the measured
                  DMMA issue rate
  k1 / k3         the element-wise and merge kernels in GB/s against MEASURED_PEAKS.json hbm_gbs
  strong_cfg5     a bounded strong-scaling leg on BASELINE configs[4] (4096^2 pupil -> 2048^2 samples, 16 field points):
                  the SAME total work at every N, wavelengths dealt to the ranks, (16, 2048, 2048) stack all-reduced
  parity_*        this very run against the CPU reference: one wavelength at N = 1, the all-reduced multi-rank PSF at N > 1
  cpu_baseline    the reference (real lentil staged under oracle/_ref, else the oracle port) on the host cores, on a bounded
                  sample of the same workload; cpu_baseline_1thread the same with BLAS pinned to one thread
`--impl reference` runs only that CPU leg, on rank 0, and prints the same kind of line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n=1024, radius=500, obscuration=1 / 3, nzern=15, nlam=100, lam0=500e-9, lam1=900e-9,
                det=512, oversample=2, dx=1 / 1000, z=20.0, du=5e-6)
# bounded strong-scaling leg on BASELINE configs[4]
CFG5 = dict(n=4096, radius=2040, nlam=128, lam0=500e-9, lam1=900e-9, det=1024, oversample=2, dx=1 / 4080, z=20.0,
            du=5e-6, field_points=16, field_halfwidth=20e-6)


def make_inputs(nlam):
    from lentil_b200 import synth
    w = WORKLOAD
    mask = synth.annulus((w["n"], w["n"]), w["radius"], w["obscuration"])
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, np.random.default_rng(0).normal(size=w["nzern"]) * 30e-9)
    wls = np.linspace(w["lam0"], w["lam1"], nlam)
    wts = np.full(nlam, 1.0 / nlam)
    return amp, opd, wls, wts


def plane_flops(m, n, M, N):
    """8*M*n*(m+N) real flops per plane (complex MAC = 8)."""
    return 8.0 * M * n * (m + N)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 8:
                continue
            try:
                parsed.append((self._stamp(parts[0]), float(parts[1]), float(parts[2]),
                               [n for n, v in zip(names, parts[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue

        def summary(rows):
            return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                    "sm_max_mhz": max(r[2] for r in rows) if rows else None, "samples": len(rows),
                    "reasons": sorted({n for r in rows for n in r[3]})}
        # samples whose nvidia-smi timestamp lies inside the timed region (the sampler runs from before the warm-up, so that it
        # is up when the region starts); the samples of the whole loaded phase (warm-up included) are summarised beside them
        inside = [r for r in parsed if r[0] is not None and self.t0 is not None and self.t1 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        out = summary(inside)
        out["timed_region_s"] = None if self.t0 is None or self.t1 is None else self.t1 - self.t0
        out["under_load_incl_warmup"] = summary(parsed)
        if not inside and parsed:            # region shorter than nvidia-smi's sampling period: report the loaded phase
            out.update({k: out["under_load_incl_warmup"][k] for k in ("sm_mhz", "sm_max_mhz", "reasons")})
            out["note"] = "no nvidia-smi sample fell inside the timed region; values from the loaded phase around it"
        return out


# ------------------------------------------------------------------------------------------------
def _reference_module():
    """The real lentil staged by oracle/build_ref.sh (travels to the GPU box in oracle/_ref/), else None."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_loader
        return ref_loader.reference()
    except Exception:
        return None


def cpu_reference_psf(amp, opd, wls, wts):
    """The reference algorithm on the host: the real lentil when a copy is staged under oracle/_ref,
    else the oracle port (numpy + BLAS, bit-identical arithmetic)."""
    w = WORKLOAD
    ref = _reference_module()
    if ref is not None:
        p = ref.Pupil(amplitude=amp, opd=opd, pixelscale=w["dx"], focal_length=w["z"])
        img = np.zeros((w["det"] * w["oversample"],) * 2)
        for wl, wt in zip(wls, wts):
            wf = ref.propagate_dft(ref.Wavefront(wl) * p, pixelscale=w["du"], shape=(w["det"],) * 2,
                                   oversample=w["oversample"])
            img = wf.insert(img, wt)
        return img, "reference"
    import lentil_oracle as oc
    img = oc.psf(amp, opd, None, wls, wts, (w["dx"], w["dx"]), w["z"], w["du"], (w["det"],) * 2, None,
                 w["oversample"])
    return img, "port"


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def _limit_threads(n):
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        return None


def time_cpu(amp, opd, wls, wts, budget_s, max_planes):
    """Planes/s of the CPU reference on a bounded sample: wavelengths taken evenly from the
    workload until `budget_s` seconds or `max_planes` are spent (first plane = warm-up)."""
    pick = np.linspace(0, len(wls) - 1, max_planes).round().astype(int)
    cpu_reference_psf(amp, opd, wls[pick[:1]], wts[pick[:1]])           # BLAS thread spin-up
    done, t0, kind = 0, time.perf_counter(), "port"
    for k in pick:
        _, kind = cpu_reference_psf(amp, opd, wls[k:k + 1], wts[k:k + 1])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _limit_threads(os.cpu_count())      # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host thread
    amp, opd, wls, wts = make_inputs(WORKLOAD["nlam"])
    per_step = 4
    cpu_reference_psf(amp, opd, wls[:1], wts[:1])
    for _ in range(args.warmup):
        cpu_reference_psf(amp, opd, wls[:1], wts[:1])
    pick = np.linspace(0, len(wls) - 1, per_step * max(args.steps, 1)).round().astype(int)
    t0 = time.perf_counter()
    kind = "port"
    for s in range(args.steps):
        idx = pick[s * per_step:(s + 1) * per_step]
        _, kind = cpu_reference_psf(amp, opd, wls[idx], wts[idx])
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    cores = blas_threads()
    sample = (f"{per_step} of the {WORKLOAD['nlam']} wavelengths per step (evenly spaced), full "
              f"Wavefront*Pupil -> propagate_dft -> insert per wavelength")
    print(json.dumps({
        "impl": "reference", "metric": "psf_planes_per_sec", "value": value, "unit": "planes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "planes/s", "cores": cores, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count(),
                         "what": "andykee/lentil 0.8.8 itself (oracle/_ref, staged by oracle/build_ref.sh)" if kind == "reference"
                         else "oracle port of lentil's propagate_dft (oracle/lentil_oracle.py, pinned bit-for-bit)"},
        "e2e": {"value": value, "unit": "planes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(n_gpus):
    w = WORKLOAD
    return {"workload": "BASELINE configs[1]: polychromatic obscured-aperture PSF, 1024^2 annular pupil "
                        "(bbox 1001^2) + 15 Zernike WFE -> 512^2 detector x oversample 2 (1024^2 samples), "
                        f"{w['nlam']} wavelengths 500-900 nm per GPU",
            "plane": "1001x1001 -> 1024x1024 complex128", "planes_per_step_per_gpu": w["nlam"],
            "gflop_per_plane": plane_flops(1001, 1001, 1024, 1024) / 1e9,
            "parallelism": f"wavelength-sharded x{n_gpus}, NCCL all-reduce of the PSF per step" if n_gpus > 1
            else "single GPU",
            "l2": "per-step working set ~2.6 GB (transposed intermediates + per-plane intensities) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
def _private_stdout():
    """stdout carries exactly ONE JSON line, but libraries write there too (NCCL prints its version banner on file
    descriptor 1 when a communicator is created).  Keep a private duplicate of the real stdout for the JSON line and point
    descriptor 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def _event_pair(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_over_ranks(torch, dist, dev, values):
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def strong_cfg5_leg(lentil, torch, dist, dev, world, rank):
    """BASELINE configs[4], bounded: the same 128 wavelengths x 16 field points (2048 planes of 4081^2 -> 2048^2) at every
    N, wavelengths dealt to the ranks, the (16, 2048, 2048) float64 stack all-reduced.  Time = CUDA events around the whole
    call (planning, K1, K2a, K3, all-reduce), max over ranks."""
    from lentil_b200 import synth
    c = CFG5
    mask = synth.annulus((c["n"], c["n"]), c["radius"])
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, np.random.default_rng(5).normal(size=10) * 30e-9)
    pupil = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=c["dx"], focal_length=c["z"])
    pupil.freeze()
    k = int(round(c["field_points"] ** 0.5))
    hw = c["field_halfwidth"]
    tilts = [[rx, ry] for rx in np.linspace(-hw, hw, k) for ry in np.linspace(-hw, hw, k)]
    wl = np.linspace(c["lam0"], c["lam1"], c["nlam"])

    def run(wls):
        return lentil.propagate_dft_batch(pupil, wls, c["du"], (c["det"],) * 2, oversample=c["oversample"],
                                          weights=np.full(len(wls), 1.0 / len(wls)), tilts=tilts,
                                          distributed=world > 1, return_device=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    run(wl[:: max(len(wl) // (2 * world), 1)][: 2 * world])              # warm-up: two wavelengths per rank
    barrier()
    e0, e1 = _event_pair(torch)
    e0.record()
    stack = run(wl)
    e1.record()
    barrier()
    ms, = _max_over_ranks(torch, dist, dev, [e0.elapsed_time(e1)])
    planes = len(wl) * len(tilts)
    from lentil_b200 import _lib
    d = (_lib.MftDesc * 1)()
    d[0].m = d[0].n = 2 * c["radius"] + 1
    d[0].M = d[0].N = c["det"] * c["oversample"]
    out = {"workload": f"BASELINE configs[4], bounded: 4096^2 annular pupil (bbox 4081^2) -> 1024^2 det x os2, {len(wl)} wavelengths x "
                       f"{len(tilts)} field points = {planes} planes, the same total at every N (strong scaling)",
           "n_gpus": world, "planes": planes, "seconds": ms * 1e-3, "planes_per_s": planes / (ms * 1e-3),
           "k2a_execution": {0: "direct", 1: "folded", 2: "czt"}[_lib.lib().lfd_mft_execution(d, 1)],
           "tflops_algorithmic": planes * plane_flops(d[0].m, d[0].n, d[0].M, d[0].N) / (ms * 1e-3) / 1e12,
           "stack_shape": list(stack.shape), "stack_sum": float(stack.sum()), "stack_max": float(stack.max())}
    del stack, pupil
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    out_stream = _private_stdout()
    import torch
    import lentil_b200 as lentil
    from lentil_b200 import device, fourier, _lib
    from lentil_b200 import field as lfield

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = device.set_device(local)
    cpus_before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_cpus = device.bind_cpu_affinity(local) if world > 1 else None     # one process per GPU: stay on the GPU's NUMA node
    lib = _lib.lib()
    w = WORKLOAD
    nlam_total = w["nlam"] * world
    amp, opd, wls, wts = make_inputs(nlam_total)
    shape = (w["det"], w["det"])
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)

    pupil = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=w["dx"], focal_length=w["z"])
    pupil.freeze()                                         # operands resident in HBM

    def step_resident(**kw):
        return lentil.propagate_dft_batch(pupil, wls, w["du"], shape, oversample=w["oversample"], weights=wts,
                                          distributed=world > 1, return_device=True, **kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler is started first: nvidia-smi needs ~0.1 s to deliver its first sample) -----------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        psf = step_resident()
    barrier()

    # ---- timed: resident inputs, device time, with K2a and K3 timed on their own ---------------
    fourier.TIMERS, lfield.TIMERS = [], []
    launches0 = device.launch_count()
    e0, e1 = _event_pair(torch)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(steps):
        psf = step_resident()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = device.launch_count() - launches0
    mft_ms = sum(t[0].elapsed_time(t[1]) for t in fourier.TIMERS)
    mft_flops = sum(t[2] for t in fourier.TIMERS)
    mft_exec = sum(t[3] for t in fourier.TIMERS)
    mft_launches = len(fourier.TIMERS)
    k3_ms = sum(t[0].elapsed_time(t[1]) for t in lfield.TIMERS)
    k3_bytes = sum(t[2] for t in lfield.TIMERS)
    k3_launches = len(lfield.TIMERS)
    fourier.TIMERS = lfield.TIMERS = None
    clocks = sampler.stop() if rank == 0 else None
    ms, mft_ms, k3_ms = _max_over_ranks(torch, dist, dev, [ms, mft_ms, k3_ms])
    planes = w["nlam"] * world * steps
    value = planes / (ms * 1e-3)

    # ---- optional complex64 mode (K2b), both executions: 3xTF32 on tcgen05 and the FP32 chirp-z row transform ---
    c64 = {}
    for name, execution in (("c64_3xtf32", "folded"), ("c64_czt", "czt")):
        for _ in range(3):
            psf32 = step_resident(precision="c64", execution=execution)
        barrier()
        c0, c1 = _event_pair(torch)
        c0.record()
        for _ in range(steps):
            psf32 = step_resident(precision="c64", execution=execution)
        c1.record()
        barrier()
        c64_ms, = _max_over_ranks(torch, dist, dev, [c0.elapsed_time(c1)])
        c64[name] = {"value": planes / (c64_ms * 1e-3), "unit": "planes/s",
                     "peak_normalised_error_vs_fp64": float((psf32 - psf).abs().max() / psf.max())}
    c64["c64_3xtf32"]["note"] = "optional complex64 mode, tcgen05 kind::tf32 with TMEM accumulators (forced: execution='folded'); gate 1e-5"
    c64["c64_czt"]["note"] = "optional complex64 mode, FP32 build of the chirp-z row transform (what precision='c64' runs by default); gate 1e-5"

    # ---- the FP64 tensor-core execution (folded DMMA form) of the same workload, forced, beside the default ------
    probe_desc = (_lib.MftDesc * 1)()
    probe_desc[0].m = probe_desc[0].n = 2 * w["radius"] + 1          # bounding box of the pupil
    probe_desc[0].M = probe_desc[0].N = w["det"] * w["oversample"]
    execution = {0: "direct", 1: "folded", 2: "czt"}[lib.lfd_mft_execution(probe_desc, 1)]
    dmma = None
    if execution != "folded":
        for _ in range(3):
            psf_d = step_resident(execution="folded")          # per-call choice (lfd_mft_desc.execution), no process-wide switch
        barrier()
        fourier.TIMERS = []
        d0, d1 = _event_pair(torch)
        d0.record()
        nd = min(steps, 10)
        for _ in range(nd):
            psf_d = step_resident(execution="folded")
        d1.record()
        barrier()
        d_mft_ms = sum(t_[0].elapsed_time(t_[1]) for t_ in fourier.TIMERS)
        d_exec = sum(t_[3] for t_ in fourier.TIMERS)
        d_alg = sum(t_[2] for t_ in fourier.TIMERS)
        d_launches = len(fourier.TIMERS)
        fourier.TIMERS = None
        d_ms, = _max_over_ranks(torch, dist, dev, [d0.elapsed_time(d1)])
        dmma = {"value": w["nlam"] * world * nd / (d_ms * 1e-3), "unit": "planes/s", "steps": nd,
                "avg_launch_ms": d_mft_ms / max(d_launches, 1), "launches": d_launches,
                "flops_per_launch": d_exec / max(d_launches, 1),
                "achieved": d_exec / (d_mft_ms * 1e-3) / 1e12,
                "algorithmic_tflops": d_alg / (d_mft_ms * 1e-3) / 1e12,
                "max_abs_diff_vs_default_over_peak": float((psf_d - psf).abs().max() / psf.max())}

    # ---- K1 on its own (it is fused into K2a's load on the headline path): phasors of one step's wavelengths ------
    ops = pupil._operands()
    k1_lam = wls[:w["nlam"]]
    phasors = torch.empty(len(k1_lam), ops["total"], dtype=torch.complex128, device=dev)
    for _ in range(2):
        pupil._phasors_into(phasors, k1_lam, ops)
    torch.cuda.synchronize()
    k0, k1e = _event_pair(torch)
    k0.record()
    k1_reps = 5
    for _ in range(k1_reps):
        pupil._phasors_into(phasors, k1_lam, ops)
    k1e.record()
    torch.cuda.synchronize()
    k1_ms = k0.elapsed_time(k1e) / k1_reps
    k1_bytes = 16.0 * ops["total"] * len(k1_lam) + 16.0 * ops["total"]        # phasors written + amp/opd read once
    del phasors

    # ---- e2e: public API, host arrays in (pinned), numpy PSF out, copies inside the timed region ---
    amp_pin = torch.from_numpy(amp).pin_memory()
    opd_pin = torch.from_numpy(opd).pin_memory()
    amp_h, opd_h = amp_pin.numpy(), opd_pin.numpy()

    def step_e2e():
        p = lentil.Pupil(amplitude=amp_h, opd=opd_h, pixelscale=w["dx"], focal_length=w["z"])
        return lentil.propagate_dft_batch(p, wls, w["du"], shape, oversample=w["oversample"], weights=wts,
                                          distributed=world > 1)

    for _ in range(2):
        out = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step_e2e()
    barrier()
    e2e_s, = _max_over_ranks(torch, dist, dev, [time.perf_counter() - t0])
    e2e_value = planes / e2e_s
    h2d = amp.nbytes + opd.nbytes                      # amplitude and opd (the mask is derived from the amplitude on the device)
    d2h = out.nbytes

    # ---- parity of the distributed sum (N > 1): 2 wavelengths per rank, all-reduced, against the CPU reference ----
    parity_distributed = None
    if world > 1:
        sub = np.linspace(0, nlam_total - 1, 2 * world).round().astype(int)
        got = lentil.propagate_dft_batch(pupil, wls[sub], w["du"], shape, oversample=w["oversample"], weights=wts[sub],
                                         distributed=True)
        if rank == 0:
            if numa_cpus and cpus_before:
                os.sched_setaffinity(0, cpus_before)
            _limit_threads(os.cpu_count())
            ref, _ = cpu_reference_psf(amp, opd, wls[sub], wts[sub])
            if numa_cpus:
                os.sched_setaffinity(0, numa_cpus)
            parity_distributed = {"peak_normalised_error": float(np.max(np.abs(got - ref)) / np.max(ref)),
                                  "wavelengths": int(len(sub)), "ranks": world,
                                  "what": "propagate_dft_batch(distributed=True) over all ranks (NCCL all-reduce) vs the CPU "
                                          "reference of the same wavelengths on rank 0"}
        barrier()

    # ---- strong scaling on BASELINE configs[4] (bounded) ---------------------------------------------------
    try:
        strong = strong_cfg5_leg(lentil, torch, dist, dev, world, rank)
    except Exception as exc:                                # the headline line must survive a failure of this leg
        strong = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rooflines -----------------------------------------------------------------------------------------
    probe = device.probe_fp64()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_gbs = peaks.get("hbm_gbs") or 6550.0
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6550 GB/s (B200_PROFILING.md)"
    peak_note = ("measured in this run by lfd_probe_fp64 (register-resident issue loops on all SMs: DMMA.8x8x4 %.1f and DFMA %.1f "
                 "TFLOP/s — one FP64 pipe serves both); MEASURED_PEAKS.json carries HBM and bf16 only; nominal FP64 = "
                 "148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s" % (probe["dmma_tflops"], probe["dfma_tflops"]))
    executed = mft_exec / (mft_ms * 1e-3) / 1e12 if mft_ms > 0 else None
    algorithmic = mft_flops / (mft_ms * 1e-3) / 1e12 if mft_ms > 0 else None
    ncu_traffic = None
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "czt_ncu_summary.json" if execution == "czt"
                                                  else "mft_ncu_summary.json")))
    except Exception:
        pass
    if execution == "czt":
        peak = probe["dfma_tflops"]
        roofline = {
            "bound": "fp64",
            "bound_detail": "FP64 pipe of the SM (DADD/DMUL/DFMA; the same unit executes DMMA), second limiter shared-memory "
                            "bandwidth; not HBM-bound and no tensor-core instruction is issued. ncu: profiles/",
            "kernel": "czt::f64::czt_stage_kernel<11,true>+<11,false> (K2a, chirp-z execution: per row FFT_2048 -> x FFT(chirp) -> "
                      "IFFT_2048 in shared memory, radix-16 passes; one launch = tables + row stage + column stage of the batch)",
            "achieved": executed, "peak": peak, "unit": "TFLOP/s", "frac": executed / peak if executed else None,
            "traffic": None,
            "traffic_note": "not measured in this run; the ncu --set full capture committed under profiles/ is quoted in traffic_ncu",
            "traffic_ncu": ncu_traffic,
            "flops_per_launch": mft_exec / max(mft_launches, 1),
            "flops_model": "EXECUTED flops: per row transform 2 FFTs of length L (5 L log2 L each, the textbook count; the kernel's "
                           "half-length pruning of the first / last butterfly executes ~3 % fewer) + 6 (L + n_in + n_out) for "
                           "the three point-wise products; rows = m + N per plane; L = 2048",
            "avg_launch_ms": mft_ms / max(mft_launches, 1), "launches": mft_launches,
            "share_of_step": mft_ms / ms,
            "algorithmic_tflops": algorithmic,
            "algorithmic_flops_per_launch": mft_flops / max(mft_launches, 1),
            "algorithmic_speedup": mft_flops / mft_exec if mft_exec else None,
            "algorithmic_note": "SURVEY.md 8(d) counts the dense contraction, 8*M*n*(m+N) per plane; the chirp-z execution "
                                "evaluates the same sums with ~33x fewer operations, so algorithmic_tflops exceeds the FP64 "
                                "peak — it is a work saving, not a roofline fraction",
            "peak_source": peak_note,
        }
    else:
        peak = probe["dmma_tflops"]
        roofline = {
            "bound": "tensor", "kernel": "mft_folded_kernel<true>+<false> (K2a, FP64 DMMA.8x8x4; one launch = fold + "
                                         "row stage + column stage of the whole 100-plane batch)",
            "achieved": executed, "peak": peak, "unit": "TFLOP/s", "frac": executed / peak if executed else None,
            "traffic": None, "traffic_ncu": ncu_traffic,
            "flops_per_launch": mft_exec / max(mft_launches, 1),
            "avg_launch_ms": mft_ms / max(mft_launches, 1), "launches": mft_launches, "share_of_step": mft_ms / ms,
            "algorithmic_tflops": algorithmic, "algorithmic_speedup": mft_flops / mft_exec if mft_exec else None,
            "peak_source": peak_note,
        }
    roofline_tensor = None
    if dmma is not None:
        roofline_tensor = dict(dmma)
        roofline_tensor.update({
            "bound": "tensor", "peak": probe["dmma_tflops"], "unit": "TFLOP/s",
            "frac": dmma["achieved"] / probe["dmma_tflops"],
            "kernel": "mft_folded_kernel<true>+<false> (FP64 DMMA.8x8x4, forced per call with lfd_mft_desc.execution = 1 + LFD_MFT_FOLDED)",
            "note": "the north star's FP64 tensor-core execution on the same workload: achieved = EXECUTED DMMA flops "
                    "(two real x complex GEMMs of ceil(M/2) x ceil(K/2) per stage; an upper bound, the row stage skips K tiles "
                    "its support map marks empty) / launch time; frac = % of the measured FP64 tensor-core (DMMA) issue peak. "
                    "The default (LFD_MFT_AUTO) runs the chirp-z execution for this shape because it is faster at the same accuracy",
            "peak_source": peak_note})
    k3 = {"kernel": "accum_kernel<true> (K3: weighted sum of the per-wavelength |F|^2 planes into the PSF)",
          "bound": "hbm", "achieved": k3_bytes / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else None, "peak": hbm_gbs, "unit": "GB/s",
          "frac": (k3_bytes / (k3_ms * 1e-3) / 1e9) / hbm_gbs if k3_ms > 0 else None, "peak_source": hbm_src,
          "bytes_per_launch": k3_bytes / max(k3_launches, 1), "avg_launch_ms": k3_ms / max(k3_launches, 1),
          "launches": k3_launches, "share_of_step": k3_ms / ms,
          "bytes_model": "8 B (intensity windows) or 16 B (complex windows) per covered (pixel, window) + 16 B per output pixel (RMW)"}
    k1 = {"kernel": "pupil_prep_kernel (K1: amp * mask * exp(2 pi i opd / lambda) for one step's wavelengths; timed on its own — "
                    "on the headline path it is fused into K2a's first load and these bytes never exist)",
          "bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
          "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9) / hbm_gbs, "peak_source": hbm_src,
          "bytes_per_launch": k1_bytes, "avg_launch_ms": k1_ms,
          "note": "one FP64 sincospi per element: the FP64 pipe, not HBM, limits this kernel"}

    # ---- CPU baseline (bounded sample of the same workload) -------------------------------------------
    if numa_cpus and cpus_before:
        os.sched_setaffinity(0, cpus_before)               # the CPU legs may use every host core again
    _limit_threads(os.cpu_count())
    cpu_value, cpu_planes, cpu_s, kind = time_cpu(amp, opd, wls[:w["nlam"]], wts[:w["nlam"]],
                                                  budget_s=12.0 if world == 1 else 3.0, max_planes=40)
    cores = blas_threads()
    cpu1 = None
    if world == 1:
        _limit_threads(1)
        v1, p1, s1, _ = time_cpu(amp, opd, wls[:w["nlam"]], wts[:w["nlam"]], budget_s=6.0, max_planes=6)
        cpu1 = {"value": v1, "unit": "planes/s", "cores": 1, "kind": kind,
                "sample": f"{p1} wavelengths in {s1:.1f} s with the BLAS pool limited to one thread (BASELINE.md 5.4)"}
        _limit_threads(os.cpu_count())
    # parity spot check of this very run against the CPU reference (one wavelength)
    ref1, _ = cpu_reference_psf(amp, opd, wls[:1], wts[:1])
    got1 = lentil.propagate_dft_batch(pupil, wls[:1], w["du"], shape, oversample=w["oversample"], weights=wts[:1])
    parity = float(np.max(np.abs(got1 - ref1)) / np.max(ref1))

    line = {
        "metric": "psf_planes_per_sec", "value": value, "unit": "planes/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world),
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "planes/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "api": "lentil_b200.propagate_dft_batch(Pupil(numpy...)) -> numpy",
                "cpu_affinity": (f"{len(numa_cpus)} cores local to the GPU (NVML)" if numa_cpus else "unchanged")},
        "roofline": roofline,
        "roofline_tensor": roofline_tensor,
        "k1": k1, "k3": k3,
        "cpu_baseline": {"value": cpu_value, "unit": "planes/s", "cores": cores, "kind": kind,
                         "host_cpus": os.cpu_count(),
                         "sample": f"{cpu_planes} of the {w['nlam']} wavelengths (evenly spaced) in {cpu_s:.1f} s, "
                                   "full Wavefront*Pupil -> propagate_dft -> insert per wavelength"},
        "cpu_baseline_1thread": cpu1,
        "parity_peak_normalised_error": parity,
        "parity_distributed": parity_distributed,
        "k2a_execution": execution,
        "strong_cfg5": strong,
        "c64_3xtf32": c64["c64_3xtf32"], "c64_czt": c64["c64_czt"],
    }
    print(json.dumps(line), file=out_stream, flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
