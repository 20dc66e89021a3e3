#!/usr/bin/env python
"""bench.py — PSF planes/s on the BASELINE.json headline workload.

Workload (BASELINE.json configs[1]): polychromatic obscured-aperture PSF — 1024^2 annular pupil
(bbox 1001^2) with Zernike WFE, 100 wavelengths 500-900 nm, 512^2 detector at oversample 2
(1024^2 samples).  One *plane* = one dft2 (one Field at one wavelength); one *step* = one pass of
the hot path over the 100-wavelength batch: K1 pupil prep -> K2a matrix Fourier transform
(FP64 DMMA) -> K3 |E|^2 accumulate.  With N ranks every rank takes 100 wavelengths of a 100*N
wavelength PSF (weak scaling) and the local PSFs are summed with one NCCL all-reduce per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event time, max over ranks.
`e2e`: the public API with host (pinned) numpy arrays in and a numpy PSF out, copies inside the
timed region.  `roofline`: the MFT kernel against the FP64 tensor (DMMA) issue rate.
`cpu_baseline` / `--impl reference`: the reference algorithm (oracle port, numpy + BLAS) on the
host cores, on a bounded sample of the same workload.
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
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_psf(amp, opd, wls, wts):
    """The reference algorithm on the host: the real lentil if an install is present under
    baseline/_ref (driver-provided), else the oracle port (numpy + BLAS, same arithmetic)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    w = WORKLOAD
    if os.path.isdir(os.path.join(ref_dir, "lentil")):
        sys.path.insert(0, ref_dir)
        import lentil
        p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=w["dx"], focal_length=w["z"])
        img = np.zeros((w["det"] * w["oversample"],) * 2)
        for wl, wt in zip(wls, wts):
            wf = lentil.propagate_dft(lentil.Wavefront(wl) * p, pixelscale=w["du"], shape=(w["det"],) * 2,
                                      oversample=w["oversample"])
            img = wf.insert(img, wt)
        return img, "reference"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
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
    try:   # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host thread
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
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
                         "host_cpus": os.cpu_count()},
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
            "l2": "per-step working set ~4.9 GB (phasors + intermediates + fields) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
def _private_stdout():
    """stdout carries exactly ONE JSON line, but libraries write there too (NCCL prints its version banner on file
    descriptor 1 when a communicator is created).  Keep a private duplicate of the real stdout for the JSON line and point
    descriptor 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def run_ours(args):
    out_stream = _private_stdout()
    import torch
    import lentil_b200 as lentil
    from lentil_b200 import device, fourier

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = device.set_device(local)
    w = WORKLOAD
    nlam_total = w["nlam"] * world
    amp, opd, wls, wts = make_inputs(nlam_total)
    shape = (w["det"], w["det"])

    pupil = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=w["dx"], focal_length=w["z"])
    pupil.freeze()                                         # operands resident in HBM

    def step_resident():
        return lentil.propagate_dft_batch(pupil, wls, w["du"], shape, oversample=w["oversample"], weights=wts,
                                          distributed=world > 1, return_device=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        psf = step_resident()
    barrier()

    # ---- timed: resident inputs, device time, with the MFT kernel timed on its own ------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fourier.TIMERS = []
    launches0 = device.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        psf = step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = device.launch_count() - launches0
    mft_ms = sum(t[0].elapsed_time(t[1]) for t in fourier.TIMERS)
    mft_flops = sum(t[2] for t in fourier.TIMERS)
    mft_exec = sum(t[3] for t in fourier.TIMERS)
    mft_launches = len(fourier.TIMERS)
    fourier.TIMERS = None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, mft_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, mft_ms = float(t[0]), float(t[1])
    planes = w["nlam"] * world * args.steps
    value = planes / (ms * 1e-3)

    # ---- optional complex64 / 3xTF32 mode (K2b, tcgen05): same workload, reported beside the FP64 headline ---
    def step_c64():
        return lentil.propagate_dft_batch(pupil, wls, w["du"], shape, oversample=w["oversample"], weights=wts,
                                          distributed=world > 1, return_device=True, precision="c64")
    for _ in range(3):
        psf32 = step_c64()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(args.steps):
        psf32 = step_c64()
    c1.record()
    barrier()
    t = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c64_value = planes / (float(t[0]) * 1e-3)
    c64_err = float((psf32 - psf).abs().max() / psf.max())

    # ---- the FP64 tensor-core execution (folded DMMA form) of the same workload, forced, beside the default ------
    from lentil_b200 import _lib
    lib = _lib.lib()
    probe_desc = (_lib.MftDesc * 1)()
    probe_desc[0].m = probe_desc[0].n = 2 * w["radius"] + 1          # bounding box of the pupil
    probe_desc[0].M = probe_desc[0].N = w["det"] * w["oversample"]
    execution = {0: "direct", 1: "folded", 2: "czt"}[lib.lfd_mft_execution(probe_desc, 1)]
    configured = lib.lfd_get_mft_variant()
    dmma = None
    if execution != "folded":
        lib.lfd_set_mft_variant(1)
        for _ in range(3):
            psf_d = step_resident()
        barrier()
        fourier.TIMERS = []
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        nd = min(args.steps, 10)
        for _ in range(nd):
            psf_d = step_resident()
        d1.record()
        barrier()
        d_mft_ms = sum(t_[0].elapsed_time(t_[1]) for t_ in fourier.TIMERS)
        d_exec = sum(t_[3] for t_ in fourier.TIMERS)
        d_alg = sum(t_[2] for t_ in fourier.TIMERS)
        fourier.TIMERS = None
        lib.lfd_set_mft_variant(configured)
        t = torch.tensor([d0.elapsed_time(d1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dmma = {"value": w["nlam"] * world * nd / (float(t[0]) * 1e-3), "unit": "planes/s", "steps": nd,
                "mft_ms_per_launch": d_mft_ms / nd, "executed_tflops": d_exec / (d_mft_ms * 1e-3) / 1e12,
                "algorithmic_tflops": d_alg / (d_mft_ms * 1e-3) / 1e12,
                "max_abs_diff_vs_default_over_peak": float((psf_d - psf).abs().max() / psf.max())}

    # ---- e2e: public API, host arrays in (pinned), numpy PSF out, copies inside the timed region ---
    amp_pin = torch.from_numpy(amp).pin_memory()
    opd_pin = torch.from_numpy(opd).pin_memory()
    amp_h, opd_h = amp_pin.numpy(), opd_pin.numpy()

    def step_e2e():
        p = lentil.Pupil(amplitude=amp_h, opd=opd_h, pixelscale=w["dx"], focal_length=w["z"])
        return lentil.propagate_dft_batch(p, wls, w["du"], shape, oversample=w["oversample"], weights=wts,
                                          distributed=world > 1)

    out = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = planes / float(t[0])
    h2d = amp.nbytes + opd.nbytes + amp.size          # amplitude, opd, uint8 mask
    d2h = out.nbytes

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    probe = device.probe_fp64()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    achieved = mft_flops / (mft_ms * 1e-3) / 1e12 if mft_ms > 0 else None
    executed = mft_exec / (mft_ms * 1e-3) / 1e12 if mft_ms > 0 else None
    traffic = None
    peak_note = ("measured in this run by lfd_probe_fp64 (register-resident issue loops on all SMs: DMMA.8x8x4 %.1f and DFMA %.1f "
                 "TFLOP/s — one FP64 pipe serves both); MEASURED_PEAKS.json carries HBM and bf16 only (hbm_gbs=%s, "
                 "bf16_tflops=%s); nominal FP64 = 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s"
                 % (probe["dmma_tflops"], probe["dfma_tflops"], peaks.get("hbm_gbs"), peaks.get("bf16_tflops")))
    if execution == "czt":
        try:        # ncu --set full capture (profiles/): DRAM bytes per plane, both stages, scaled to one timed launch
            traffic = json.load(open(os.path.join(ROOT, "profiles", "czt_ncu_summary.json")))["dram_bytes_per_plane"] * w["nlam"]
        except Exception:
            pass
        peak = probe["dfma_tflops"]
        roofline = {
            "bound": "tensor",
            "bound_detail": "the FP64 pipe of the SM — the unit that executes both DMMA (the FP64 'tensor core' path) and "
                            "DFMA/DADD/DMUL, which is what this kernel issues; its second limiter is shared-memory "
                            "bandwidth (ncu: l1tex throughput ~80 %). Not HBM-bound (DRAM ~6 % busy)",
            "kernel": "czt_stage_kernel<11,true>+<11,false> (K2a, chirp-z execution: per row FFT_2048 -> x FFT(chirp) -> "
                      "IFFT_2048 in shared memory; one launch = tables + row stage + column stage of the whole batch)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if achieved else None,
            "traffic": traffic,
            "note": "achieved counts ALGORITHMIC flops, 8*M*n*(m+N) per plane (SURVEY.md 8d); the chirp-z execution "
                    "evaluates the same sums with ~30x fewer (two FFTs per row instead of a dense product), so frac > 1 is "
                    "the algorithmic saving, not a timing artefact; executed_* is what the FP64 pipe ran "
                    "(10 L log2 L + 6 (L + n_in + n_out) per row transform)",
            "executed_tflops": executed, "executed_frac": executed / peak if executed else None,
            "peak_source": peak_note,
            "flops_per_launch": mft_flops / max(mft_launches, 1),
            "avg_launch_ms": mft_ms / max(mft_launches, 1), "launches": mft_launches,
            "share_of_step": mft_ms / ms,
        }
    else:
        try:
            # ncu --set full capture (profiles/): DRAM bytes per plane per GEMM stage, scaled to one timed
            # launch here (= both stages of every plane of the step on this rank)
            per_stage = json.load(open(os.path.join(ROOT, "profiles", "mft_ncu_summary.json")))["dram_bytes_per_plane_stage"]
            traffic = per_stage * 2 * w["nlam"]
        except Exception:
            pass
        roofline = {
            "bound": "tensor", "kernel": "mft_folded_kernel<true>+<false> (K2a, FP64 DMMA.8x8x4; one launch = fold + "
                                         "row stage + column stage of the whole 100-plane batch)",
            "achieved": achieved, "peak": probe["dmma_tflops"], "unit": "TFLOP/s",
            "frac": achieved / probe["dmma_tflops"] if achieved else None,
            "traffic": traffic,
            "note": "achieved counts ALGORITHMIC flops, 8*M*n*(m+N) per plane (SURVEY.md 8d); the folded kernel "
                    "executes ~4x fewer (even/odd folding of both DFT axes -> real twiddles), so frac > 1 is the "
                    "algorithmic saving, not a timing artefact; executed_* is what the DMMA pipe ran (an upper bound: the "
                    "row stage also skips the K tiles its support map marks empty, ~21% of them for a disc)",
            "executed_tflops": executed,
            "executed_frac": executed / probe["dmma_tflops"] if executed else None,
            "peak_source": peak_note,
            "flops_per_launch": mft_flops / max(mft_launches, 1),
            "avg_launch_ms": mft_ms / max(mft_launches, 1), "launches": mft_launches,
            "share_of_step": mft_ms / ms,
        }
    if dmma is not None:
        dmma["executed_frac_of_dmma_peak"] = dmma["executed_tflops"] / probe["dmma_tflops"]
        dmma["note"] = ("the FP64 tensor-core execution of the north star (mft_folded_kernel, DMMA.8x8x4), forced with "
                        "lfd_set_mft_variant(LFD_MFT_FOLDED) on the same workload; the default (LFD_MFT_AUTO) runs the "
                        "chirp-z execution for this shape because it is faster at the same accuracy")

    # ---- CPU baseline (bounded sample of the same workload) -------------------------------------------
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    cpu_value, cpu_planes, cpu_s, kind = time_cpu(amp, opd, wls[:w["nlam"]], wts[:w["nlam"]],
                                                  budget_s=12.0 if world == 1 else 3.0, max_planes=40)
    # parity spot check of this very run against the CPU reference (one wavelength)
    ref1, _ = cpu_reference_psf(amp, opd, wls[:1], wts[:1])
    got1 = lentil.propagate_dft_batch(pupil, wls[:1], w["du"], shape, oversample=w["oversample"], weights=wts[:1])
    parity = float(np.max(np.abs(got1 - ref1)) / np.max(ref1))

    line = {
        "metric": "psf_planes_per_sec", "value": value, "unit": "planes/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world),
        "tflops_algorithmic": value * plane_flops(1001, 1001, 1024, 1024) / 1e12,
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "planes/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "api": "lentil_b200.propagate_dft_batch(Pupil(numpy...)) -> numpy"},
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_value, "unit": "planes/s", "cores": blas_threads(), "kind": kind,
                         "host_cpus": os.cpu_count(),
                         "sample": f"{cpu_planes} of the {w['nlam']} wavelengths (evenly spaced) in {cpu_s:.1f} s, "
                                   "full Wavefront*Pupil -> propagate_dft -> insert per wavelength"},
        "parity_peak_normalised_error": parity,
        "k2a_execution": execution, "fp64_dmma_folded": dmma,
        "c64_3xtf32": {"value": c64_value, "unit": "planes/s", "peak_normalised_error_vs_fp64": c64_err,
                       "note": "optional complex64 mode (K2b: tcgen05 kind::tf32, TMEM accumulators); gate 1e-5"},
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
