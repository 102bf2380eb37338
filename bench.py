#!/usr/bin/env python
"""Benchmark of the B200 discrete-ordinates radiance solve (contract: see the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic input: the BASELINE.json configs[1]/[4]
shape (pseudo-spherical DO, 16 streams, 100 layers, Rayleigh + aerosol + absorbers, 10 ground-viewing LOS)
with `--nwavel` wavelengths PER GPU (weak scaling: every rank solves its own contiguous wavelength block,
no data-path collective; only the timing reduction uses torch.distributed).

  value  LOS x wavelength radiances / s, inputs resident in HBM, kernels only (CUDA events inside the library)
  e2e    the same metric through Engine.calculate_radiance(atmosphere) — the reference-facing call — with
         pinned HOST buffers: H2D of the inputs and D2H of the results inside the timed region
  --impl reference   the CPU implementation of the path (oracle port, OpenMP over wavelengths, all host
         cores) on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "LOS x wavelength radiances/sec (16-stream, 100 layers)"
UNIT = "radiances/s"


def flop_model(nstr, nlayers, nlos, m_list, ngroups=1, adjoint_refactor=False):
    """Algorithmic flops per wavelength, split by timed kernel group.

    layer, bvp: SURVEY.md section 8d / Appendix B (values only; the BVP figure is LAPACK's banded LU + solve count,
    the staircase elimination does about a third of it).  Weighting functions (reverse mode):
      wf_adjoint  one banded solve (LAPACK dgbtrs count) per line of sight with the forward factors, as the
                  reference's backprop does; `adjoint_refactor` adds the second factorisation (of A^T) that the
                  SK_B200_ADJOINT=refactor path and few-LOS shapes perform;
      wf_layer    the layer-local linearisation with NL = ngroups + 4 lanes [eps_g | tau | omega | t | s]: NL times the
                  values-only particular + LOS work (forward-mode count of the reference's layer duals,
                  sktran_do_rte.cpp:903-1332, sktran_do_opticallayer.cpp:94-555) + 7 N^3 per eigen-derivative lane
                  (ngroups + 1 lanes, sktran_do_rte.cpp:198-298) + the adjoint contraction 8 N^2 NL per LOS."""
    N, K, L = nstr // 2, nstr, nlayers
    NL, NH = ngroups + 4, ngroups + 1
    layer = bvp = wf_adjoint = wf_layer = 0.0
    for m in m_list:
        homog = L * (6 * N * N * (K - m) + 29 * N**3 + 4 * N * N)            # S+-, eigen-decomposition, W+-
        part = L * (12 * N * (K - m) + 10 * N * N + 60 * N)                  # Green's function particular solution
        post = nlos * L * (6 * N * (K - m) + 8 * N * N + 120 * N)            # LOS source multipliers
        factor = 4 * N * L * (3 * N - 1) * (6 * N - 2)                       # banded LU (LAPACK count)
        solve = 4 * N * L * (9 * N - 3)                                      # one banded solve
        layer += homog + part + post
        bvp += factor + solve
        wf_adjoint += (factor if adjoint_refactor else 0) + nlos * solve
        wf_layer += NL * (part + post) + L * NH * 7 * N**3 + nlos * L * 8 * N * N * NL
    return {"layer": layer, "bvp": bvp, "wf_adjoint": wf_adjoint, "wf_layer": wf_layer, "total": layer + bvp,
            "total_wf": layer + bvp + wf_adjoint + wf_layer}


# DRAM bytes per wavelength (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, divided by
# the 600 wavelengths of the profiled launch) for the default shape: 16 streams, 100 layers, 10 LOS, weighting
# functions with one scattering group - profiles/ncu_r01_v12_summary.csv.  Other shapes report traffic = null.
NCU_DRAM_BYTES_PER_WAVELENGTH = {"layer": (0.813 + 2.020 + 3.324) * 1e9 / 600, "bvp": (6.151 + 7.813) * 1e9 / 600,
                                 "wf_adjoint": (10.033 + 2.442) * 1e9 / 600, "wf_layer": (3.041 + 0.453) * 1e9 / 600}


def tsolve_bytes_model(nstr, nlayers, nlos, n_orders):
    """Algorithmic HBM bytes per wavelength of the transposed solves (k_bvp_tsolve): per
    (order, pivot) one factor row (FS doubles) and one multiplier row (LS doubles) read once, per (order, LOS) the
    right-hand side read, y written and read back, z written (2 N L doubles each)."""
    N = nstr // 2
    fs = ((4 * N + 2) & ~1) + 2
    ls = (3 * N + 2) & ~1
    n = 2 * N * nlayers
    return 8.0 * n_orders * (n * (fs + ls) + nlos * 4 * n)


def bytes_model(nloc, nleg, nlos, nwf_out=0):
    """Algorithmic HBM bytes per wavelength: inputs 8*nloc*(2+nleg) + outputs 8*nlos*(1 + sum nout)."""
    return 8.0 * nloc * (2 + nleg) + 8.0 * nlos * (1 + nwf_out)


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.maxclk = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxclk = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.maxclk, "reasons": sorted(self.reasons)}


def oracle_inputs(sc):
    return dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
                earth_radius=sc.earth_radius, los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa,
                ext=sc.total_extinction, leg=sc.leg_coeff, albedo=sc.albedo)


def time_oracle(sc, sample, threads, with_wf=False):
    """Times the CPU port (oracle) on `sample` wavelengths of the workload with `threads` OpenMP threads."""
    from oracle import oracle

    pick = np.linspace(0, sc.nwavel - 1, sample).astype(int)
    inp = oracle_inputs(sc)
    inp["ssa"] = np.asfortranarray(sc.ssa[:, pick])
    inp["ext"] = np.asfortranarray(sc.total_extinction[:, pick])
    inp["leg"] = np.asfortranarray(sc.leg_coeff[:, :, pick])
    inp["albedo"] = sc.albedo[pick]
    oracle.lib()
    extra = {}
    if with_wf and "wf_aerosol_extinction" in sc.mappings:
        extra = dict(d_leg=np.asfortranarray(sc.mappings["wf_aerosol_extinction"]["d_legendre"][:, :, pick][..., None]),
                     calc_derivs=True)
    t0 = time.perf_counter()
    oracle.do_radiance(**inp, nthreads=threads, **extra)
    dt = time.perf_counter() - t0
    return sample * sc.nlos / dt, dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path.  The reference itself cannot be built here
    (needs Eigen + Rust/cxx generated headers), so this is the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sasktran2_b200 import scenarios

    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    sc = scenarios.config2(nwavel=max(sample, 64), nlayers=args.layers, nstr=args.nstr, nlos=args.nlos)
    time_oracle(sc, min(sample, cores), cores)  # warm-up (thread pool, page-in)
    vals, dts = [], []
    for _ in range(max(args.steps, 1)):
        v, dt = time_oracle(sc, sample, cores)
        vals.append(v)
        dts.append(dt)
    v = float(np.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.median(dts)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pseudo-spherical DO, {args.nstr} streams, {args.layers} layers, {args.nlos} LOS, "
                               f"CPU port on a {sample}-wavelength sample per step, VALUES ONLY (upper bound for the CPU "
                               f"path with weighting functions: the port has no reverse-mode linearisation)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} wavelengths x {args.nlos} LOS per step, OpenMP over wavelengths"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nwavel", type=int, default=int(os.environ.get("SK_BENCH_NWAVEL", "20000")),
                    help="wavelengths per GPU per step")
    ap.add_argument("--nstr", type=int, default=16)
    ap.add_argument("--layers", type=int, default=100)
    ap.add_argument("--nlos", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=4000,
                    help="wavelengths of the CPU-port sample (values only; ~10 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workspace-gb", type=float, default=48.0,
                    help="device workspace per wavelength chunk (B200: 180 GB HBM3e)")
    ap.add_argument("--cpu-wf-sample", type=int, default=8,
                    help="also time the CPU port WITH weighting functions (forward-mode duals) on this many wavelengths")
    ap.add_argument("--wf", type=int, default=1, help="1: with weighting functions (O3, NO2, aerosol mappings + albedo), 0: radiances only")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import sasktran2_b200 as sk
    from sasktran2_b200 import scenarios

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- synthetic workload: every rank builds its own contiguous wavelength block of the global spectrum
    nw_total = args.nwavel * world
    sc_all_small = None
    from sasktran2_b200.parallel import wavelength_block

    start, count = wavelength_block(nw_total, rank, world)
    with_wf = bool(args.wf)
    # every rank builds only its own block of the global spectrum (same formulae as the full scenario)
    sc = scenarios.config2(nwavel=nw_total, nlayers=args.layers, nstr=args.nstr, nlos=args.nlos, with_wf=with_wf,
                           block=(start, count))
    sk._lib.check(sk._lib.lib().sk_b200_set_device(local_rank), "set_device")
    _, geo, view, eng, atm = sk.engine_for_scenario(sc)
    if with_wf:
        atm.surface.enable_albedo_derivative("wf_albedo")
    eng.set_workspace_gb(args.workspace_gb)
    nloc, nleg, nlos, nw = sc.nloc, sc.leg_coeff.shape[0], sc.nlos, sc.nwavel

    # pin the caller-side buffers for the e2e path (the C ABI takes plain host pointers)
    rad_buf = np.zeros((nw, nlos, 1))
    pinned = []
    cudart = torch.cuda.cudart()
    for arr in (atm.storage.ssa, atm.storage.total_extinction, atm.storage.leg_coeff, atm.storage.solar_irradiance,
                atm.surface.albedo, rad_buf):
        rc = cudart.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
        pinned.append((arr, int(rc) == 0 or "success" in str(rc).lower()))

    h2d_bytes = sum(a.nbytes for a in (atm.storage.ssa, atm.storage.total_extinction, atm.storage.leg_coeff,
                                       atm.storage.solar_irradiance, atm.surface.albedo))
    d2h_bytes = rad_buf.nbytes
    nwf_out = 0
    if with_wf:
        for mp in sc.mappings.values():
            h2d_bytes += sum(v.nbytes for v in mp.values() if isinstance(v, np.ndarray))
            nwf_out += sc.nloc
        nwf_out += 1
        d2h_bytes += 8 * nwf_out * nw * nlos

    # ---- device-resident timing (value)
    eng.stage(atm, radiance_buffer=rad_buf)
    info = eng.info()
    m_list = list(range(info["num_azimuth"]))
    for _ in range(args.warmup):
        eng.solve_staged()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    per_kernel = {}
    launches = 0
    for _ in range(args.steps):
        eng.solve_staged()
        t = eng.timings_ms()
        dev_ms += t["kernels_total"]
        # weighting functions: adjoint BVP, layer derivatives, cross-layer chain, mapping
        for k in ("optics", "layer", "bvp", "radiance", "wf_adjoint", "wf_layer", "wf_chain", "wf_map"):
            per_kernel[k] = per_kernel.get(k, 0.0) + t.get(k, 0.0)
        launches += eng.kernel_launches()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    dev_ms = reduce_max(dev_ms)
    wall_ms = reduce_max(wall_ms)
    units_per_step_all = float(nw_total * nlos)
    ms_per_step = dev_ms / args.steps
    value = units_per_step_all / (ms_per_step * 1e-3)
    launches_all = int(reduce_sum(launches))
    check = eng.fetch()["radiance"]
    assert np.all(np.isfinite(check)) and np.all(check > 0), "non-finite radiance in the bench workload"

    # ---- end to end through the reference-facing call with host buffers (e2e): every step copies the inputs from
    # (page-locked) host memory, solves, and copies radiances + weighting functions back into page-locked host arrays
    # that the engine hands out again on the next call (reuse_output_buffers)
    eng.reuse_output_buffers = True
    for _ in range(min(args.warmup, 2)):
        eng.calculate_radiance(atm, rad_buf)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.calculate_radiance(atm, rad_buf)
    barrier()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = units_per_step_all / (e2e_ms * 1e-3)
    t_last = eng.timings_ms()
    e2e_breakdown = {"h2d_ms": t_last["h2d"], "kernels_ms": t_last["kernels_total"], "d2h_ms": t_last["d2h"]}
    e2e_breakdown["host_ms"] = e2e_ms - sum(e2e_breakdown.values())

    # ---- roofline of the dominant kernel (FP64 pipe; peak measured live by a DFMA micro-benchmark)
    reuse = bool(sk._lib.lib().sk_b200_adjoint_reuses_factors(args.nstr // 2, nlos))
    fm = flop_model(args.nstr, args.layers, nlos, m_list, ngroups=1 if with_wf else 0, adjoint_refactor=not reuse)
    fast = args.nstr in (4, 8, 16) and os.environ.get("SK_B200_GENERIC", "0") != "1"
    kernel_names = {
        "layer": "k_eig_setup + k_eig_jacobi + k_layer_post (+ k_los_atten)" if fast else "k_layer_solve",
        "bvp": "k_bvp_v2 (blocked elimination)" if args.nstr <= 16 else "k_bvp",
        "wf_adjoint": ("k_bvp_tsolve" if reuse else "k_bvp_adjoint_v2") if args.nstr <= 16 else "k_bvp_adjoint",
        "wf_layer": "k_wf_layer_fast" if fast else "k_wf_layer",
    }
    launches_per_chunk = {"layer": 4 if fast else 1, "bvp": 1, "wf_adjoint": 1, "wf_layer": 1}
    chunk = info["chunk_wavelengths"]
    nchunks = int(np.ceil(nw / chunk))
    fp64_peak = None
    try:
        import ctypes as C

        fn = sk._lib.lib().sk_b200_measure_fp64_tflops
        fn.restype = C.c_double
        fp64_peak = float(fn())
    except Exception:
        fp64_peak = None
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    step_kernel_ms = max(sum(per_kernel.values()), 1e-12)

    def kernel_roofline(k):
        ms = per_kernel.get(k, 0.0)
        if ms <= 0.0:
            return None
        n_launch = args.steps * nchunks * launches_per_chunk[k]
        tf = fm[k] * nw * args.steps / (ms * 1e-3) / 1e12   # algorithmic flops of the group / its device time
        return {"kernel": kernel_names[k], "achieved": tf, "frac": (tf / fp64_peak) if fp64_peak else None,
                "share_of_step": ms / step_kernel_ms, "avg_launch_ms": ms / max(n_launch, 1),
                "flops_per_launch": fm[k] * (nw / nchunks) / launches_per_chunk[k]}

    per_k = {k: kernel_roofline(k) for k in ("layer", "bvp", "wf_adjoint", "wf_layer")}
    per_k = {k: v for k, v in per_k.items() if v}
    if reuse and "wf_adjoint" in per_k:
        # the transposed solves stream the factors once: bounded by HBM, not by the FP64 pipe
        by = tsolve_bytes_model(args.nstr, args.layers, nlos, len(m_list))
        gbs = by * nw * args.steps / (per_kernel["wf_adjoint"] * 1e-3) / 1e9
        per_k["wf_adjoint"]["hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": gbs / hbm_peak, "bytes_per_wavelength": by}
    dom = max(per_k, key=lambda k: per_k[k]["share_of_step"])
    total_flops = fm["total_wf"] if with_wf else fm["total"]
    roofline = {
        "bound": "fp64", "kernel": per_k[dom]["kernel"], "achieved": per_k[dom]["achieved"], "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": per_k[dom]["frac"],
        "traffic": (NCU_DRAM_BYTES_PER_WAVELENGTH[dom] * (nw / nchunks) / launches_per_chunk[dom]
                    if (args.nstr, args.layers, nlos, with_wf) == (16, 100, 10, True) else None),
        "traffic_source": "profiles/ncu_r01_v12_summary.csv (dram bytes per wavelength of the profiled launch x wavelengths per launch)",
        "peak_source": "DFMA micro-benchmark run inside this bench (MEASURED_PEAKS.json has no FP64 figure)",
        "share_of_step": per_k[dom]["share_of_step"], "avg_launch_ms": per_k[dom]["avg_launch_ms"],
        "flops_per_launch": per_k[dom]["flops_per_launch"],
        "kernels": per_k,
        "whole_step": {"flops_per_wavelength": total_flops,
                       "achieved_tflops": total_flops * (nw_total / world) / (ms_per_step * 1e-3) / 1e12},
        "hbm_view": {"bound": "hbm", "achieved": bytes_model(nloc, nleg, nlos, nwf_out) * (nw_total / world) / (ms_per_step * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s",
                     "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }
    roofline["hbm_view"]["frac"] = roofline["hbm_view"]["achieved"] / hbm_peak

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1] shape: pseudo-spherical DO, {args.nstr} streams, {args.layers} layers, "
                               f"Rayleigh+aerosol+O3/NO2, {nlos} ground-viewing LOS, {args.nwavel} wavelengths per GPU, "
                               + ("weighting functions w.r.t. O3 VMR, NO2 VMR, aerosol extinction (3 x 101 outputs) and albedo"
                                  if with_wf else "radiances only"),
                   "wavelengths_per_gpu": args.nwavel, "azimuth_orders": len(m_list), "chunk_wavelengths": chunk,
                   "l2": "inputs+workspace per step far exceed the 126 MB L2 (no flush needed)",
                   "wall_ms_per_step": wall_ms / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": e2e_ms, "pinned": all(ok for _, ok in pinned), "breakdown": e2e_breakdown},
        "gpu_launches": launches_all,
        "kernel_ms_per_step": {k: v / args.steps for k, v in per_kernel.items()},
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        time_oracle(sc, min(args.cpu_sample, cores), cores)
        v, dt = time_oracle(sc, args.cpu_sample, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_sample} wavelengths x {nlos} LOS of the same atmosphere, oracle port "
                                          f"with OpenMP over wavelengths, {dt:.1f} s; VALUES ONLY, i.e. an upper bound for the "
                                          f"CPU path with weighting functions: the reference's default (do_backprop = false, "
                                          f"cpp/lib/config/config.cpp:14) linearises in forward mode like the port "
                                          f"(with_wf_forward_mode below), its optional reverse mode costs ~2.5-3x values only"}
        if args.cpu_wf_sample > 0:
            v2, dt2 = time_oracle(sc, args.cpu_wf_sample, cores, with_wf=True)
            line["cpu_baseline"]["with_wf_forward_mode"] = {"value": v2, "sample": f"{args.cpu_wf_sample} wavelengths, {dt2:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    for arr, ok in pinned:
        if ok:
            cudart.cudaHostUnregister(arr.ctypes.data)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
