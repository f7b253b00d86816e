#!/usr/bin/env python
"""bench.py — painted halo-pixels/s of Painter.spray at nside 8192 (BASELINE.json's metric).

Workload (every N): BASELINE config 3 — ONE Websky-shaped full-sky catalog of --halos (1e7) halos painted with
Battaglia16.kSZ_T (tau_2D_interp x v_r) at nside 8192, R_times 5.  Strong scaling: with N ranks (one process
per GPU) each rank owns one ring band of the map (an equal share of the sphere's area) and paints the halos
whose disc touches it; there is no partial map and no reduction (astropaint_b200/_bands.py, parallel.py).

One JSON line on stdout (rank 0).  `value` is device-resident throughput (the rank's halo slice already in HBM,
map band zeroed and painted inside the timed step); `e2e` is the same metric through the public API on a
host catalog: Canvas.clean(); Painter.spray(canvas); canvas.pixels -- pinned host->device copies of each
rank's catalog slice and device->host copies of every rank's band into the node's shared host map inside the
timed region.

  python bench.py                                   # N=1, defaults finish in a few minutes
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference                  # the CPU restatement of the reference loop
  python bench.py --config C4 --steps 2 --warmup 1  # one of the other BASELINE configs alone (for ncu)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSIDE = 8192
R_TIMES = 5
METRIC = "painted halo-pixels/s for Painter.spray at nside 8192"
UNIT = "halo-pixels/s"
#: fp64 instructions per integrand evaluation in the hot loop of los_nodes_kernel<b16_tau> (SASS of qk15i_rm's
#: four-evaluation trip, profiles/README.md): the ALGORITHMIC fp64 work of one QUADPACK function evaluation
FP64_INSTR_PER_EVAL = 56


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--halos", type=int, default=10_000_000, help="halos of the WHOLE catalog (BASELINE config 3: 1e7)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE config 3: ONE catalog of --halos halos over all ranks); weak: the catalog "
                         "grows with the ranks (--halos per rank; every rank still owns one band of the one map)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-halos", type=int, default=400, help="halos of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short C1/C2/C4/C5 runs after the headline")
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"],
                    help="C3 = the headline workload; another value times only that BASELINE config (N=1)")
    return ap.parse_args()


def total_halos(args, n_gpus):
    return args.halos * (n_gpus if args.scaling == "weak" else 1)


def config(args, n_gpus):
    return {"workload": f"BASELINE config 3: Websky-shaped synthetic full-sky catalog, {total_halos(args, n_gpus)} halos in total, "
                        f"Battaglia16.kSZ_T (tau_2D_interp x v_r), nside {NSIDE}, R_times {R_TIMES}",
            "halos_total": total_halos(args, n_gpus), "nside": NSIDE, "R_times": R_TIMES, "template": "Battaglia16.kSZ_T",
            "parallelism": f"ring-band ownership x{n_gpus}: each rank paints the halos touching its band of the map, "
                           "no partial maps, no reduction; bands copied to one shared host map over each GPU's own PCIe link",
            "catalog_staging": "Catalog.pin_memory(): columns page-locked in colatitude order, once per catalog, "
                               "outside the timed region",
            "l2": "fp64 map of 6.44 GB and >1 GB of catalog columns and tables per step: inputs larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference loop on a bounded sample
# ---------------------------------------------------------------------------------------------
def _cpu_worker(job):
    df, lo, hi = job
    from oracle import reference_loop as ref, profiles_np as oprof
    import warnings
    warnings.simplefilter("ignore")
    discs = ref.Discs(df, NSIDE, R_TIMES)
    out = []
    for h in range(lo, hi):
        pix = discs.pixel_index(h)
        R = discs.cent2pix_mpc(h, pix)
        val = oprof.b16_kSZ_T(R, df["R_200c"].values[h], df["M_200c"].values[h], df["v_r"].values[h],
                              df["redshift"].values[h])
        out.append((pix, np.asarray(val, dtype=np.float64) * np.ones(len(pix))))
    return out


def cpu_sample(n_halos, n_procs, seed=1):
    """Reference loop (oracle port) on the first n_halos halos of the workload.  Returns
    (halo-pixels, seconds).  n_procs > 1 splits the halos over processes (np.array_split, like the
    reference's n_cpus batches) and scatter-adds the returned (pixel, value) lists in the parent."""
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    ap.paint_bucket.verbose = False
    df = ap.Catalog(synthetic.websky_like(n_halos, seed=seed), device=False).data
    pixels = np.zeros(12 * NSIDE * NSIDE)
    bounds = np.linspace(0, n_halos, n_procs + 1).astype(int)
    jobs = [(df, bounds[i], bounds[i + 1]) for i in range(n_procs)]
    t0 = time.perf_counter()
    if n_procs == 1:
        results = [_cpu_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(n_procs) as pool:
            results = pool.map(_cpu_worker, jobs)
    n_pairs = 0
    for res in results:
        for pix, val in res:
            np.add.at(pixels, pix, val)
            n_pairs += len(pix)
    dt = time.perf_counter() - t0
    return n_pairs, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 64)
    # a bounded sample of the workload per step: ~100 halos per core (about 4 s of QUADPACK-dominated CPU
    # work per step); the reference's cost is a per-halo sum, so the rate does not depend on which halos
    n = max(args.cpu_halos, 100 * cores)
    for _ in range(args.warmup):
        cpu_sample(n, cores)
    times, pairs = [], 0
    steps = max(1, args.steps)
    for _ in range(steps):
        pairs, dt = cpu_sample(n, cores)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = pairs / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"first {n} halos of the workload catalog per step ({pairs} halo-pixels), oracle "
                                       f"restatement of the reference loop split over {cores} processes"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        os.unlink(self.f.name)
        return out


def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except (OSError, ValueError, KeyError):
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def fp64_instr_per_eval_from_sass():
    """fp64 instructions per integrand evaluation, counted in the SASS of the library that is actually loaded: the
    largest loop of los_nodes_kernel<LOS_B16_TAU, FAST> is one trip of the Gauss-Kronrod rule = FOUR evaluations
    (profiles/r2_sass_quadpack_trip.txt).  Returns (count, source); falls back to the committed count when cuobjdump
    is not on the box."""
    import re
    import shutil
    from astropaint_b200 import _capi
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", "-fun", "_Z16los_nodes_kernelILi0ELi1EEvlPKdS1_S1_PK7double4iS1_PKiPd",
                              _capi.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
        ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", out)]
        loops = []
        for a, t in ins:
            if "BRA" in t:
                mm = re.search(r"0x([0-9a-f]+)", t)
                if mm and int(mm.group(1), 16) < a:
                    loops.append((a - int(mm.group(1), 16), int(mm.group(1), 16), a))
        _, lo, hi = max(loops)
        ops = [re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for a, t in ins if lo <= a <= hi]
        n = sum(op in ("DFMA", "DMUL", "DADD") for op in ops)
        if 100 < n < 400:
            return n / 4.0, f"cuobjdump -sass of the loaded library: {n} DFMA/DMUL/DADD in the rule's {len(ops)}-instruction trip of 4 evaluations"
    except (OSError, ValueError, subprocess.SubprocessError):
        pass
    return float(FP64_INSTR_PER_EVAL), "committed count (profiles/r2_sass_quadpack_trip.txt); cuobjdump unavailable"


def measure_fp64_peak(dev, sustained_s=2.0):
    """thread-level fp64 instructions per second this GPU delivers (ap_fp64_peak: 8 DFMA chains per thread, 8 blocks
    of 256 threads per SM, CUDA events): (burst, sustained, sustained_sm_mhz).  burst = best of three 40 ms launches on
    a cool, idle GPU; sustained = the rate over the second half of `sustained_s` seconds of back-to-back launches --
    what a kernel that runs inside a long step can count on once the 1 kW board power cap has pulled the SM clock."""
    import ctypes as C
    import torch
    from astropaint_b200 import _capi
    L = _capi.lib()
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    n = C.c_uint64(0)
    blocks = 148 * 8
    sp = _capi.c_ptr(torch.cuda.current_stream().cuda_stream)
    _capi.check(L.ap_fp64_peak(2000, blocks, _capi.ptr(out), C.byref(n), None))
    torch.cuda.synchronize()
    burst = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _capi.check(L.ap_fp64_peak(40000, blocks, _capi.ptr(out), C.byref(n), sp))
        e1.record()
        e1.synchronize()
        burst = max(burst, n.value / (e0.elapsed_time(e1) / 1e3))
    # sustained: back-to-back launches, rate of the second half
    per_launch_s = n.value / burst
    launches = max(4, int(sustained_s / per_launch_s))
    sampler = ClockSampler(dev.index or 0)
    events = []
    for _ in range(launches):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        events.append(ev)
        _capi.check(L.ap_fp64_peak(40000, blocks, _capi.ptr(out), C.byref(n), sp))
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    end.synchronize()
    clocks = sampler.stop()
    half = launches // 2
    sustained = n.value * (launches - half) / (events[half].elapsed_time(end) / 1e3)
    return burst, sustained, clocks


def mean_neval(dev, halos, cols, nodes, n_nodes, sample=200_000):
    """QUADPACK's own evaluation count (scipy's `neval`) averaged over a sample of this rank's (halo, node)
    quadratures: ap_b16_tau_eval returns it per quadrature."""
    import torch
    from astropaint_b200 import _capi
    live = torch.nonzero(n_nodes > 0).squeeze(1)
    if live.numel() == 0:
        return 0.0
    ns = nodes.shape[1]
    g = torch.Generator(device="cpu").manual_seed(0)
    pick = live[torch.randint(0, live.numel(), (sample,), generator=g).to(dev)]
    node = torch.randint(0, ns, (sample,), generator=g).to(dev)
    R = nodes[pick, node].contiguous()
    a = [cols[k][pick].contiguous() for k in ("R_200c", "M_200c", "redshift")]
    res = torch.empty(sample, dtype=torch.float64, device=dev)
    nev = torch.empty(sample, dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().ap_b16_tau_eval(sample, _capi.ptr(R), _capi.ptr(a[0]), _capi.ptr(a[1]), _capi.ptr(a[2]),
                                            _capi.ptr(res), None, _capi.ptr(nev), None))
    return float(nev.double().mean().item())


def run_b200(args):
    import torch
    import torch.distributed as dist
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic, _engine as eng, _bands
    from astropaint_b200.profiles import Battaglia16

    ap.paint_bucket.verbose = False
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if args.config != "C3":
        assert world == 1, "the other BASELINE configs are single-GPU lines"
        emit({"metric": METRIC, "unit": UNIT, "n_gpus": 1, "configs": run_configs(dev, [args.config], args.steps, args.warmup)})
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks(vals, op):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return [float(x) for x in t]

    MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world > 1 else (None, None)

    # ---- inputs: ONE catalog (same seed on every rank), staged once; this rank's slice resident in HBM ----
    t_stage = time.perf_counter()
    n_total = total_halos(args, world)
    cat = ap.Catalog(synthetic.websky_like(n_total, seed=1))
    cat.pin_memory()
    stage_s = time.perf_counter() - t_stage
    band = _bands.Band(NSIDE, rank, world)
    order = cat.theta_order()[0]
    plan = _bands.Plan(band, cat.disc_reach(R_TIMES), n_chunks=1)
    lo, hi = plan.halo_lo, plan.halo_hi
    halos = eng.DeviceHalos(cat.data, R_TIMES, dev, rows=order[lo:hi], pinned=cat.staged([]), pinned_range=(lo, hi))
    cols = {k: halos.col(k) for k in ("R_200c", "M_200c", "redshift", "v_r")}
    T_cmb, c_kms, ns = Battaglia16.T_cmb, Battaglia16.c, 15
    npix = 12 * NSIDE * NSIDE
    d_band = torch.zeros(band.npix, dtype=torch.float64, device=dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    args_t = [cols["R_200c"], cols["M_200c"], cols["redshift"]]
    launches_per_step = 6     # K1, nodes, QUADPACK, spline, K3, bypass (the zero-fill and the scale are torch kernels)

    def step(timers=None):
        """One spray of this rank's share into its band of the map through the engine's pipeline (the functions
        Painter.spray drives): zero the band, then K1 / nodes / QUADPACK / spline / K3 / bypass."""
        if timers is not None:
            timers.tick("zero_fill")
        d_band.zero_()
        scale = (-cols["v_r"] / c_kms * T_cmb * (1 + cols["redshift"])).contiguous()
        eng.run_chunks(plan, lambda a, e, k: eng.table_chunk(halos.view(a, e), NSIDE, "b16_tau", [t[a:e] for t in args_t],
                                                          scale[a:e], ns, "linspace", d_band, d_count, timers))

    for _ in range(args.warmup):
        step()
    barrier()
    d_count.zero_()
    timers = [eng.Timers() for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        step(timers[s])
    t1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_rank = t0.elapsed_time(t1)
    pairs_rank = int(d_count.item()) // args.steps
    ms_total, = over_ranks([ms_rank], MAX)
    pairs_all, halos_painted = over_ranks([float(pairs_rank), float(halos.n)], SUM)
    ms_step = ms_total / args.steps
    value = pairs_all / (ms_step / 1e3)
    phase_ms = {}
    for tm in timers:
        tm.phase_ms(phase_ms, 1.0 / args.steps)
    # the map itself: a checksum that must agree between N = 1 and any N (each pixel is painted by one rank)
    csum = over_ranks([float(d_band.sum()), float(d_band.abs().sum()), float((d_band != 0).sum())], SUM)
    checksum = {"sum": float(f"{csum[0]:.10e}"), "sum_abs": float(f"{csum[1]:.10e}"), "pixels_painted": int(csum[2])}

    # ---- roofline of the HBM-bound scatter kernel and of the dominant (fp64-bound) kernel -------
    hbm_peak, peak_src = load_peaks()
    n_pix, _, rmin_t, rmax_t = eng.disc_count(halos, NSIDE, want_range=True)   # whole-disc counts of this rank's halos
    nodes, n_nodes = eng.table_nodes(halos, n_pix, rmin_t, rmax_t, ns, "linspace")
    torch.cuda.synchronize()
    n_big = int((n_nodes > 0).sum().item())
    # 16 B per halo-pixel (fp64 read-modify-write) + 8*(7+n_params) B per halo (n_params = 4
    # for kSZ) + 16*n_nodes B of table per halo; the pixels are the ones THIS rank painted (its band)
    alg_bytes = 16.0 * pairs_rank + (8.0 * (7 + 4) + 16.0 * ns) * n_big
    paint_gbs = alg_bytes / (phase_ms["paint_table"] / 1e3) / 1e9
    roofline = {"kernel": "paint_kernel<P_TABLE> (fused disc query + geometry + spline + fp64 RED scatter)",
                "bound": "hbm", "achieved": paint_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": paint_gbs / hbm_peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": phase_ms["paint_table"],
                "share_of_step": phase_ms["paint_table"] / ms_step}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            if world == 1 and tr.get("halos") == n_total:
                roofline["traffic"] = tr.get("paint_table_bytes_per_launch")
        except (OSError, ValueError):
            pass
    dominant = max(phase_ms, key=phase_ms.get)
    # the dominant kernel is the QUADPACK node kernel: fp64-pipe bound (DRAM < 1% of peak), so its roofline is the
    # fp64 instruction rate, measured here with a DFMA probe; its algorithmic work = QUADPACK's own number of
    # integrand evaluations (measured on a sample with ap_b16_tau_eval) x the fp64 instructions of one evaluation
    fp64_burst, fp64_peak, fp64_clocks = measure_fp64_peak(dev)
    nev = mean_neval(dev, halos, cols, nodes, n_nodes)
    n_quad = ns * n_big
    per_eval, per_eval_src = fp64_instr_per_eval_from_sass() if rank == 0 else (float(FP64_INSTR_PER_EVAL), "")
    fp64_instr = n_quad * nev * per_eval
    fp64_rate = fp64_instr / (phase_ms["los_nodes"] / 1e3)
    roofline_dominant = {"kernel": "los_nodes_kernel<b16_tau> (QUADPACK dqagie per (halo, node))", "bound": "fp64-pipe",
                         "achieved": fp64_rate, "peak": fp64_peak, "unit": "fp64 thread-instructions/s",
                         "frac": fp64_rate / fp64_peak, "launch_ms": phase_ms["los_nodes"],
                         "share_of_step": phase_ms["los_nodes"] / ms_step, "quadratures_per_launch": n_quad,
                         "mean_neval": nev, "fp64_instr_per_eval": per_eval, "fp64_instr_per_eval_source": per_eval_src,
                         "peak_source": "ap_fp64_peak DFMA probe timed in this run (8 chains x 8 blocks x 256 threads per SM), "
                                        "SUSTAINED: second half of 2 s of back-to-back launches (this kernel runs inside a long "
                                        "step); the burst figure of a 40 ms launch on an idle GPU is given beside it",
                         "peak_tflops_fp64": 2 * fp64_peak / 1e12, "peak_burst": fp64_burst,
                         "frac_of_burst_peak": fp64_rate / fp64_burst, "sustained_probe_clocks": fp64_clocks}
    del nodes, n_nodes, n_pix, rmin_t, rmax_t

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(args, world),
            "halo_pixels_per_step": pairs_all, "halos_per_s": n_total / (ms_step / 1e3),
            "halos_painted_all_ranks": halos_painted, "phase_ms": phase_ms, "dominant_phase": dominant,
            "gpu_launches": launches_per_step * len(plan.chunks) * args.steps, "torch_elementwise_launches": 5 * args.steps,
            "map_checksum": checksum, "catalog_stage_s": stage_s,
            "roofline": roofline, "roofline_dominant": roofline_dominant, "clocks": clocks}

    # ---- e2e: the public API on a host catalog --------------------------------------------------
    if not args.no_e2e:
        del halos, cols, args_t, d_band
        torch.cuda.empty_cache()
        # every step reads the map back: size the bands by each GPU's device->host bandwidth (equal areas at N = 1)
        canvas = ap.Canvas(cat, NSIDE, R_times=R_TIMES, band_balance="readback")
        painter = ap.Painter(Battaglia16.kSZ_T)
        h2d_cols = ["x", "y", "z", "theta", "phi", "D_a", "R_ang_200c", "R_200c", "M_200c", "redshift", "v_r"]

        def e2e_step():
            canvas.clean()
            painter.spray(canvas)
            return canvas.pixels  # every rank: its band device -> the node's shared pinned host map

        for _ in range(3):   # W >= 3 warm-up steps (the first one allocates the page-locked host map)
            m = e2e_step()
            del m
        barrier()
        t_start = time.perf_counter()
        for _ in range(args.e2e_steps):
            m = e2e_step()
            del m            # nobody keeps the array: the next step may reuse the buffer
        barrier()
        dt = (time.perf_counter() - t_start) / args.e2e_steps
        dt, = over_ranks([dt], MAX)
        pairs, n_h2d = over_ranks([float(int(painter.last_spray["d_count"].item())), float(painter.last_spray["n_halos"])], SUM)
        m = canvas.pixels
        e2e_sum = float(f"{float(m.sum()):.10e}") if rank == 0 else None
        del m
        line["e2e"] = {"value": pairs / dt, "unit": UNIT, "h2d_bytes_per_step": int(8 * n_h2d * len(h2d_cols)),
                       "d2h_bytes_per_step": 8 * npix, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
                       "path": painter.last_spray["path"], "chunks": painter.last_spray["chunks"], "map_sum": e2e_sum,
                       "band_balance": "readback" if world > 1 else "area",
                       "readback_GBps_per_rank": [round(x, 1) for x in canvas._band().weights] if world > 1 else None,
                       "api": "Catalog.pin_memory(); Canvas.clean(); Painter(Battaglia16.kSZ_T).spray(canvas); canvas.pixels "
                              "(all ranks; each copies its band into the node's shared host map)"}
        del canvas, painter
        torch.cuda.empty_cache()

    # ---- the other BASELINE configs, short (N = 1 only) ------------------------------------------
    if world == 1 and not args.no_configs:
        del cat
        line["configs"] = run_configs(dev, ["C1", "C2", "C4", "C5"], 3, 2)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload -----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pairs, dt = cpu_sample(args.cpu_halos, 1)
        line["cpu_baseline"] = {"value": pairs / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"first {args.cpu_halos} halos of the workload catalog ({pairs} halo-pixels, "
                                          f"{dt:.1f} s): oracle restatement of the reference's serial spray loop"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# BASELINE configs 1, 2, 4, 5 (single GPU, short): device time of Painter.spray / stack_cutouts on a staged catalog
# ---------------------------------------------------------------------------------------------
def run_configs(dev, which, steps, warmup):
    import torch
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import spherical, NFW, Battaglia16
    hbm_peak, _ = load_peaks()
    out = {}

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / steps

    def spray_config(cat, nside, tmpl, bytes_per_halo, label):
        """`ms` / `value`: CUDA events around Canvas.clean() + Painter.spray (zero-fill of the map, pinned H2D of the
        catalog columns, every kernel).  `kernel_ms` / `roofline_frac`: the painting kernel(s) alone on resident
        inputs (closed-form templates: ap_paint_profile; tabulated: the K1..bypass pipeline), against the HBM roofline
        with SURVEY 8(d)'s algorithmic bytes."""
        from astropaint_b200 import _engine as eng
        cat.pin_memory()
        canvas = ap.Canvas(cat, nside, R_times=R_TIMES)
        p = ap.Painter(tmpl)

        def one():
            canvas.clean()
            p.spray(canvas)
        ms = timed(one)
        pairs = int(p.last_spray["d_count"].item())
        alg = 16.0 * pairs + bytes_per_halo * cat.size
        out = {"workload": label, "halo_pixels": pairs, "ms": ms, "value": pairs / (ms / 1e3), "unit": UNIT,
               "path": p.last_spray["path"], "algorithmic_bytes": alg}
        # the kernels alone
        halos = eng.DeviceHalos(cat.data, R_TIMES, dev)
        d_map = canvas._device_band()
        info = getattr(tmpl, "_ap_device", None)
        if info is not None:
            cols = [halos.col(a) for a in info["args"]] + [torch.full((1,), float(v), dtype=torch.float64, device=dev)
                                                           for v in info["kwonly"].values()]
            kms = timed(lambda: eng.paint_profile(halos, nside, info["id"], cols, d_map))
            out["kernel"] = f"paint_kernel<{tmpl.__name__}>"
        else:
            ti = tmpl._ap_device_table
            targs = [halos.col(a) for a in ti["args"]]
            tm = eng.Timers()
            kms = timed(lambda: eng.table_chunk(halos, nside, ti["kind"], targs, None, ti["n_samples"], ti["sampling_method"],
                                                d_map, None, tm))
            ph = tm.phase_ms(scale=1.0 / (steps + warmup))
            out["phase_ms"] = ph
            out["kernel"] = "K1 + nodes + QUADPACK + spline + K3 + bypass (fp64-bound: see phase_ms)"
        out["kernel_ms"] = kms
        out["roofline_frac"] = alg / (kms / 1e3) / 1e9 / hbm_peak
        del halos
        return canvas, out

    cat6 = None
    if "C1" in which:
        cat = ap.Catalog(synthetic.random_box(1000, seed=0))
        _, out["C1"] = spray_config(cat, 512, spherical.gaussian_2D, 8.0 * (7 + 4),
                                    "1e3 random-box halos, README Gaussian, nside 512, R_times 5")
    if "C2" in which or "C4" in which or "C5" in which:
        cat6 = ap.Catalog(synthetic.websky_like(1_000_000, seed=1))
    if "C2" in which:
        _, out["C2"] = spray_config(cat6, 4096, Battaglia16.tau_2D_interp, 8.0 * (7 + 3) + 16.0 * 15,
                                    "1e6 Websky-shaped halos, Battaglia16.tau_2D_interp, nside 4096, R_times 5")
    canvas = None
    if "C4" in which or "C5" in which:
        canvas, res = spray_config(cat6, NSIDE, NFW.BG, 8.0 * (7 + 8),
                                   "1e6 Websky-shaped halos, NFW.BG (R_vec), nside 8192, R_times 5")
        if "C4" in which:
            out["C4"] = res
    if "C5" in which:
        halo_list = np.arange(100_000)

        def stack():
            canvas.stack_cutouts(halo_list, lon_range=[-0.2, 0.2], lat_range=[-0.2, 0.2], xpix=200)
        ms = timed(stack)
        samples = 100_000 * 200 * 200
        from astropaint_b200 import _engine as eng
        d_map = canvas.pixels_device
        d_lon = eng.to_device(cat6.data["lon"].values[halo_list], dev)
        d_lat = eng.to_device(cat6.data["lat"].values[halo_list], dev)
        kms = timed(lambda: eng.stack_cutouts(d_map, NSIDE, d_lon, d_lat, [-0.2, 0.2], [-0.2, 0.2], 200, 200))
        out["C5"] = {"workload": "stack_cutouts of 1e5 halos, 0.4 x 0.4 deg, 200 x 200, from the C4 map (nside 8192)",
                     "samples": samples, "ms": ms, "value": 1e5 / (ms / 1e3), "unit": "cutouts/s",
                     "algorithmic_bytes": 8.0 * samples, "kernel": "cutout_kernel<STACK>", "kernel_ms": kms,
                     "roofline_frac": 8.0 * samples / (kms / 1e3) / 1e9 / hbm_peak,
                     "timed": "ms: CUDA events around Canvas.stack_cutouts (H2D of lon/lat, fused project + gather + "
                              "accumulate kernel, D2H of the 200 x 200 stack); kernel_ms: the kernel alone"}
    return out


_JSON_OUT = None


def emit(line):
    """the ONE JSON line, on the process's original stdout"""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse()
    # Libraries write banners to file descriptor 1 (NCCL prints "NCCL version ..." there when /etc/nccl.conf or
    # the environment asks for NCCL_DEBUG=VERSION): keep the original stdout for the JSON line and point fd 1
    # at stderr for everything else.
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
