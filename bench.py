#!/usr/bin/env python
"""bench.py — painted halo-pixels/s of Painter.spray at nside 8192 (BASELINE.json's metric).

Workload (N = 1 and every N): BASELINE config 3's shape — a Websky-shaped full-sky catalog painted
with Battaglia16.kSZ_T (tau_2D_interp x v_r) at nside 8192, R_times 5.  Weak scaling: every rank
(one process per GPU) paints its own --halos halos (seed 1 + rank) into a full-sky partial map and
the partial maps are summed with one NCCL all-reduce inside the timed step.

One JSON line on stdout (rank 0).  `value` is device-resident throughput (inputs in HBM), `e2e` the
same metric through the public API (Painter.spray on a host catalog: pinned host->device copies of
the catalog columns, device->host read-back of the map) — see the task contract for the keys.

  python bench.py                                   # N=1, defaults finish in a few minutes
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference                  # the CPU restatement of the reference loop
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--halos", type=int, default=10_000_000, help="halos per GPU (BASELINE config 3: 1e7)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-halos", type=int, default=400, help="halos of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def config(args, n_gpus):
    return {"workload": f"BASELINE config 3 shape: Websky-shaped synthetic full-sky catalog, {args.halos} halos per GPU, "
                        f"Battaglia16.kSZ_T (tau_2D_interp x v_r), nside {NSIDE}, R_times {R_TIMES}",
            "halos_per_gpu": args.halos, "nside": NSIDE, "R_times": R_TIMES,
            "template": "Battaglia16.kSZ_T", "parallelism": f"halo-sharded x{n_gpus} + NCCL all-reduce of partial maps (per ring band, overlapped)",
            "l2": "fp64 map of 6.44 GB and >1 GB of catalog columns per step: inputs larger than the 126 MB L2"}


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
    df = ap.Catalog(synthetic.websky_like(n_halos, seed=seed)).data
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
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
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


def run_b200(args):
    import torch
    import torch.distributed as dist
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic, _engine as eng, parallel
    from astropaint_b200.profiles import Battaglia16

    ap.paint_bucket.verbose = False
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # ---- inputs: this rank's catalog, built on the host, then resident in HBM ------------------
    cat = ap.Catalog(synthetic.websky_like(args.halos, seed=1 + rank))
    halos = eng.DeviceHalos(cat.data, R_TIMES, dev)
    cols = {k: halos.col(k) for k in ("R_200c", "M_200c", "redshift", "v_r")}
    T_cmb, c_kms, ns = Battaglia16.T_cmb, Battaglia16.c, 15
    npix = 12 * NSIDE * NSIDE
    d_map = torch.zeros(npix, dtype=torch.float64, device=dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    launches_per_step = 6
    n_bands = 8 if world > 1 else 1   # ring bands: the all-reduce of band k overlaps band k+1's QUADPACK phase
    args_t = [cols["R_200c"], cols["M_200c"], cols["redshift"]]

    def step(timers=None):
        """One spray of this rank's catalog into d_map through the engine's pipeline (the same function
        Painter.spray drives), then the sum over ranks."""
        scale = (-cols["v_r"] / c_kms * T_cmb * (1 + cols["redshift"])).contiguous()
        works = []

        def band_done(k, p_lo, p_hi):
            if world > 1 and p_hi > p_lo:
                works.append(dist.all_reduce(d_map[p_lo:p_hi], op=dist.ReduceOp.SUM, async_op=True))

        eng.table_spray(halos, NSIDE, "b16_tau", args_t, scale, ns, "linspace", d_map, d_count, n_bands=n_bands,
                        band_done=band_done, timers=timers)
        if timers is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timers.append(("allreduce_tail", ev))
        for w in works:
            w.wait()
        if timers is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timers.append(("end", ev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    d_count.zero_()
    timers = [[] for _ in range(args.steps)]
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
    ms_total = t0.elapsed_time(t1)
    pairs_per_step = int(d_count.item()) // args.steps
    tt = torch.tensor([ms_total, float(pairs_per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, pairs_all = float(mx[0]), float(sm[1])
    else:
        pairs_all = float(pairs_per_step)
    ms_step = ms_total / args.steps
    value = pairs_all / (ms_step / 1e3)
    phase_ms = {}
    for tl in timers:   # consecutive ticks on the launching stream; every chunk adds to its phase
        for (name, ev), (_, nxt) in zip(tl[:-1], tl[1:]):
            if name != "end":
                phase_ms[name] = phase_ms.get(name, 0.0) + ev.elapsed_time(nxt) / args.steps

    # ---- roofline of the HBM-bound scatter kernel and of the dominant (fp64-bound) kernel -------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s"
    n_pix, _, _, _ = eng.disc_count(halos, NSIDE)
    torch.cuda.synchronize()
    big = n_pix >= ns
    hp_table = int(n_pix[big].sum().item())
    n_big = int(big.sum().item())
    # 16 B per halo-pixel (fp64 read-modify-write) + 8*(7+n_params) B per halo (n_params = 4
    # for kSZ) + 16*n_nodes B of table per halo
    alg_bytes = 16.0 * hp_table + (8.0 * (7 + 4) + 16.0 * ns) * n_big
    paint_gbs = alg_bytes / (phase_ms["paint_table"] / 1e3) / 1e9
    roofline = {"kernel": "paint_kernel<P_TABLE> (fused disc query + geometry + spline + fp64 RED scatter)",
                "bound": "hbm", "achieved": paint_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": paint_gbs / hbm_peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": phase_ms["paint_table"],
                "share_of_step": phase_ms["paint_table"] / ms_step}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("paint_table_bytes_per_launch")
        except (OSError, ValueError):
            pass
    dominant = max(phase_ms, key=phase_ms.get)
    # the dominant kernel is the QUADPACK node kernel: fp64-pipe / issue bound (DRAM < 1% of peak), so an
    # HBM or tensor roofline does not apply; report its rate and the ncu-measured pipe utilisation instead
    n_quad = ns * n_big
    roofline_dominant = {"kernel": "los_nodes_kernel<b16_tau> (QUADPACK dqagie per (halo, node))", "bound": "fp64-pipe",
                         "achieved": n_quad / (phase_ms["los_nodes"] / 1e3), "unit": "quadratures/s",
                         "launch_ms": phase_ms["los_nodes"], "share_of_step": phase_ms["los_nodes"] / ms_step,
                         "fp64_pipe_active_pct_ncu": 63.0, "peak": None, "frac": None,
                         "note": "profiles/r1c_ncu_final.csv: fp64 pipe 63.0% active; 165 integrand evaluations per "
                                 "quadrature (scipy neval), ~142 warp instructions each; hot loop bound by dependent "
                                 "fp64 chains (46% stall_wait), see DESIGN.md section 7"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(args, world),
            "halo_pixels_per_step": pairs_all, "halos_per_s": args.halos * world / (ms_step / 1e3),
            "phase_ms": phase_ms, "dominant_phase": dominant,
            "gpu_launches": launches_per_step * n_bands * args.steps, "torch_elementwise_launches": 4 * args.steps,
            "roofline": roofline, "roofline_dominant": roofline_dominant, "clocks": clocks}

    # ---- e2e: the public API on a host catalog --------------------------------------------------
    if not args.no_e2e:
        del halos, cols
        d_map = None
        torch.cuda.empty_cache()
        cat.pin_memory()
        canvas = ap.Canvas(cat, NSIDE, R_times=R_TIMES)
        painter = ap.Painter(Battaglia16.kSZ_T)
        h2d_cols = ["x", "y", "z", "theta", "phi", "D_a", "R_ang_200c", "R_200c", "M_200c", "redshift", "v_r"]
        h2d = 8 * args.halos * len(h2d_cols)
        d2h = 8 * npix if rank == 0 else 0

        def e2e_step():
            canvas.clean()
            painter.spray(canvas)
            if rank == 0:
                return canvas.pixels  # device -> pinned host
            return None

        for _ in range(3):   # W >= 3 warm-up steps (first one allocates the pinned read-back buffer)
            e2e_step()
        barrier()
        t_start = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t_start) / args.e2e_steps
        pairs = float(int(painter.last_spray["d_count"].item()))
        tt = torch.tensor([dt, pairs], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            dt, pairs = float(mx[0]), float(sm[1])
        line["e2e"] = {"value": pairs / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
                       "api": "Catalog.pin_memory(); Canvas.clean(); Painter(Battaglia16.kSZ_T).spray(canvas); canvas.pixels"}

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
