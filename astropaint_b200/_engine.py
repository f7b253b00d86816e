"""Device-side plumbing of the painting path: catalog columns in HBM and kernel launches.

PyTorch is used for device memory, streams and (in `parallel.py`) torch.distributed; every
compute step is a call into libastropaint_b200.so (csrc/k_*.cu) through `_capi`.
"""
import ctypes as C

import warnings

import numpy as np
import torch

# pandas hands out read-only column views; they are only ever read (copied to pinned memory / HBM)
warnings.filterwarnings("ignore", message="The given NumPy array is not writable")

from . import _capi
from ._capi import ApHalos, c_ptr, c_i32, ptr, check
from .lib import transform  # noqa: F401  (reference expressions cited in DeviceHalos)

GEOMETRY_COLUMNS = ("x", "y", "z", "theta", "phi", "D_a")

#: (halo, pixel) pairs per batch of the flat work-list path (R: 8 B, pixel id: 8 B, value: 8 B each)
FLAT_BATCH_PIXELS = 1 << 25


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("astropaint_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _capi.lib()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def stream_ptr():
    return c_ptr(torch.cuda.current_stream().cuda_stream)


def to_device(arr, device):
    """host float64 column -> device tensor (pinned staging so the copy is a single DMA)."""
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))
    t = torch.from_numpy(a)
    if t.numel() > (1 << 16):
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


class DeviceHalos:
    """Catalog columns of the halos being painted, resident in HBM (SoA, fp64).

    `radius` is R_times * arcmin2rad(R_ang_200c), computed on the host with the reference's own
    numpy expressions (paint_bucket.py:1226-1227, lib/transform.py:143-150) so that the disc
    radius rounds identically.
    """

    def __init__(self, data, R_times, device, rows=None, pinned=None, pinned_range=None, lazy=False):
        """rows: catalog row labels of the halos held, in this object's order (None = every row, catalog order).
        pinned: {column: page-locked host tensor}.  With pinned_range = (lo, hi) the tensors are in the
        catalog's colatitude order (Catalog.pin_memory) and [lo:hi] of them ARE the rows (one contiguous DMA
        per column); otherwise they are in catalog order and `rows` is gathered from them.
        lazy: upload nothing yet -- the caller stages the columns piecewise with upload_pieces()."""
        self.device = device
        self._data = data
        self._rows = rows
        self._pinned = pinned or {}
        self._pinned_range = pinned_range
        self._R_times = float(R_times)
        self.cols = {}
        self._n_all = len(data)
        self.n = self._n_all if rows is None else len(rows)
        if lazy:
            return
        # disc radius = R_times * arcmin2rad(R_ang_200c) = R_times * ((R_ang / 60) * (pi / 180))
        # (paint_bucket.py:1226-1227, lib/transform.py:148-150).  Formed on the device from the uploaded
        # column with the same three IEEE operations numpy performs (ap_disc_radius; torch's own
        # `x / 60` multiplies by a reciprocal and rounds differently), so it rounds identically
        # (tests/test_gpu_parity.py::test_device_radius_is_numpy_radius) and costs no host pass.
        r_ang = self.col("R_ang_200c")
        radius = torch.empty_like(r_ang)
        check(_capi.lib().ap_disc_radius(r_ang.numel(), ptr(r_ang), float(R_times), ptr(radius), stream_ptr()))
        self.cols["__radius__"] = radius
        for k in GEOMETRY_COLUMNS:
            self.col(k)

    def upload_pieces(self, names, pieces, stream):
        """Stage the geometry columns and `names` piece by piece: piece k = halos [a_k, e_k) of this object goes
        host -> device on `stream` (one DMA per column and piece, straight out of the page-locked colatitude-ordered
        staging of Catalog.pin_memory); returns one event per piece.  The painting stream waits for piece k's event,
        calls radius_piece(a_k, e_k) and paints it while the later pieces are still in flight."""
        assert self._pinned_range is not None
        lo = self._pinned_range[0]
        names = list(dict.fromkeys(list(GEOMETRY_COLUMNS) + ["R_ang_200c"] + list(names)))
        L = _capi.lib()
        piecewise = []
        for name in names:
            if name in self.cols:
                continue
            if name in self._pinned:
                self.cols[name] = torch.empty(self.n, dtype=torch.float64, device=self.device)
                piecewise.append(name)
            else:
                self.col(name)                # not staged: whole column now, on the painting stream
        self.cols["__radius__"] = torch.empty(self.n, dtype=torch.float64, device=self.device)
        stream.wait_stream(torch.cuda.current_stream())   # the fresh buffers may be recycled blocks still in use
        sp = c_ptr(stream.cuda_stream)
        events = []
        for (a, e) in pieces:
            for name in piecewise:
                src = self._pinned[name]
                check(L.ap_copy_h2d_async(c_ptr(self.cols[name].data_ptr() + 8 * a), c_ptr(src.data_ptr() + 8 * (lo + a)),
                                          8 * (e - a), sp))
            ev = torch.cuda.Event()
            ev.record(stream)
            events.append(ev)
        return events

    def radius_piece(self, a, e):
        """disc radius of halos [a, e) on the painting stream (their R_ang_200c has arrived)"""
        if e > a:
            check(_capi.lib().ap_disc_radius(e - a, c_ptr(self.cols["R_ang_200c"].data_ptr() + 8 * a), self._R_times,
                                             c_ptr(self.cols["__radius__"].data_ptr() + 8 * a), stream_ptr()))

    @property
    def index(self):
        return np.arange(self._n_all) if self._rows is None else np.asarray(self._rows)

    def _take(self, v):
        """rows of a host column: a view for all rows / a contiguous block, a gather otherwise"""
        if self._pinned_range is not None and isinstance(v, torch.Tensor):
            return v[self._pinned_range[0]:self._pinned_range[1]]
        if self._rows is None:
            return v
        r = self._rows
        if len(r) and r[-1] - r[0] + 1 == len(r) and np.all(np.diff(r) == 1):
            return v[int(r[0]):int(r[-1]) + 1]
        return v[torch.as_tensor(r)] if isinstance(v, torch.Tensor) else v[r]

    def _host_col(self, name):
        if name in self._pinned:
            return self._take(self._pinned[name]).numpy()
        return self._take(self._data[name].values)

    def col(self, name):
        if name not in self.cols:
            if name in self._pinned:
                v = self._take(self._pinned[name])
                v = v if v.is_pinned() else v.pin_memory()
                self.cols[name] = v.to(self.device, non_blocking=True)
            else:
                self.cols[name] = to_device(self._host_col(name), self.device)
        return self.cols[name]

    def struct(self, lo=0, hi=None):
        hi = self.n if hi is None else hi
        s = ApHalos()
        s.n = hi - lo

        def p(name):
            t = self.cols[name]
            return t.data_ptr() + 8 * lo

        s.d_x, s.d_y, s.d_z = p("x"), p("y"), p("z")
        s.d_radius = p("__radius__")
        s.d_theta, s.d_phi, s.d_D_a = p("theta"), p("phi"), p("D_a")
        return s

    def gathered(self, perm):
        """The same halos in the order `perm` (int64 device tensor): one gather per resident column."""
        out = DeviceHalos.__new__(DeviceHalos)
        out.device, out._data, out._pinned, out._rows, out._pinned_range = self.device, self._data, {}, None, None
        out._n_all, out.n = self.n, self.n
        out.cols = {k: v.index_select(0, perm) for k, v in self.cols.items()}
        return out

    def view(self, lo, hi):
        """Halos [lo, hi) of this object without copying (column views)."""
        out = DeviceHalos.__new__(DeviceHalos)
        out.device, out._data, out._pinned, out._rows, out._pinned_range = self.device, self._data, {}, None, None
        out._n_all, out.n = hi - lo, hi - lo
        out.cols = {k: v[lo:hi] for k, v in self.cols.items()}
        return out

    def subset(self, idx):
        """New DeviceHalos holding rows `idx` (positions within this object)."""
        sub = DeviceHalos.__new__(DeviceHalos)
        sub.device = self.device
        sub._data = self._data
        sub._pinned = {}
        sub._pinned_range = None
        sub._rows = self.index[idx]
        sub._n_all = self._n_all
        sub.n = len(idx)
        ti = torch.as_tensor(np.asarray(idx), device=self.device, dtype=torch.int64)
        sub.cols = {k: v.index_select(0, ti) for k, v in self.cols.items()}
        return sub


class band_clip:
    """`with band_clip(band):` -- every painting launch of this thread inside the block paints only the rings
    the rank owns and addresses the map relative to the band's first pixel (ap_band_clip; _bands.Band)."""

    def __init__(self, band):
        self.band = band

    def __enter__(self):
        b = self.band
        if b is not None and not b.whole:
            check(_capi.lib().ap_band_clip(b.ring_lo, b.ring_hi))
        return self

    def __exit__(self, *exc):
        check(_capi.lib().ap_band_clip(0, 0))
        return False


class emit_pairs:
    """`with emit_pairs(n_max, labels, npix, device) as e:` -- deterministic accumulation.  Every painting launch of
    this thread inside the block appends (pixel << shift | halo label, value) pairs instead of adding atomically
    (ap_emit_begin); `e.accumulate(d_map, pix0)` afterwards sorts the keys and adds every pixel's values in label
    order with plain stores (ap_segment_accumulate): the summation order of the reference's serial loop, bit-reproducible
    from run to run."""

    def __init__(self, n_max, labels, npix, device):
        self.n_max = int(n_max)
        pix_bits = max(1, int(npix - 1).bit_length())
        self.shift = min(40, 63 - pix_bits)
        self.labels = labels
        self.keys = torch.empty(max(self.n_max, 1), dtype=torch.int64, device=device)
        self.vals = torch.empty(max(self.n_max, 1), dtype=torch.float64, device=device)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)

    def __enter__(self):
        check(_capi.lib().ap_emit_begin(ptr(self.keys), ptr(self.vals), ptr(self.count), self.n_max, ptr(self.labels),
                                        self.shift))
        return self

    def __exit__(self, *exc):
        check(_capi.lib().ap_emit_end())
        return False

    def accumulate(self, d_map, pix0=0):
        n = int(self.count.item())
        if n > self.n_max:
            raise RuntimeError(f"deterministic accumulation: {n} pairs painted but room for {self.n_max}")
        if n == 0:
            return 0
        keys, perm = torch.sort(self.keys[:n])      # keys are unique (a halo paints a pixel once): any sort is deterministic
        vals = self.vals[:n].index_select(0, perm)
        check(_capi.lib().ap_segment_accumulate(n, ptr(keys), ptr(vals), self.shift, int(pix0), ptr(d_map),
                                                int(d_map.dtype == torch.float32), stream_ptr()))
        return n


class Timers:
    """CUDA-event ticks on the launching stream; phase_ms() sums the time between consecutive ticks by name."""

    def __init__(self):
        self.ticks = []

    def tick(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.ticks.append((name, ev))

    def phase_ms(self, into=None, scale=1.0):
        out = {} if into is None else into
        for (name, ev), (_, nxt) in zip(self.ticks[:-1], self.ticks[1:]):
            if name != "end":
                out[name] = out.get(name, 0.0) + ev.elapsed_time(nxt) * scale
        return out


def table_chunk(halos, nside, kind, args, scale, n_samples, method, d_map, d_count=None, timers=None):
    """@interpolate(@LOS_integrate(profile_3D)) for one chunk of halos: K1 stats (+ the bypass work list) ->
    nodes -> QUADPACK -> spline -> K3 paint -> bypass.  Six launches, nothing read back by the host."""
    uniform = method == "linspace"
    tick = timers.tick if timers is not None else (lambda name: None)
    tick("disc_count")
    n_pix, rmin, rmax, items, n_items = disc_count_items(halos, nside, n_samples)
    tick("table_nodes")
    nodes, n_nodes = table_nodes(halos, n_pix, rmin, rmax, n_samples, method)
    tick("los_nodes")
    values = los_nodes(kind, halos, args, nodes, n_nodes)
    tick("spline_build")
    coef = spline_build(nodes, values, n_nodes, uniform=uniform)
    tick("paint_table")
    paint_table(halos, nside, nodes, coef, n_nodes, scale, d_map, d_count, uniform=uniform)
    tick("paint_los_items")
    paint_los_items(kind, halos, nside, args, scale, items, n_items, d_map, d_count)
    tick("end")
    return n_pix


def run_chunks(plan, paint_chunk, chunk_done=None):
    """Drive one rank's share of a spray: paint_chunk(lo, hi) for every chunk of `plan` (positions within the
    rank's halo slice), then chunk_done(rel_lo, rel_hi) with the pixel range of the rank's band (relative to its
    first pixel) that became final -- the caller ships it on a copy stream.  Inside band_clip(plan.band)."""
    from ._bands import ring_start
    band = plan.band
    prev = band.ring_lo
    with band_clip(band):
        for k, (a, e, done) in enumerate(plan.chunks):
            if e > a:
                paint_chunk(a - plan.halo_lo, e - plan.halo_lo, k)
            if chunk_done is not None and done > prev:
                chunk_done(ring_start(band.nside, prev) - band.pix_lo, ring_start(band.nside, done) - band.pix_lo)
                prev = done


# ---------------------------------------------------------------------------------------------
# kernel wrappers
# ---------------------------------------------------------------------------------------------
CATALOG_OUT = ("D_c", "redshift", "lat", "lon", "theta", "phi", "D_a", "R_200c", "c_200c", "R_ang_200c", "rho_s",
               "R_s", "v_r", "v_th", "v_ph", "v_lat", "v_lon")   # order of the AP_CAT_* indices


def copy_d2h(host_tensor, d_src, stream=None):
    """asynchronous device -> host copy of a contiguous tensor into (page-locked) host memory"""
    assert host_tensor.numel() == d_src.numel() and host_tensor.element_size() == d_src.element_size()
    sp = c_ptr(stream.cuda_stream) if stream is not None else stream_ptr()
    check(_capi.lib().ap_copy_d2h_async(c_ptr(host_tensor.data_ptr()), ptr(d_src), d_src.numel() * d_src.element_size(), sp))


def catalog_build(cols, default_redshift, crit_dens_0, h, device=None):
    """Catalog.build_dataframe's derived columns on the device (ap_catalog_build).  cols: dict of host fp64
    arrays x, y, z, v_x, v_y, v_z, M_200c and optionally redshift.  Returns {column: numpy view} over ONE
    pinned (17, n) host buffer (kept alive by the arrays)."""
    dev = require_cuda(device)
    n = len(cols["x"])
    names = ["x", "y", "z", "v_x", "v_y", "v_z", "M_200c"] + (["redshift"] if "redshift" in cols else [])
    d_in = [to_device(np.ascontiguousarray(cols[k], dtype=np.float64), dev) for k in names]
    if "redshift" not in cols:
        d_in.append(None)
    d_out = torch.empty((len(CATALOG_OUT), n), dtype=torch.float64, device=dev)
    outs = (c_ptr * len(CATALOG_OUT))(*[d_out[k].data_ptr() for k in range(len(CATALOG_OUT))])
    check(_capi.lib().ap_catalog_build(n, *[ptr(t) for t in d_in], float(default_redshift), float(crit_dens_0),
                                       float(h), outs, stream_ptr()))
    h_out = torch.empty((len(CATALOG_OUT), n), dtype=torch.float64, pin_memory=True)
    h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    arr = h_out.numpy()
    return {name: arr[k] for k, name in enumerate(CATALOG_OUT)}


class disc_mode:
    """`with disc_mode(canvas.inclusive):` -- every disc query issued by this thread inside the block is
    healpy's query_disc(inclusive=True) (oversampling factor 4, healpy's default; Canvas(..., inclusive=True),
    paint_bucket.py:782, :1228) instead of the exclusive one.  Thread-local in the C library (ap_disc_mode)."""

    def __init__(self, inclusive, fact=4):
        self.fact = int(fact) if inclusive else 0

    def __enter__(self):
        self.prev = _capi.lib().ap_disc_mode(self.fact)
        return self

    def __exit__(self, *exc):
        _capi.lib().ap_disc_mode(self.prev)
        return False


def disc_count(halos, nside, inclusive=False, want_range=False, want_rings=False):
    L = _capi.lib()
    dev = halos.device
    n_pix = torch.empty(halos.n, dtype=torch.int64, device=dev)
    n_rings = torch.empty(halos.n, dtype=torch.int32, device=dev) if want_rings else None
    rmin = torch.empty(halos.n, dtype=torch.float64, device=dev) if want_range else None
    rmax = torch.empty(halos.n, dtype=torch.float64, device=dev) if want_range else None
    s = halos.struct()
    check(L.ap_disc_count(nside, C.byref(s), int(bool(inclusive)), ptr(n_pix), ptr(n_rings), ptr(rmin),
                          ptr(rmax), stream_ptr()))
    return n_pix, n_rings, rmin, rmax


def disc_tie_audit(halos, nside, tol=1.0e-9):
    """Near-tie recheck of the disc query.  The device lists the halos that have a span boundary (or a
    ring-range decision) within `tol` pixels of an integer -- the only place where CUDA's libm and the CPU's can
    disagree on a pixel set; the host recomputes exactly those discs with its own libm (ap_query_disc_host) and the
    pixel count and pixel-id sum of both are compared.  Returns {"n_halos", "n_flagged", "n_flips", "flagged",
    "flips", "min_margin"}: `flips` are positions (within `halos`) whose device pixel set differs from the host's."""
    L = _capi.lib()
    dev = halos.device
    n = halos.n
    n_flag = torch.zeros(1, dtype=torch.int64, device=dev)
    flagged = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    margin = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    s = halos.struct()
    check(L.ap_disc_near_ties(nside, C.byref(s), float(tol), ptr(n_flag), ptr(flagged), n, ptr(margin), stream_ptr()))
    k = int(n_flag.item())
    idx = torch.sort(flagged[:k]).values
    out = {"n_halos": n, "n_flagged": k, "n_flips": 0, "flagged": idx.cpu().numpy(), "flips": np.empty(0, dtype=np.int64),
           "min_margin": float(margin[:n].min().item()) if n else float("inf")}
    if k == 0:
        return out
    sub = halos.view(0, n)
    sub.cols = {c: v.index_select(0, idx) for c, v in halos.cols.items()}
    sub.n = sub._n_all = k
    n_pix, _, _, _ = disc_count(sub, nside)
    off, pix, _, _ = disc_pixels(sub, nside, n_pix)
    seg = torch.repeat_interleave(torch.arange(k, device=dev), n_pix)
    d_sum = torch.zeros(k, dtype=torch.int64, device=dev).index_add_(0, seg, pix)
    host = {c: np.ascontiguousarray(sub.cols[c].cpu().numpy()) for c in ("x", "y", "z", "__radius__")}
    h_count = np.empty(k, dtype=np.int64)
    h_sum = np.empty(k, dtype=np.int64)
    check(L.ap_query_disc_host(nside, k, ptr(host["x"]), ptr(host["y"]), ptr(host["z"]), ptr(host["__radius__"]),
                               L.ap_disc_mode(-1), ptr(h_count), ptr(h_sum), None, 0))
    bad = (n_pix.cpu().numpy() != h_count) | (d_sum.cpu().numpy() != h_sum)
    out["n_flips"] = int(bad.sum())
    out["flips"] = out["flagged"][bad]
    return out


def disc_pixels(halos, nside, n_pix, lo=0, hi=None, want_R=False, want_vec=False):
    """Flat pixel list of halos [lo, hi): returns (offsets[int64, hi-lo+1], pix, R or None, R_vec or None)."""
    L = _capi.lib()
    dev = halos.device
    hi = halos.n if hi is None else hi
    off = torch.zeros(hi - lo + 1, dtype=torch.int64, device=dev)
    torch.cumsum(n_pix[lo:hi], 0, out=off[1:])
    total = int(off[-1].item())
    pix = torch.empty(total, dtype=torch.int64, device=dev)
    R = torch.empty(total, dtype=torch.float64, device=dev) if want_R else None
    Rv = torch.empty((total, 3), dtype=torch.float64, device=dev) if want_vec else None
    s = halos.struct(lo, hi)
    check(L.ap_disc_pixels(nside, C.byref(s), ptr(off), ptr(pix), ptr(R), ptr(Rv), stream_ptr()))
    return off, pix, R, Rv


def _fn(name, d_map):
    """fp64 or fp32-accumulate entry point according to the map's dtype"""
    return getattr(_capi.lib(), name + ("_f32" if d_map.dtype == torch.float32 else ""))


def scatter_add(pix, val, d_map):
    check(_fn("ap_scatter_add", d_map)(pix.numel(), ptr(pix), ptr(val), ptr(d_map), stream_ptr()))


def paint_profile(halos, nside, profile_id, params, d_map, d_count=None):
    """params: list of 1-element (broadcast) or n-element device tensors in device-argument order."""
    L = _capi.lib()
    n = len(params)
    arr = (c_ptr * max(n, 1))()
    strides = (c_i32 * max(n, 1))()
    for i, t in enumerate(params):
        assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
        arr[i] = t.data_ptr()
        strides[i] = 0 if t.numel() == 1 and halos.n != 1 else 1
        if strides[i] == 1 and t.numel() != halos.n:
            raise ValueError(f"template argument #{i} has {t.numel()} values for {halos.n} halos")
    s = halos.struct()
    check(_fn("ap_paint_profile", d_map)(int(profile_id), nside, C.byref(s), arr, strides, n, ptr(d_map),
                                         ptr(d_count), stream_ptr()))


def table_nodes(halos, n_pix, rmin, rmax, n_samples, method):
    L = _capi.lib()
    dev = halos.device
    nodes = torch.empty((halos.n, n_samples), dtype=torch.float64, device=dev)
    n_nodes = torch.empty(halos.n, dtype=torch.int32, device=dev)
    check(L.ap_table_nodes(halos.n, ptr(n_pix), ptr(rmin), ptr(rmax), n_samples,
                           {"linspace": 0, "logspace": 1}[method], ptr(nodes), ptr(n_nodes), stream_ptr()))
    return nodes, n_nodes


LOS_KINDS = {"b16_tau": 0, "b16_ne": 1, "b16_rho_gas": 2, "nfw_rho": 3}


def los_nodes(kind, halos, args, nodes, n_nodes):
    """QUADPACK projection of the 3-D profile `kind` at every (halo, node); args = its per-halo arguments."""
    values = torch.empty_like(nodes)
    a = list(args) + [None] * (3 - len(args))
    scratch = torch.empty((halos.n, 4), dtype=torch.float64, device=nodes.device)   # per-halo profile constants
    check(_capi.lib().ap_los_nodes(LOS_KINDS[kind], halos.n, ptr(a[0]), ptr(a[1]), ptr(a[2]), nodes.shape[1], ptr(nodes),
                                   ptr(n_nodes), ptr(values), ptr(scratch), stream_ptr()))
    return values


def disc_count_items(halos, nside, n_samples):
    """K1 with the bypass work list: (n_pix, rmin, rmax, items, n_items).  items holds halo * 32 + pixel ordinal
    for every pixel of every halo with 0 < n_pix < n_samples; n_items (device uint64) is their number."""
    dev = halos.device
    n = halos.n
    n_pix = torch.empty(n, dtype=torch.int64, device=dev)
    rmin = torch.empty(n, dtype=torch.float64, device=dev)
    rmax = torch.empty(n, dtype=torch.float64, device=dev)
    cap = n * (n_samples - 1)
    items = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
    n_items = torch.zeros(1, dtype=torch.int64, device=dev)
    s = halos.struct()
    check(_capi.lib().ap_disc_count_items(nside, C.byref(s), ptr(n_pix), ptr(rmin), ptr(rmax), n_samples, ptr(items), cap,
                                          ptr(n_items), stream_ptr()))
    return n_pix, rmin, rmax, items, n_items


def paint_los_items(kind, halos, nside, args, scale, items, n_items, d_map, d_count=None):
    """The `len(R) < n_samples` bypass of @interpolate from the device-built work list: one QUADPACK quadrature
    per (halo, pixel) item, every lane busy, the item count read on the device."""
    s = halos.struct()
    a = list(args) + [None] * (3 - len(args))
    check(_fn("ap_paint_los_items", d_map)(LOS_KINDS[kind], nside, C.byref(s), ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(scale),
                                           ptr(n_items), items.numel(), ptr(items), ptr(d_map), ptr(d_count), stream_ptr()))


def spline_build(nodes, values, n_nodes, uniform=False):
    n, ns = nodes.shape
    coef = torch.empty((n, ns - 1, 4), dtype=torch.float64, device=nodes.device)
    fn = _capi.lib().ap_spline_build_uniform if uniform else _capi.lib().ap_spline_build
    check(fn(n, ns, ptr(nodes), ptr(values), ptr(n_nodes), ptr(coef), stream_ptr()))
    return coef


def paint_table(halos, nside, nodes, coef, n_nodes, scale, d_map, d_count=None, uniform=False):
    s = halos.struct()
    check(_fn("ap_paint_table", d_map)(nside, C.byref(s), nodes.shape[1], int(bool(uniform)), ptr(nodes), ptr(coef),
                                       ptr(n_nodes),
                                     ptr(scale), ptr(d_map), ptr(d_count), stream_ptr()))


def cutouts(d_map, nside, lon, lat, lon_range, lat_range, xsize, ysize):
    d_map = d_map if d_map.dtype == torch.float64 else d_map.double()
    out = torch.empty((lon.numel(), ysize, xsize), dtype=torch.float64, device=d_map.device)
    check(_capi.lib().ap_cutouts(nside, ptr(d_map), lon.numel(), ptr(lon), ptr(lat), float(lon_range[0]),
                                 float(lon_range[1]), float(lat_range[0]), float(lat_range[1]), xsize, ysize,
                                 ptr(out), stream_ptr()))
    return out


def stack_cutouts(d_map, nside, lon, lat, lon_range, lat_range, xsize, ysize, weight=None, stack=None):
    d_map = d_map if d_map.dtype == torch.float64 else d_map.double()
    if stack is None:
        stack = torch.zeros((ysize, xsize), dtype=torch.float64, device=d_map.device)
    check(_capi.lib().ap_stack_cutouts(nside, ptr(d_map), lon.numel(), ptr(lon), ptr(lat), float(lon_range[0]),
                                       float(lon_range[1]), float(lat_range[0]), float(lat_range[1]), xsize,
                                       ysize, ptr(weight), ptr(stack), stream_ptr()))
    return stack


def ud_grade(d_map, nside_in, nside_out, pess=False, power=None):
    """healpy.ud_grade of a device RING map (fp64) -> new device map at nside_out"""
    d_map = d_map if d_map.dtype == torch.float64 else d_map.double()
    out = torch.empty(12 * nside_out * nside_out, dtype=torch.float64, device=d_map.device)
    check(_capi.lib().ap_ud_grade(nside_in, ptr(d_map), nside_out, int(bool(pess)), float(power or 0.0),
                                  int(bool(power)), ptr(out), stream_ptr()))
    return out


def ang2pix(nside, theta, phi, dev=None):
    """hp.ang2pix(nside, theta, phi) (RING) on the device; host arrays in, host int64 array out"""
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    th = np.ascontiguousarray(np.atleast_1d(theta), dtype=np.float64)
    ph = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(phi), th.shape), dtype=np.float64)
    d_th, d_ph = to_device(th, dev), to_device(ph, dev)
    d_pix = torch.empty(th.size, dtype=torch.int64, device=dev)
    check(_capi.lib().ap_ang2pix(int(nside), th.size, ptr(d_th), ptr(d_ph), ptr(d_pix), stream_ptr()))
    pix = d_pix.cpu().numpy()
    if (pix < 0).any():
        raise ValueError("THETA is out of range [0,pi]")     # healpy's message
    return pix


def pix2ang(nside, pix, dev=None):
    """hp.pix2ang(nside, ipix) (RING) on the device; returns host (theta, phi)"""
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    p = np.ascontiguousarray(np.atleast_1d(pix), dtype=np.int64)
    if p.size and (p.min() < 0 or p.max() >= 12 * int(nside) ** 2):
        raise ValueError("Wrong pixel number (it is not 0<=pixel<12*nside**2)")
    d_p = torch.from_numpy(p).to(dev)
    d_th = torch.empty(p.size, dtype=torch.float64, device=dev)
    d_ph = torch.empty_like(d_th)
    check(_capi.lib().ap_pix2ang(int(nside), p.size, ptr(d_p), ptr(d_th), ptr(d_ph), stream_ptr()))
    return d_th.cpu().numpy(), d_ph.cpu().numpy()


def _win(win, dev):
    if win is None:
        return None, None
    wy, wx = win
    return to_device(np.ascontiguousarray(wy, dtype=np.float64), dev), to_device(np.ascontiguousarray(wx, dtype=np.float64), dev)


def patch_rotate(d_patches, angle_deg, win=None, weight=None, stack=None):
    """transform.rotate_patch = scipy.ndimage.rotate(patch, angle, reshape=False) on a batch of device cutouts
    (n, ysize, xsize): cubic prefilter IN PLACE (d_patches becomes spline coefficients), then rotation;
    optionally fused with a taper window (wy, wx), per-cutout weights and accumulation into `stack`."""
    from scipy import special
    n, ysize, xsize = d_patches.shape
    dev = d_patches.device
    L = _capi.lib()
    check(L.ap_patch_spline_prefilter(n, ysize, xsize, ptr(d_patches), stream_ptr()))
    ang = np.broadcast_to(np.asarray(angle_deg, dtype=np.float64), (n,))
    d_c, d_s = to_device(np.ascontiguousarray(special.cosdg(ang)), dev), to_device(np.ascontiguousarray(special.sindg(ang)), dev)
    wy, wx = _win(win, dev)
    out = stack if stack is not None else torch.empty_like(d_patches)
    check(L.ap_patch_rotate(n, ysize, xsize, ptr(d_patches), ptr(d_c), ptr(d_s), ptr(wy), ptr(wx), ptr(weight),
                            int(stack is not None), ptr(out), stream_ptr()))
    return out


def patch_scale(d_patches, win=None, weight=None, stack=None):
    """transform.taper_patch (window = outer(hanning, hanning)) and / or per-cutout weights; in place, or
    accumulated into `stack`."""
    n, ysize, xsize = d_patches.shape
    wy, wx = _win(win, d_patches.device)
    check(_capi.lib().ap_patch_scale(n, ysize, xsize, ptr(d_patches), ptr(wy), ptr(wx), ptr(weight), ptr(stack),
                                     stream_ptr()))
    return stack if stack is not None else d_patches


# ---------------------------------------------------------------------------------------------
# spray flows
# ---------------------------------------------------------------------------------------------
def spray_flat(halos, nside, template, R_mode, host_cols, d_map, d_count=None, inclusive=False):
    """Arbitrary Python template: device disc query + geometry + scatter, template on the host.

    Follows the reference loop exactly (paint_bucket.py:2380-2389): one call per halo with the
    halo's full R (or R_vec) array and its catalog row as keyword arguments.
    """
    n_pix, _, _, _ = disc_count(halos, nside, inclusive)
    counts = n_pix.cpu().numpy()
    bounds = [0]
    run = 0
    for h in range(halos.n):
        if run > 0 and run + counts[h] > FLAT_BATCH_PIXELS:
            bounds.append(h)
            run = 0
        run += counts[h]
    bounds.append(halos.n)
    names = list(host_cols.keys())
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi == lo:
            continue
        off, pix, R, Rv = disc_pixels(halos, nside, n_pix, lo, hi, want_R=(R_mode == "R"),
                                      want_vec=(R_mode == "R_vec"))
        off_h = off.cpu().numpy()
        geo = (R if R_mode == "R" else Rv).cpu().numpy()
        vals = np.empty(off_h[-1], dtype=np.float64)
        for h in range(lo, hi):
            a, b = off_h[h - lo], off_h[h - lo + 1]
            kw = {k: host_cols[k][h] for k in names}
            kw[R_mode] = geo[a:b]
            v = template(**kw)
            if np.iscomplexobj(v):
                v = np.real(v)
            vals[a:b] = v
        if off_h[-1] > 0:
            scatter_add(pix, to_device(vals, halos.device), d_map)
        if d_count is not None:
            d_count += int(off_h[-1])
    return n_pix
