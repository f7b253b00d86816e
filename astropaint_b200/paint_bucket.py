"""Catalog / Canvas / Painter — the reference's object API over the B200 painting path.

Mirrors `astropaint/paint_bucket.py` of syasini/AstroPaint for the hot path only:

    Catalog(data)                              paint_bucket.py:41-364   (host, pandas; once per catalog)
    Canvas(catalog, nside, R_times, inclusive) paint_bucket.py:773-955, 1175-1323, 1601-1966
    Painter(template).spray(canvas, **kwargs)  paint_bucket.py:2263-2676

Everything per (halo, pixel) runs in hand-written sm_100a CUDA (csrc/k_*.cu) behind the C ABI
of include/astropaint_b200.h.  There is no CPU implementation of the path in this package.
"""
import inspect
import os
import re

import numpy as np
import pandas as pd

from .lib import transform
from . import _engine as eng

verbose = True


def _say(*a):
    if verbose:
        print(*a)


#########################################################
#                  Halo Catalog Object
#########################################################
class Catalog:
    """Halo catalog: positions [Mpc], velocities [km/s], M_200c [M_sun] (+ derived columns)."""

    def __init__(self, data=None, calculate_redshifts=False, default_redshift=0, *, device=None):
        """`device` (extra, keyword-only): where build_dataframe computes the derived columns -- None = the
        device kernel for large catalogs when a GPU is present, True / False to force."""
        self._build_counter = 0
        self.calculate_redshifts = calculate_redshifts
        self.default_redshift = default_redshift
        self.device_build = device
        if data is None:
            self.data = self._initialize_catalog(1)
        elif isinstance(data, str):
            if re.match(".*random.*box", data, re.IGNORECASE):
                self.generate_random_box()
            elif re.match(".*random.*shell", data, re.IGNORECASE):
                self.generate_random_shell()
            elif re.match(".*test.*", data, re.IGNORECASE):
                self.generate_test_box(configuration=["all"])
            else:
                self.load_from_csv(data)
        else:
            self.data = data

    # ------------------------ properties ------------------------
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, val):
        self._pinned = None
        self._order = None
        self._data = pd.DataFrame(val).reset_index(drop=True)
        self.size = len(self._data)
        self.box_size = self._get_box_size()
        if self._build_counter > 0:
            _say("Catalog data has been modified...\n")
        self.build_dataframe(calculate_redshifts=self.calculate_redshifts,
                             default_redshift=self.default_redshift, device=self.device_build)

    @staticmethod
    def _fingerprint(values):
        """cheap identity of a column's contents (microseconds, checked on every spray): where the array lives, its
        length and a CRC of ~2000 evenly spaced elements.  Catches a replaced column and bulk in-place edits
        (`df["v_r"] *= 2`); NOT an edit of a few elements -- reassign `catalog.data` (or call pin_memory() again)
        after such edits."""
        import zlib
        v = np.asarray(values)
        step = max(1, v.shape[0] // 2048) if v.ndim else 1
        return (v.__array_interface__["data"][0], v.shape, zlib.crc32(np.ascontiguousarray(v[::step]).tobytes()))

    def _geometry_fingerprint(self):
        d = self._data
        return tuple(self._fingerprint(d[k].values) for k in ("x", "y", "z", "R_ang_200c"))

    def staged(self, names):
        """The page-locked staging of pin_memory() if it is still current for the columns `names` (stale columns are
        staged again), else None."""
        if not getattr(self, "_pinned", None):
            return None
        for k in names:
            if k in self._pinned and self._pinned_fp.get(k) != self._fingerprint(self._data[k].values):
                import torch
                order = self.theta_order()[0]
                self._pinned[k] = torch.from_numpy(
                    np.ascontiguousarray(self._data[k].values[order], dtype=np.float64)).pin_memory()
                self._pinned_fp[k] = self._fingerprint(self._data[k].values)
        return self._pinned

    def theta_order(self):
        """(order, theta_sorted, r_ang_max): the catalog rows by increasing colatitude of the disc centre
        (formed from x, y, z exactly like hp.query_disc's centre, paint_bucket.py:1223-1225), that colatitude,
        and the largest R_ang_200c [arcmin].  The painting path walks the halos in this order: the halos that
        can touch one ring band of the map are then ONE contiguous slice (astropaint_b200/_bands.py), and
        consecutive halos write neighbouring rings.  Computed once per catalog, invalidated with `data`."""
        if getattr(self, "_order", None) is not None and self._order_fp != self._geometry_fingerprint():
            self._order = None           # x, y, z or R_ang_200c were edited in place: the order, reach and staging are stale
            self._pinned = None
        if getattr(self, "_order", None) is None:
            d = self._data
            x, y, z = (d[k].values.astype(np.float64, copy=False) for k in ("x", "y", "z"))
            key = np.arctan2(np.hypot(x, y), z)
            order = np.argsort(key, kind="stable")
            r_ang_max = float(np.nanmax(d["R_ang_200c"].values)) if len(d) else 0.0
            self._order = (order, key[order], r_ang_max)
            self._order_fp = self._geometry_fingerprint()
            self._reach = {}
        return self._order

    def disc_reach(self, R_times, extra=0.0):
        """(north, south) for the discs of radius R_times * R_ang_200c (+ extra [rad]) in colatitude order:
        north[i] = the smallest colatitude any disc of halos i, i+1, ... reaches, south[i] = the largest colatitude
        any disc of halos 0 .. i reaches.  Both are non-decreasing, so the halos that can touch a ring band, and the
        rings no later halo can touch, come out of binary searches (astropaint_b200/_bands.py).  Cached."""
        order, theta_sorted, _ = self.theta_order()
        key = (float(R_times), float(extra))
        if key not in self._reach:
            r = float(R_times) * transform.arcmin2rad(np.nan_to_num(self._data["R_ang_200c"].values[order])) + float(extra)
            north = np.minimum.accumulate((theta_sorted - r)[::-1])[::-1]
            south = np.maximum.accumulate(theta_sorted + r)
            self._reach = {key: (np.ascontiguousarray(north), south)}      # one entry: a catalog is sprayed with one R_times
        return self._reach[key]

    def pin_memory(self):
        """Stage every numeric column in page-locked host memory (torch pinned tensors), rows in colatitude
        order (`theta_order`), so that each spray uploads the slice of halos it needs with one DMA per column
        instead of pageable, gathered copies.  Optional; invalidated whenever `data` is reassigned."""
        import torch
        order = self.theta_order()[0]
        self._pinned = {k: torch.from_numpy(np.ascontiguousarray(self._data[k].values[order], dtype=np.float64)).pin_memory()
                        for k in self._data.columns if np.issubdtype(self._data[k].dtype, np.number)}
        self._pinned_fp = {k: self._fingerprint(self._data[k].values) for k in self._pinned}
        return self

    # ------------------------ sample data ------------------------
    @staticmethod
    def _data_dir():
        return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

    def load_from_csv(self, sample_name="MICE"):
        if not sample_name.endswith(".csv"):
            sample_name += ".csv"
        fname = sample_name if os.path.exists(sample_name) else os.path.join(self._data_dir(), sample_name)
        _say(f"Catalog loaded from:\n{fname}")
        self.data = pd.read_csv(fname, index_col=0)

    def save_to_csv(self, sample_name):
        if not sample_name.endswith(".csv"):
            sample_name += ".csv"
        fname = sample_name if os.path.dirname(sample_name) else os.path.join(self._data_dir(), sample_name)
        os.makedirs(os.path.dirname(fname), exist_ok=True)
        self.data.to_csv(fname)
        _say(f"Catalog saved to:\n{fname}")

    @staticmethod
    def _initialize_catalog(n_tot):
        names = ["x", "y", "z", "v_x", "v_y", "v_z", "M_200c"]
        return pd.DataFrame(np.zeros(n_tot, {"names": names, "formats": 7 * [np.float32]}))

    def _random_velocities_and_masses(self, catalog, v_max, mass_min, mass_max, n_tot):
        v = np.random.uniform(low=-v_max, high=v_max, size=(3, n_tot))
        catalog["v_x"], catalog["v_y"], catalog["v_z"] = v
        catalog["M_200c"] = np.exp(np.random.uniform(low=np.log(mass_min), high=np.log(mass_max), size=n_tot))

    def generate_random_box(self, box_size=50, v_max=100, mass_min=1E14, mass_max=1E15, n_tot=50000,
                            put_on_shell=False, inplace=True):
        """Uniform random halos in a box centred on the observer (legacy np.random call order of the
        reference: positions, velocities, log-uniform masses)."""
        catalog = self._initialize_catalog(n_tot)
        _say("generating random catalog...\n")
        x, y, z = np.random.uniform(low=-box_size / 2, high=box_size / 2, size=(3, n_tot))
        if put_on_shell:
            (x, y, z) = box_size * np.true_divide((x, y, z), np.linalg.norm((x, y, z), axis=0))
        catalog["x"], catalog["y"], catalog["z"] = x, y, z
        self._random_velocities_and_masses(catalog, v_max, mass_min, mass_max, n_tot)
        if inplace:
            self.data = pd.DataFrame(catalog)
        else:
            return pd.DataFrame(catalog)

    def generate_random_shell(self, shell_radius=50, v_max=100, mass_min=1E14, mass_max=1E15, n_tot=50000,
                              inplace=True):
        catalog = self._initialize_catalog(n_tot)
        _say("generating random catalog...\n")
        u, v = np.random.uniform(low=0, high=1, size=(2, n_tot))
        phi = 2 * np.pi * u
        theta = np.arccos(2 * v - 1)
        catalog["x"] = np.sin(theta) * np.cos(phi) * shell_radius
        catalog["y"] = np.sin(theta) * np.sin(phi) * shell_radius
        catalog["z"] = np.cos(theta) * shell_radius
        self._random_velocities_and_masses(catalog, v_max, mass_min, mass_max, n_tot)
        if inplace:
            self.data = pd.DataFrame(catalog)
        else:
            return pd.DataFrame(catalog)

    def generate_test_box(self, configuration=["all"], distance=100, mass=1E15, inplace=True):
        """Up to six halos on the +-x, +-y, +-z axes."""
        where = {"front": (1, 0, 0), "back": (-1, 0, 0), "left": (0, 1, 0), "right": (0, -1, 0),
                 "top": (0, 0, 1), "bottom": (0, 0, -1)}
        if "all" in configuration:
            configuration = where.keys()
        rows = []
        for key in configuration:
            x, y, z = where[key]
            rows.append({"x": x * distance, "y": y * distance, "z": z * distance,
                         "v_x": 0.0, "v_y": 0.0, "v_z": 0.0, "M_200c": mass})
        catalog = pd.DataFrame(rows, columns=["x", "y", "z", "v_x", "v_y", "v_z", "M_200c"])
        if inplace:
            self.data = catalog
        else:
            return catalog

    # ------------------------ methods ------------------------
    #: catalogs at least this long get their derived columns from the device kernel (ap_catalog_build) when a
    #: GPU is present: 1e7 halos take 3.2 s in numpy/pandas on the host against a 0.22 s spray
    DEVICE_BUILD_MIN = 200_000

    def build_dataframe(self, calculate_redshifts=False, default_redshift=0, device=None):
        """Derived columns the painting path reads (reference: paint_bucket.py:308-364).
        device: None = the device kernel for catalogs of DEVICE_BUILD_MIN halos or more if CUDA is available,
        True / False to force."""
        self._build_counter = 1
        d = self.data
        if device is None:
            device = len(d) >= self.DEVICE_BUILD_MIN and eng.torch.cuda.is_available()
        if device:
            return self._build_dataframe_device(calculate_redshifts, default_redshift)
        x, y, z = (d[k].values.astype(np.float64) for k in ("x", "y", "z"))
        # cartesian -> spherical: D_c, latitude in [-pi/2, pi/2], longitude in [0, 2 pi)
        s = np.hypot(x, y)
        d["D_c"] = np.hypot(s, z)
        lat = np.arctan2(z, s)
        lon = np.arctan2(y, x)
        lon = np.where(lon < 0, lon + 2 * np.pi, lon)
        if calculate_redshifts:
            d["redshift"] = transform.D_c_to_redshift(d["D_c"].values)
        elif "redshift" not in d.columns:
            d["redshift"] = pd.Series([default_redshift] * len(d))
        d["lat"], d["lon"] = lat, lon
        d["theta"] = np.pi / 2 - d["lat"]
        d["phi"] = d["lon"]
        d["lon"], d["lat"] = np.rad2deg((d["lon"], d["lat"]))
        d["D_a"] = transform.D_c_to_D_a(d["D_c"], d["redshift"])
        d["R_200c"] = transform.M_200c_to_R_200c(d["M_200c"], d["redshift"])
        d["c_200c"] = transform.M_200c_to_c_200c(d["M_200c"], d["redshift"])
        d["R_ang_200c"] = transform.radius_to_angsize(d["R_200c"], d["D_a"], arcmin=True)
        d["rho_s"] = transform.M_200c_to_rho_s(d["M_200c"], d["redshift"], d["R_200c"], d["c_200c"])
        d["R_s"] = np.true_divide(d["R_200c"], d["c_200c"])
        J = transform.get_cart2sph_jacobian(d["theta"].values, d["phi"].values)
        v_cart = np.array([d["v_x"], d["v_y"], d["v_z"]])
        d["v_r"], d["v_th"], d["v_ph"] = np.einsum("ij...,i...->j...", J, v_cart)
        d["v_lat"] = -d["v_th"]
        d["v_lon"] = d["v_ph"]

    def _build_dataframe_device(self, calculate_redshifts, default_redshift):
        """the same 17 columns from csrc/k_core.cu::catalog_build_kernel, one pinned read-back"""
        d = self.data
        cols = {k: d[k].values.astype(np.float64, copy=False) for k in ("x", "y", "z", "v_x", "v_y", "v_z", "M_200c")}
        if calculate_redshifts:   # z_at_value per halo in the reference (:321-323); table inversion here, on the host
            cols["redshift"] = np.asarray(transform.D_c_to_redshift(np.hypot(np.hypot(cols["x"], cols["y"]), cols["z"])))
        elif "redshift" in d.columns:
            cols["redshift"] = d["redshift"].values.astype(np.float64, copy=False)
        out = eng.catalog_build(cols, default_redshift, transform.crit_dens_0, transform.h)
        had_z = "redshift" in d.columns and not calculate_redshifts
        for name in eng.CATALOG_OUT:
            if name == "redshift" and had_z:
                continue   # the user's column stays as given (dtype included)
            d[name] = out[name]

    def _get_box_size(self):
        return tuple(self.data[k].max() - self.data[k].min() for k in ("x", "y", "z"))

    def move_to_box_center(self):
        for k in ("x", "y", "z"):
            self.data[k] -= (self.data[k].max() + self.data[k].min()) / 2
        self.data = self.data

    def _cut(self, column, lo, hi, inplace):
        data = self.data[(self.data[column] > lo) & (self.data[column] < hi)]
        if inplace:
            self.data = data
        else:
            return data

    def cut_M_200c(self, mass_min=0., mass_max=np.inf, inplace=True):
        return self._cut("M_200c", mass_min, mass_max, inplace)

    def cut_R_200c(self, R_min=0., R_max=np.inf, inplace=True):
        return self._cut("R_200c", R_min, R_max, inplace)

    def cut_R_ang_200c(self, R_ang_min=0., R_ang_max=np.inf, inplace=True):
        return self._cut("R_ang_200c", R_ang_min, R_ang_max, inplace)

    def cut_D_c(self, D_min=0., D_max=np.inf, inplace=True):
        return self._cut("D_c", D_min, D_max, inplace)

    def cut_D_a(self, D_min=0., D_max=np.inf, inplace=True):
        return self._cut("D_a", D_min, D_max, inplace)

    def cut_redshift(self, redshift_min=0., redshift_max=np.inf, inplace=True):
        return self._cut("redshift", redshift_min, redshift_max, inplace)

    def cut_lon_lat(self, lon_range=[0, 360], lat_range=[-90, 90], inplace=True):
        d = self.data
        data = d[(d.lon > lon_range[0]) & (d.lon < lon_range[1]) & (d.lat > lat_range[0]) & (d.lat < lat_range[1])]
        if inplace:
            self.data = data
        else:
            return data

    def cut_theta_phi(self, theta_range=[0, np.pi], phi_range=[0, 2 * np.pi], inplace=True):
        d = self.data
        data = d[(d.theta > theta_range[0]) & (d.theta < theta_range[1]) &
                 (d.phi > phi_range[0]) & (d.phi < phi_range[1])]
        if inplace:
            self.data = data
        else:
            return data


#########################################################
#                  Canvas Object
#########################################################
class Canvas:
    """Full-sky HEALPix (RING) canvas; the fp64 pixel array lives in HBM while painting."""

    def __init__(self, catalog, nside, mode="healpy", R_times=1, inclusive=False, *, accumulate="fp64",
                 deterministic=False, band_balance="area"):
        """`accumulate="fp32"` (extra, keyword-only) keeps the map in float32: templates are still
        evaluated in fp64, each value is rounded once before it is accumulated (parity bar 1e-4).
        `deterministic=True` (extra, keyword-only): sprays of templates with a device twin accumulate by
        sort-then-segmented-reduce instead of atomics -- every pixel receives its values in catalog order, as in the
        reference's serial loop, so the map is bit-reproducible from run to run (slower: the (pixel, value) pairs are
        materialised and sorted).
        `band_balance` (extra, keyword-only; several ranks only): "area" gives every rank an equal share of the sphere
        (balances the painting), "readback" sizes the bands by each GPU's measured device->host bandwidth
        (parallel.readback_weights) so that `canvas.pixels` lands on all ranks together -- for workflows that read the
        map back after every spray on hosts whose PCIe paths are not equal."""
        assert mode == "healpy", "currently only full sky is supported"
        assert accumulate in ("fp64", "fp32")
        assert band_balance in ("area", "readback")
        self.accumulate = accumulate
        self.deterministic = bool(deterministic)
        self.band_balance = band_balance
        self._nside = int(nside)
        self._npix = 12 * self._nside * self._nside
        self._lmax = 3 * self._nside - 1
        self.R_times = R_times
        self.inclusive = inclusive
        self._catalog = Catalog() if catalog is None else catalog
        self.template_name = None
        self.stack = None
        self._host = None          # _bands.SharedHostMap: page-locked host copy the device map is read back into
        self._host_stale = False   # clean() / the setter ran since the buffer was last handed out
        self._prefetch = None      # (copy stream, band pixels already queued) of a streaming spray
        self._dfull = None
        self._band_cache = None
        self._band_override = None
        self._dmap = None
        self._dmap_spare = None
        self.clean()
        self.instantiate_discs()

    # immutables
    @property
    def nside(self):
        return self._nside

    @property
    def npix(self):
        return self._npix

    @property
    def lmax(self):
        return self._lmax

    @property
    def ell(self):
        return np.arange(self._lmax + 1)

    @property
    def R_max(self):
        return self.R_times * self.catalog.data["R_200c"].max()

    @property
    def centers_D_a(self):
        return self._catalog.data.D_a

    # ---- pixel state: "zero" | "host" | "device" ------------------------------------------------
    # "device": the authoritative copy is the rank's ring band of the map in HBM (the whole map when there is
    # one rank); "host": it is the numpy array `_pixels` (every rank sees the same one); "zero": nothing allocated.
    def _band(self):
        from . import parallel, _bands
        if self._band_override is not None:      # tests: paint one rank's band of an N-rank layout on one GPU
            return self._band_override
        rank, ws = parallel.world()
        b = self._band_cache
        if b is None or (b.rank, b.world_size, b.nside) != (rank, ws, self.nside):
            if b is not None and self._state == "device":
                raise RuntimeError("the process group changed while the canvas held a device map; read canvas.pixels "
                                   "(or clean()) before initialising / destroying torch.distributed")
            weights = parallel.readback_weights() if (ws > 1 and self.band_balance == "readback") else None
            self._band_cache = b = _bands.Band(self.nside, rank, ws, weights)
        return b

    @property
    def pixels(self):
        """Host view of the map (np.ndarray, fp64).  Reading it after a spray copies the device map back once
        (only the part the spray has not already streamed); the host array is then authoritative (users index
        and mutate it in place, e.g. README.md:40-48), so the next spray uploads it again.  `pixels_device`
        avoids both copies.  With several ranks the call is collective: every rank copies its own band into one
        host array shared by the node and every rank gets a view of the whole map."""
        if self._state == "zero":
            self._pixels = np.zeros(self.npix, dtype=self._np_dtype)
        elif self._state == "device":
            import torch
            from . import parallel
            band = self._band()
            if self._prefetch is not None:
                # the spray already streamed the finished rings to the host on a copy stream; fetch the rest
                stream, upto = self._prefetch
                host = self._host
                self._prefetch = None
            else:
                host = self._reusable_host() or self._new_host()
                stream, upto = None, 0
            if upto < band.npix:
                eng.copy_d2h(host.band_view(upto, band.npix), self._dmap[upto:])
            if stream is not None:
                stream.synchronize()
            torch.cuda.current_stream().synchronize()
            parallel.barrier()           # every rank's band has landed before anyone reads the map
            self._pixels = host.array
            self._dmap_spare = self._dmap
            self._dmap = None
            self._dfull = None
        self._state = "host"
        return self._pixels

    def _new_host(self):
        from . import _bands
        self._host = _bands.SharedHostMap(self.npix, self._np_dtype, self._band())
        return self._host

    def _reusable_host(self):
        """The read-back buffer of an earlier `.pixels`, if it may be written again: never once the array has been
        handed to the user and is still referenced (the reference's clean() allocates a fresh array,
        paint_bucket.py:1316, so a map kept across clean() must survive).  Collective with several ranks."""
        from . import parallel
        if self._host is None:
            return None
        ok = (not self._host_stale) or (not self._host.handed_out())
        ok = ok and self._host.dtype == np.dtype(self._np_dtype) and self._host.npix == self.npix
        if not parallel.all_agree(ok):
            self._host = None
        self._host_stale = False
        return self._host

    def _cancel_prefetch(self):
        if self._prefetch is not None:
            self._prefetch[0].synchronize()
            self._prefetch = None

    @pixels.setter
    def pixels(self, val):
        self._cancel_prefetch()
        self._pixels = val
        self._dmap = None
        self._dfull = None
        self._host_stale = True
        self._state = "host"
        self._pixel_is_outdated = False
        self._alm_is_outdated = True
        self._Cl_is_outdated = True

    @property
    def _np_dtype(self):
        return np.float32 if self.accumulate == "fp32" else np.float64

    @property
    def _torch_dtype(self):
        import torch
        return torch.float32 if self.accumulate == "fp32" else torch.float64

    def _device_band(self):
        """The rank's band of the map as a torch CUDA tensor (the whole map with one rank); allocated, zeroed or
        uploaded from the host copy on first use.  This is what the painting kernels accumulate into."""
        import torch
        dev = eng.require_cuda()
        band = self._band()
        if self._state != "device":
            self._cancel_prefetch()
            self._dfull = None
        if self._state == "zero":
            sp = self._dmap_spare
            if sp is not None and sp.device == dev and sp.numel() == band.npix and sp.dtype == self._torch_dtype:
                self._dmap = sp.zero_()
            else:
                self._dmap = torch.zeros(band.npix, dtype=self._torch_dtype, device=dev)
            self._dmap_spare = None
        elif self._state == "host":
            host = np.broadcast_to(np.asarray(self._pixels, dtype=self._np_dtype), (self.npix,))
            host = np.ascontiguousarray(host[band.pix_lo:band.pix_hi])
            self._dmap = torch.from_numpy(host).to(dev)
            self._pixels = None
        self._state = "device"
        return self._dmap

    @property
    def pixels_device(self):
        """The whole map as a torch CUDA tensor (fp64[npix]).  With one rank this is the tensor the kernels paint
        into; with several it is gathered from the owners' bands over NVLink (collective) and cached until the
        map changes."""
        d_band = self._device_band()
        band = self._band()
        if band.whole:
            return d_band
        if self._dfull is None:
            from . import parallel
            self._dfull = parallel.gather_bands(d_band, band, self._torch_dtype)
        return self._dfull

    # ---- map sinks ---------------------------------------------------------------
    def ud_grade(self, nside_out, inplace=True, **kwargs):
        """healpy.ud_grade of canvas.pixels on the device (reference: paint_bucket.py:2233-2257).  kwargs: `pess`,
        `power` (healpy's); RING in and out.  inplace=True also switches the canvas to nside_out."""
        unknown = set(kwargs) - {"pess", "power"}
        if unknown:
            raise TypeError(f"ud_grade() got unsupported healpy arguments {sorted(unknown)}")
        whole = self._band().whole
        new_map = eng.ud_grade(self.pixels_device, self.nside, int(nside_out), kwargs.get("pess", False),
                               kwargs.get("power", None))
        if self.accumulate == "fp32":
            new_map = new_map.float()
        if not inplace:
            return new_map.cpu().numpy()
        self._cancel_prefetch()
        self._state = "zero"           # the old map is dropped below; lets _band() re-derive the ownership
        self._band_cache = None
        self._nside = int(nside_out)
        self._npix = 12 * self._nside * self._nside
        self._lmax = 3 * self._nside - 1
        if not whole:     # several ranks: hold the new map on the host (every rank has all of it)
            self._dmap, self._dmap_spare, self._host, self._dfull = None, None, None, None
            self._pixels = new_map.cpu().numpy()
            self._state = "host"
        else:
            self._dmap, self._dmap_spare, self._host, self._pixels, self._dfull = new_map, None, None, None, None
            self._state = "device"
        self.discs = self.Disc(self.catalog, self.nside, self.R_times, self.inclusive)
        self._mark_painted()

    # ---- map files (reference: paint_bucket.py:1478-1558; hp.write_map / hp.read_map restated in lib/fits_map.py)
    def _map_filename(self, filename, prefix, suffix):
        if prefix and str(prefix)[-1] != "_":
            prefix = "".join([prefix, "_"])
        if suffix and str(suffix)[0] != "_":
            suffix = "".join(["_", suffix])
        if filename is None:
            filename = f"{str(prefix or '')}{self.template_name}_NSIDE={self.nside}{str(suffix or '')}.fits"
        return filename

    def save_map_to_file(self, filename=None, prefix=None, suffix=None, overwrite=True):
        """save the map as a HEALPix FITS file (healpy.write_map's layout: one float32 column, RING).  With several
        ranks the call is collective (`pixels` assembles the bands) and rank 0 writes."""
        from .lib import fits_map
        from . import parallel
        filename = self._map_filename(filename, prefix, suffix)
        pixels = self.pixels
        if parallel.world()[0] == 0:
            fits_map.write_map(filename, pixels, overwrite=overwrite)
        parallel.barrier()

    def load_map_from_file(self, filename=None, prefix=None, suffix=None, inplace=True):
        """load a HEALPix FITS map (healpy.read_map) into canvas.pixels, or return it"""
        from .lib import fits_map
        m = fits_map.read_map(self._map_filename(filename, prefix, suffix))
        if inplace:
            self.pixels = m
        else:
            return m

    def discs_coverage(self):
        """The 0/1 map Canvas.show_discs draws (paint_bucket.py:1433-1458: junk_pixels[disc] = 1 for every halo's
        disc), built on the device; plotting itself is outside the path."""
        import torch
        dev = eng.require_cuda()
        halos = eng.DeviceHalos(self.catalog.data, self.R_times, dev)
        cover = torch.zeros(self.npix, dtype=torch.float64, device=dev)
        one = torch.ones(1, dtype=torch.float64, device=dev)
        with eng.disc_mode(self.inclusive):
            eng.paint_profile(halos, self.nside, 0, [one], cover)      # AP_CONSTANT: hits per pixel
        return (cover > 0).to(torch.float64).cpu().numpy()

    def _mark_painted(self):
        # the reference's serial spray leaves alm/Cl flags untouched (:2387) while the threaded one
        # goes through the setter (:2433); a drop-in marks them stale in both cases.
        self._pixel_is_outdated = False
        self._alm_is_outdated = True
        self._Cl_is_outdated = True

    @property
    def alm(self):
        if self._alm_is_outdated:
            raise NotImplementedError("spherical-harmonic transforms are outside the painting hot path")
        return self._alm

    @alm.setter
    def alm(self, val):
        self._alm = val
        self._alm_is_outdated = False
        self._pixel_is_outdated = True
        self._Cl_is_outdated = True

    @property
    def Cl(self):
        if self._Cl_is_outdated:
            raise NotImplementedError("spherical-harmonic transforms are outside the painting hot path")
        return self._Cl

    @property
    def catalog(self):
        return self._catalog

    @catalog.setter
    def catalog(self, val):
        self._catalog = val
        self.instantiate_discs()

    def clean(self):
        """Set all pixels to zero (paint_bucket.py:1307-1323)."""
        self._cancel_prefetch()
        self._pixels = None
        self._dfull = None
        self._host_stale = True    # a host array handed out before this point is the user's from now on
        # keep the HBM allocation for the next spray instead of returning 8*npix bytes to the allocator
        self._dmap_spare = self._dmap if self._dmap is not None else self._dmap_spare
        self._dmap = None
        self._state = "zero"
        self._pixel_is_outdated = False
        self._alm = None
        self._alm_is_outdated = True
        self._Cl = np.zeros(self._lmax + 1)
        self._Cl_is_outdated = True

    # ------------------------ Disc inner class ------------------------
    class Disc:
        """Per-halo disc generators (paint_bucket.py:1175-1290), computed by the device kernels in
        batches and yielded as numpy arrays."""

        BATCH = 4096

        def __init__(self, catalog, nside, R_times, inclusive):
            self.catalog = catalog
            self.nside = nside
            self.R_times = R_times
            self.inclusive = inclusive
            self.center_D_a = self.catalog.data.D_a

        def _halos(self, halo_list):
            try:
                if isinstance(halo_list, str) and halo_list.lower() == "all":
                    halo_list = range(self.catalog.size)
            except AttributeError:
                pass
            assert hasattr(halo_list, "__iter__")
            return np.asarray(list(halo_list), dtype=np.int64)

        def _gen(self, halo_list, want_R=False, want_vec=False):
            idx = self._halos(halo_list)
            dev = eng.require_cuda()
            for b in range(0, len(idx), self.BATCH):
                rows = idx[b:b + self.BATCH]
                halos = eng.DeviceHalos(self.catalog.data, self.R_times, dev, rows=rows)
                with eng.disc_mode(self.inclusive):
                    n_pix, _, _, _ = eng.disc_count(halos, self.nside)
                    off, pix, R, Rv = eng.disc_pixels(halos, self.nside, n_pix, want_R=want_R, want_vec=want_vec)
                off = off.cpu().numpy()
                pix = pix.cpu().numpy()
                R = R.cpu().numpy() if R is not None else None
                Rv = Rv.cpu().numpy() if Rv is not None else None
                for i in range(len(rows)):
                    a, e = off[i], off[i + 1]
                    yield rows[i], pix[a:e], (R[a:e] if R is not None else None), (Rv[a:e] if Rv is not None else None)

        def _center_ipix(self, halo_list):
            """hp.ang2pix of the listed centres in one device launch (ap_ang2pix)"""
            d = self.catalog.data
            rows = np.asarray(list(self._halos(halo_list)), dtype=np.int64)
            return rows, eng.ang2pix(self.nside, d.theta.values[rows], d.phi.values[rows])

        def gen_center_ipix(self, halo_list="All"):
            """ipix of the halo centres (paint_bucket.py:1175-1185)"""
            for ipix in self._center_ipix(halo_list)[1]:
                yield int(ipix)

        def gen_center_ang(self, halo_list="All", snap2pixel=False):
            """(theta, phi) of the halo centres; snap2pixel=True returns the centre pixel's own coordinates
            (paint_bucket.py:1187-1203)"""
            if snap2pixel:
                theta, phi = eng.pix2ang(self.nside, self._center_ipix(halo_list)[1])
                for t, f in zip(theta, phi):
                    yield (t, f)
                return
            d = self.catalog.data
            for halo in self._halos(halo_list):
                yield (d.theta[halo], d.phi[halo])

        def find_centers_indx(self):
            """sets .center_index to the centre pixel of every halo (paint_bucket.py:984-1000)"""
            self.center_index = self._center_ipix("All")[1]

        def find_centers_ang(self):
            """sets .center_ang to [theta, phi] of every halo (paint_bucket.py:1002-1016)"""
            d = self.catalog.data
            self.center_ang = np.asarray([d.theta.to_list(), d.phi.to_list()])

        def gen_center_vec(self, halo_list="All"):
            d = self.catalog.data
            for halo in self._halos(halo_list):
                th, ph = d.theta[halo], d.phi[halo]
                yield np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])

        def gen_pixel_index(self, halo_list="All"):
            for _, pix, _, _ in self._gen(halo_list):
                yield pix

        def gen_pixel_ang(self, halo_list="All"):
            """(theta, phi) of the pixels of each halo's disc: hp.pix2ang(nside, pixel_index)
            (paint_bucket.py:1231-1238)"""
            for pix in self.gen_pixel_index(halo_list):
                theta, phi = eng.pix2ang(self.nside, pix) if len(pix) else (np.empty(0), np.empty(0))
                yield (theta, phi)

        def gen_cent2pix_rad(self, halo_list="All"):
            for halo, _, R, _ in self._gen(halo_list, want_R=True):
                yield R / self.center_D_a[halo]

        def gen_cent2pix_mpc(self, halo_list="All"):
            for _, _, R, _ in self._gen(halo_list, want_R=True):
                yield R

        def gen_cent2pix_mpc_vec(self, halo_list="All"):
            for _, _, _, Rv in self._gen(halo_list, want_vec=True):
                yield Rv

        def gen_cent2pix_hat(self, halo_list="All"):
            for _, _, _, Rv in self._gen(halo_list, want_vec=True):
                yield Canvas._normalize_vec(Rv)

        def gen_pixel_vec(self, halo_list="All"):
            for (_, _, _, Rv), c, halo in zip(self._gen(halo_list, want_vec=True), self.gen_center_vec(halo_list),
                                              self._halos(halo_list)):
                yield Rv / self.center_D_a[halo] + c

    def instantiate_discs(self):
        self.discs = self.Disc(self.catalog, self.nside, self.R_times, self.inclusive)

    @staticmethod
    def _normalize_vec(vec, axis=-1):
        return np.true_divide(vec, np.expand_dims(np.linalg.norm(vec, axis=axis), axis=axis))

    # ------------------------ cutouts / stacking ------------------------
    def _cutout_setup(self, halo_list, lon_range, lat_range, xpix, ypix):
        if lat_range is None:
            lat_range = lon_range
        try:
            if isinstance(halo_list, str) and halo_list == "all":
                halo_list = range(self.catalog.size)
        except ValueError:
            pass
        idx = np.asarray(list(halo_list), dtype=np.int64)
        lonra = [float(lon_range[0]), float(lon_range[1])]
        latra = [float(lat_range[0]), float(lat_range[1])]
        span = np.mod(lonra[1] - lonra[0], 360) or 360.0
        if ypix is None:
            ypix = int(round(xpix * (latra[1] - latra[0]) / span))
        return idx, lonra, latra, int(xpix), int(ypix)

    CUTOUT_BATCH_BYTES = 1 << 28

    def cutouts(self, halo_list="all", lon_range=[-1, 1], lat_range=None, xpix=200, ypix=None,
                apply_func=None, func_kwargs=dict()):
        """Generate (ypix, xpix) Cartesian-projection cutouts centred on each halo, optionally passed
        through `apply_func` with per-halo or shared `func_kwargs` (paint_bucket.py:1601-1743)."""
        idx, lonra, latra, xpix, ypix = self._cutout_setup(halo_list, lon_range, lat_range, xpix, ypix)
        if apply_func is not None:
            if not isinstance(apply_func, list):
                apply_func = [apply_func]
                assert not isinstance(func_kwargs, list)
                func_kwargs = [func_kwargs]
            else:
                assert isinstance(func_kwargs, list)
            for f_kw in func_kwargs:
                for key, value in f_kw.items():
                    if not hasattr(value, "__len__"):
                        f_kw[key] = [value]
        d_map = self.pixels_device
        lon = self.catalog.data["lon"].values
        lat = self.catalog.data["lat"].values
        batch = max(1, self.CUTOUT_BATCH_BYTES // (8 * xpix * ypix))
        plan = self._patch_plan(apply_func, func_kwargs, xpix, ypix)
        for b in range(0, len(idx), batch):
            rows = idx[b:b + batch]
            d_lon = eng.to_device(lon[rows], d_map.device)
            d_lat = eng.to_device(lat[rows], d_map.device)
            d_cuts = eng.cutouts(d_map, self.nside, d_lon, d_lat, lonra, latra, xpix, ypix)
            if plan is not None:      # every function of the chain has a device twin: no per-halo host work
                cuts = self._run_patch_plan(plan, d_cuts, rows).cpu().numpy()
                for i in range(len(rows)):
                    yield cuts[i]
                continue
            cuts = d_cuts.cpu().numpy()
            for i, halo in enumerate(rows):
                cut_out = cuts[i]
                if apply_func is not None:
                    for func, f_kw in zip(apply_func, func_kwargs):
                        func_dict = {key: (value[halo] if len(value) > 1 else value[0])
                                     for key, value in f_kw.items()}
                        cut_out = func(cut_out, **func_dict)
                yield cut_out

    @staticmethod
    def _patch_plan(apply_func, func_kwargs, xpix, ypix):
        """[(op, per-halo-or-scalar value)] when every callable of the (already list-normalised) apply_func
        chain is one of the library operators with a device twin (transform.rotate_patch / taper_patch /
        scale_patch), else None."""
        if apply_func is None:
            return None
        plan = []
        for func, f_kw in zip(apply_func, func_kwargs):
            spec = getattr(func, "_ap_patch_op", None)
            if spec is None:
                return None
            op, key = spec
            extra = set(f_kw) - ({key} if key else set())
            if extra or (key and key not in f_kw and op == "rotate"):
                return None
            if op == "taper" and xpix != ypix:
                return None       # the reference's window is square (np.hanning(patch.shape[0]) twice)
            plan.append((op, f_kw.get(key) if key else None))
        return plan

    def _run_patch_plan(self, plan, d_cuts, rows, stack=None):
        """Apply the chain to the device batch d_cuts (n, ypix, xpix) for catalog rows `rows`; with `stack`
        the last operator accumulates into it (returns stack), otherwise returns the processed batch."""
        def per_halo(value):    # the reference indexes length-N kwargs by halo label, scalars are shared
            v = np.asarray(value, dtype=np.float64).reshape(-1)
            return v[rows] if len(v) > 1 else np.full(len(rows), v[0])
        dev = d_cuts.device
        n_ops = len(plan)
        for k, (op, value) in enumerate(plan):
            st = stack if k == n_ops - 1 else None
            if op == "rotate":
                d_cuts = eng.patch_rotate(d_cuts, per_halo(value), stack=st)
            elif op == "taper":
                han = np.hanning(d_cuts.shape[1])
                d_cuts = eng.patch_scale(d_cuts, win=(han, han), stack=st)
            else:
                w = eng.to_device(per_halo(1.0 if value is None else value), dev)
                d_cuts = eng.patch_scale(d_cuts, weight=w, stack=st)
        return d_cuts

    def stack_cutouts(self, halo_list="all", lon_range=[-1, 1], lat_range=None, xpix=200, ypix=None,
                      inplace=True, parallel=False, apply_func=None, func_kwargs=dict()):
        """Sum of the cutouts of `halo_list` (paint_bucket.py:1745-1941).  Without `apply_func` the
        cutouts are never materialised: one fused kernel projects and accumulates.  `parallel` is
        accepted for compatibility (the GPU path is always parallel)."""
        if ypix is None:
            ypix = xpix
        if apply_func is None:
            from . import parallel
            idx, lonra, latra, xpix, ypix = self._cutout_setup(halo_list, lon_range, lat_range, xpix, ypix)
            d_map = self.pixels_device           # several ranks: the bands gathered over NVLink
            rank, ws = parallel.world()
            if ws > 1:                           # each rank stacks its share of the halos, the small stacks are summed
                idx = np.array_split(idx, ws)[rank]
            d_lon = eng.to_device(self.catalog.data["lon"].values[idx], d_map.device)
            d_lat = eng.to_device(self.catalog.data["lat"].values[idx], d_map.device)
            d_stack = eng.stack_cutouts(d_map, self.nside, d_lon, d_lat, lonra, latra, xpix, ypix)
            if ws > 1:
                parallel.reduce_partial_map(d_stack, "all")
            stack = d_stack.cpu().numpy()
        else:
            stack = self._stack_cutouts_device(halo_list, lon_range, lat_range, xpix, ypix, apply_func, func_kwargs)
            if stack is None:       # some callable has no device twin: the reference's loop over cutouts
                stack = np.zeros((xpix, ypix))
                for cut_out in self.cutouts(halo_list, lon_range, lat_range, xpix, ypix, apply_func, func_kwargs):
                    stack += cut_out
        _say("Checkout the result with canvas.stack")
        if inplace:
            self.stack = stack
        else:
            return stack


    def _stack_cutouts_device(self, halo_list, lon_range, lat_range, xpix, ypix, apply_func, func_kwargs):
        """stack_cutouts with an apply_func chain of library operators: cutouts, operators and the sum all
        stay on the device; only the (ypix, xpix) stack comes back."""
        import copy
        import torch
        if not isinstance(apply_func, list):
            apply_func, func_kwargs = [apply_func], [func_kwargs]
        func_kwargs = [copy.copy(kw) for kw in func_kwargs]
        plan = self._patch_plan(apply_func, func_kwargs, xpix, ypix)
        if plan is None:
            return None
        idx, lonra, latra, xpix, ypix = self._cutout_setup(halo_list, lon_range, lat_range, xpix, ypix)
        d_map = self.pixels_device
        lon = self.catalog.data["lon"].values
        lat = self.catalog.data["lat"].values
        stack = torch.zeros((ypix, xpix), dtype=torch.float64, device=d_map.device)
        batch = max(1, self.CUTOUT_BATCH_BYTES // (16 * xpix * ypix))
        for b in range(0, len(idx), batch):
            rows = idx[b:b + batch]
            d_lon = eng.to_device(lon[rows], d_map.device)
            d_lat = eng.to_device(lat[rows], d_map.device)
            d_cuts = eng.cutouts(d_map, self.nside, d_lon, d_lat, lonra, latra, xpix, ypix)
            self._run_patch_plan(plan, d_cuts, rows, stack=stack)
        return stack.cpu().numpy()


#########################################################
#                   Painter Object
#########################################################
class Painter:
    """Sprays a template (radial profile) over a Canvas."""

    #: True: every device spray records CUDA events (chunk ends, uploads, read-back copies) in last_spray["trace"]
    trace = False

    def __init__(self, template):
        self.template = template

    @property
    def template(self):
        return self._template

    @template.setter
    def template(self, val):
        self._template = val
        self.R_arg_indx = self._analyze_template()

    # ------------------------ template introspection / marshalling ------------------------
    def _analyze_template(self):
        """Argument names of the template (paint_bucket.py:2542-2596)."""
        self.template_name = self.template.__name__
        spec = inspect.getfullargspec(self.template)
        self.template_args_list = [a for a in spec.args if a not in ("self", "cls")]
        self.template_kwargs_list = list(spec.kwonlyargs)
        message = (f"The template '{self.template_name}' takes in the following arguments:\n"
                   f"{self.template_args_list}\n")
        if len(self.template_kwargs_list) > 0:
            message += f"and the following keyword-only arguments:\n{self.template_kwargs_list}"
        assert sum([arg in self.template_args_list for arg in ['R', 'R_vec']]) == 1, \
            "Either 'R' or 'R_vec' must be a template argument (only one of them and not both)."
        R_arg_index = (self.template_args_list.index("R") if "R" in self.template_args_list
                       else self.template_args_list.index("R_vec"))
        _say(message)
        return R_arg_index

    def _shake_canister(self, catalog, template_kwargs):
        """Per-halo keyword arguments of the template as a dict of host columns: catalog columns
        named like template_args_list[1:] plus the user's kwargs, scalars (and length-1 values)
        broadcast to the catalog size (paint_bucket.py:2598-2676)."""
        cols = {}
        missing = []
        self._catalog_cols = set()
        for name in self.template_args_list[1:]:
            if name in catalog.data.columns:
                cols[name] = catalog.data[name].values
                self._catalog_cols.add(name)
            else:
                missing.append(name)
        if missing:
            _say("The following parameters were not found in the canvas.catalog.data\n"
                 f"{missing}\nMake sure you pass them as kwargs (key=value) in the .spray method.")
        n = catalog.size
        for key, value in template_kwargs.items():
            self._catalog_cols.discard(key)
            if not hasattr(value, "__len__"):
                cols[key] = np.full(n, value)
            else:
                value = np.asarray(value)
                if len(value) == 1:
                    cols[key] = np.tile(value, n)
                elif len(value) == n:
                    cols[key] = value
                else:
                    raise ValueError(f"template kwarg '{key}' has length {len(value)}; expected 1 or {n}")
        return cols

    # ------------------------ spray ------------------------
    def spray(self, canvas, distance_units="Mpc", n_cpus=1, cache=False, lazy=False, lazy_grid=None,
              **template_kwargs):
        """Paint every halo of canvas.catalog onto canvas.pixels (additively) with the template.

        Same contract as the reference (paint_bucket.py:2296-2495).  `n_cpus` only validates
        (0 raises ValueError): the halos are painted by CUDA kernels, and when torch.distributed is
        initialised with more than one rank the call is collective: each rank owns one ring band of the
        map and paints the halos that touch it (parallel.py).  `cache` is accepted and ignored; `lazy`
        snaps the template's per-halo arguments to 500-point grids as the reference does.
        """
        _say("Painting the canvas...")
        canvas._cancel_prefetch()
        canvas.template_name = self.template.__name__
        assert distance_units.lower() in ["mpc", "mpcs", "megaparsecs", "megaparsec"], \
            "For now the distance unit has to be megaparsecs."
        R_mode = self.template_args_list[self.R_arg_indx]
        if cache:
            assert R_mode == "R", "cache method is not available for profiles that use 'R_vec'"
        if n_cpus == 0:
            raise ValueError(f"n_cpus = {n_cpus} is not valid. Please enter a value > 0 for n_cpus or -1 "
                             "to run on all cores.")
        host_cols = self._shake_canister(canvas.catalog, template_kwargs)
        if lazy is True and lazy_grid is None:
            for key, data in host_cols.items():
                grid = np.linspace(data.min(), data.max(), 500)
                host_cols[key] = grid[np.searchsorted(grid, data)]
            self._catalog_cols = set()
        with eng.disc_mode(canvas.inclusive):
            if self._device_info(template_kwargs) is not None:
                stats = self._spray_banded(canvas, host_cols, template_kwargs)
            else:
                stats = self._spray_host_template(canvas, R_mode, host_cols, template_kwargs)
        canvas._mark_painted()
        self.last_spray = stats
        _say("Your artwork is finished. Check it out with Canvas.show_map()")

    def _device_info(self, template_kwargs):
        """("profile" | "table", registry entry) when the template has a CUDA twin that covers this call: every
        spray kwarg must be one of the twin's parameters -- anything else (e.g. NFW.deflection_angle's
        `suppress=False`) is a feature only the Python template has, so the call takes the host-template path,
        which forwards all kwargs exactly like the reference's loop (paint_bucket.py:2386-2389)."""
        tmpl = self.template
        if self.R_arg_indx != 0:
            return None
        info = getattr(tmpl, "_ap_device", None)
        if info is not None:
            known = set(info["args"]) | set(info["kwonly"])
            return ("profile", info) if set(template_kwargs) <= known else None
        info = getattr(tmpl, "_ap_device_table", None)
        if info is not None:
            known = set(info["args"])
            if info["scale"] is not None:
                known |= set(info["scale"]["args"]) | set(info["scale"]["kwonly"])
            return ("table", info) if set(template_kwargs) <= known else None
        return None

    # ------------------------ device templates: ring-band ownership ------------------------
    #: catalogs with at least this many halos per rank stream the finished rings of the map to the host while the
    #: remaining halos are still being painted (only once the host copy has been read before)
    PREFETCH_MIN_HALOS = 200000

    def _spray_banded(self, canvas, host_cols, template_kwargs, timers=None):
        """Built-in templates (device functions and device-tabulated @interpolate(@LOS_integrate) profiles): this
        rank paints the halos whose disc touches ITS ring band of the map (parallel.py, _bands.py), in
        colatitude order, chunk by chunk; finished rings go device -> host on a copy stream meanwhile."""
        import torch
        from . import _bands
        kind, info = self._device_info(template_kwargs)
        band = canvas._band()
        cat = canvas.catalog
        nside = canvas.nside
        order = cat.theta_order()[0]
        # the inclusive query widens every disc by about one pixel radius
        reach = cat.disc_reach(canvas.R_times, 4.0 / nside if canvas.inclusive else 0.0)
        d_band = canvas._device_band()
        dev = d_band.device
        # the decision to stream must be the same on every rank: it depends on the catalog size only
        host = None
        if cat.size // band.world_size >= self.PREFETCH_MIN_HALOS and not canvas.deterministic:
            host = canvas._reusable_host()
        plan = _bands.Plan(band, reach, n_chunks=None if host is not None else 1)
        lo, hi = plan.halo_lo, plan.halo_hi
        rows = order[lo:hi]
        needed = list(eng.GEOMETRY_COLUMNS) + ["R_ang_200c"] + [a for a in self._catalog_cols]
        pinned = cat.staged(needed)
        # a page-locked catalog painted in several chunks is uploaded chunk by chunk on its own stream
        piecewise = bool(pinned) and len(plan.chunks) > 1
        halos = eng.DeviceHalos(cat.data, canvas.R_times, dev, rows=rows, pinned=pinned,
                                pinned_range=(lo, hi) if pinned else None, lazy=piecewise)
        d_count = torch.zeros(1, dtype=torch.int64, device=dev)
        trace = {"main": [], "upload": [], "copy": []} if self.trace else None

        def mark(lane, name, stream=None):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream) if stream is not None else ev.record()
                trace[lane].append((name, ev))

        mark("main", "start")
        arrived = None
        if piecewise:
            wanted = list(info["args"]) + list(info.get("kwonly", {}))
            if info.get("scale"):
                wanted += list(info["scale"]["args"]) + list(info["scale"]["kwonly"])
            wanted = [a for a in wanted if a in host_cols and a in self._catalog_cols]
            up_stream = torch.cuda.Stream(device=dev)
            arrived = halos.upload_pieces(wanted, [(a - lo, e - lo) for (a, e, _) in plan.chunks], up_stream)
            mark("upload", "all uploads", up_stream)

        def col(name, defaults=None):
            """device tensor of a template argument: spray kwarg > catalog column > kw-only default"""
            if name in host_cols:
                if name in self._catalog_cols:
                    return halos.col(name)   # unmodified catalog column: pinned slice / gathered upload
                return eng.to_device(np.asarray(host_cols[name])[rows], dev)
            if defaults and name in defaults:
                return torch.full((1,), float(defaults[name]), dtype=torch.float64, device=dev)
            raise TypeError(f"{self.template_name}() missing required argument: '{name}'")

        def cut(t, a, e):
            return t if t.numel() == 1 else t[a:e]

        def ready(a, e, k):
            if arrived is not None:       # this chunk's columns: wait for their DMA, then form the disc radii
                torch.cuda.current_stream().wait_event(arrived[k])
                halos.radius_piece(a, e)

        if kind == "profile":
            params = [col(a) for a in info["args"]] + [col(k, info["kwonly"]) for k in info["kwonly"]]

            def paint_chunk(a, e, k):
                ready(a, e, k)
                eng.paint_profile(halos.view(a, e), nside, info["id"], [cut(t, a, e) for t in params], d_band, d_count)
                mark("main", f"chunk {k}")
            path = "device_profile"
        else:
            assert info["kind"] in eng.LOS_KINDS
            args = [col(a) for a in info["args"]]
            args = [t if t.numel() == halos.n else t.expand(halos.n).contiguous() for t in args]
            sc = info["scale"]
            if sc is not None:
                sc_cols = {a: col(a) for a in sc["args"]}
                sc_kw = {k: col(k, sc["kwonly"]) for k in sc["kwonly"]}

            def paint_chunk(a, e, k):
                ready(a, e, k)
                scale = None
                if sc is not None:     # per-halo multiplier (kSZ: -v_r / c T_cmb (1 + z)), formed chunk by chunk
                    scale = sc["fn"]({n_: cut(t, a, e) for n_, t in sc_cols.items()}, sc_kw)
                    scale = (scale if scale.numel() == e - a else scale.expand(e - a)).contiguous()
                eng.table_chunk(halos.view(a, e), nside, info["kind"], [t[a:e] for t in args], scale, info["n_samples"],
                                info["sampling_method"], d_band, d_count, timers)
                mark("main", f"chunk {k}")
            path = "device_table"

        chunk_done = None
        if host is not None:
            copy_stream = torch.cuda.Stream(device=dev)

            def chunk_done(rel_lo, rel_hi):
                ev = torch.cuda.Event()
                ev.record()
                copy_stream.wait_event(ev)
                eng.copy_d2h(host.band_view(rel_lo, rel_hi), d_band[rel_lo:rel_hi], stream=copy_stream)
                mark("copy", f"pixels {rel_lo}:{rel_hi}", copy_stream)
            path += "+stream"
        if canvas.deterministic:
            # sort-then-segmented-reduce: pairs for at most every pixel of every disc of the slice (K1 counts)
            with eng.band_clip(band):
                n_max = int(eng.disc_count(halos, nside)[0].sum().item()) if halos.n else 0
            labels = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int64)).to(dev)
            with eng.emit_pairs(n_max, labels, canvas.npix, dev) as pairs:
                eng.run_chunks(plan, paint_chunk, None)
            pairs.accumulate(d_band, band.pix_lo)
            path += "+deterministic"
        else:
            eng.run_chunks(plan, paint_chunk, chunk_done)
        if host is not None:
            canvas._prefetch = (copy_stream, band.npix)
        canvas._dfull = None
        return {"path": path, "n_halos": halos.n, "d_count": d_count, "world_size": band.world_size,
                "chunks": len(plan.chunks), "band": (band.ring_lo, band.ring_hi), "trace": trace}

    # ------------------------ host templates: the reference's halo batches ------------------------
    def _spray_host_template(self, canvas, R_mode, host_cols, template_kwargs):
        """Arbitrary Python templates (and @interpolate-wrapped Python profiles): the template runs on the host
        once per halo as in the reference; disc query, geometry, spline and scatter run on the device.  With
        several ranks each paints its np.array_split block of the catalog (paint_bucket.py:2414) into a full-sky
        partial map; the partial maps are summed over NVLink and every rank keeps its own band of the sum."""
        import torch
        from . import parallel
        if canvas.deterministic:
            raise NotImplementedError("Canvas(deterministic=True) is available for templates with a device twin; "
                                      f"'{self.template_name}' runs on the host-template path (atomic scatter)")
        band = canvas._band()
        d_band = canvas._device_band()
        if band.whole:
            stats = self._paint_rows(canvas, None, R_mode, host_cols, d_band)
        else:
            rows = parallel.shard_rows(canvas.catalog.size, band.rank, band.world_size)
            partial = torch.zeros(canvas.npix, dtype=d_band.dtype, device=d_band.device)
            stats = self._paint_rows(canvas, rows, R_mode, host_cols, partial)
            parallel.reduce_partial_map(partial, "all")
            d_band += partial[band.pix_lo:band.pix_hi]
            del partial
        canvas._dfull = None
        stats["world_size"] = band.world_size
        return stats

    def _paint_rows(self, canvas, rows, R_mode, host_cols, d_map):
        """Paint catalog rows `rows` (None = all) into the full-size device map d_map; returns stats dict."""
        import torch
        dev = d_map.device
        halos = eng.DeviceHalos(canvas.catalog.data, canvas.R_times, dev, rows=rows)
        nside = canvas.nside
        d_count = torch.zeros(1, dtype=torch.int64, device=dev)
        if self._interp_info() is not None:
            self._paint_python_table(self._interp_info(), halos, nside, rows, R_mode, host_cols, d_map, d_count)
            path = "python_table"
        else:
            hc = host_cols if rows is None else {k: v[rows] for k, v in host_cols.items()}
            eng.spray_flat(halos, nside, self.template, R_mode, hc, d_map, d_count)
            path = "flat"
        return {"path": path, "n_halos": halos.n, "d_count": d_count}

    def _interp_info(self):
        info = getattr(self.template, "_ap_interp", None)
        if info is None or self.R_arg_indx != 0:
            return None
        if info["min_frac"] or info["k"] != 3 or not info["default_interpolator"]:
            return None  # per-halo node counts / custom interpolators: flat path keeps exact semantics
        if info["sampling_method"] not in ("linspace", "logspace") or not (4 <= info["n_samples"] <= 32):
            return None
        return info

    def _paint_python_table(self, info, halos, nside, rows, R_mode, host_cols, d_map, d_count):
        """A Python template wrapped in @interpolate: nodes, spline and painting on device; the
        wrapped profile is evaluated on the host at each halo's nodes exactly as the reference's
        `interpolate` does (lib/utils.py:515-516).  Halos with fewer pixels than n_samples take the
        flat path (the decorator's direct-evaluation branch, :510-511)."""
        ns = info["n_samples"]
        inner = info["inner"]
        n_pix, _, rmin, rmax = eng.disc_count(halos, nside, want_range=True)
        nodes, n_nodes = eng.table_nodes(halos, n_pix, rmin, rmax, ns, info["sampling_method"])
        nodes_h = nodes.cpu().numpy()
        counts = n_pix.cpu().numpy()
        hc = host_cols if rows is None else {k: v[rows] for k, v in host_cols.items()}
        pos = [a for a in self.template_args_list[1:]]
        kwo = [k for k in self.template_kwargs_list if k in hc]
        values = np.zeros_like(nodes_h)
        for h in np.nonzero(counts >= ns)[0]:
            rest = [hc[a][h] for a in pos]
            kws = {k: hc[k][h] for k in kwo}
            values[h] = [inner(x, *rest, **kws) for x in nodes_h[h]]
        coef = eng.spline_build(nodes, eng.to_device(values.ravel(), halos.device).view(halos.n, ns), n_nodes,
                                uniform=(info["sampling_method"] == "linspace"))
        eng.paint_table(halos, nside, nodes, coef, n_nodes, None, d_map, d_count,
                        uniform=(info["sampling_method"] == "linspace"))
        small = np.nonzero((counts > 0) & (counts < ns))[0]
        if len(small):
            sub = halos.subset(small)
            eng.spray_flat(sub, nside, self.template, R_mode, {k: v[small] for k, v in hc.items()}, d_map, d_count)

    # ------------------------ 1-D template helpers ------------------------
    def calculate_template(self, R, catalog, halo_list=[0, ], **template_kwargs):
        """1-D profile of the template as a function of R [Mpc] for the given halos (host)."""
        assert hasattr(halo_list, "__iter__"), "halo_list must be an iterable"
        cols = self._shake_canister(catalog, template_kwargs)
        out = []
        for halo in halo_list:
            out.append(self.template(**{"R": R, **{k: v[halo] for k, v in cols.items()}}))
        return np.array(out)
