"""Ring-band ownership of the map and the host-side schedule of one spray.

The reference's parallel spray has joblib threads adding into ONE shared array
(astropaint/paint_bucket.py:2414-2433).  On several GPUs the B200 form of that is not "every GPU paints a
full-sky partial map, then reduce" (6.44 GB per rank at nside 8192) but OWNERSHIP: rank g owns the rings
[ring_lo, ring_hi) of the map -- one contiguous pixel range in RING order, an equal share of the sphere's
area -- keeps only that band in HBM and paints every halo whose disc touches it, clipped to its own rings
(ap_band_clip).  A disc that straddles a boundary is painted by both neighbours, each keeping its own
rings, so the bands tile the map exactly and no partial map, zero-fill, reduction or `base += partial` exists.

Halos are taken in order of colatitude (Catalog.theta_order): the halos that can touch a band are one
contiguous slice of that order, found on the host by two binary searches, so nothing is sorted, bucketed or
compacted per spray and the host never waits for the device.  Within a rank the slice is cut into chunks;
after a chunk, every ring farther north than (next halo's colatitude - largest disc radius) is final and can
be shipped device->host on a copy stream while the next chunk is in its QUADPACK phase.

`SharedHostMap` is the page-locked host copy of the whole map that all ranks of the node write their bands
into (one PCIe link per GPU instead of one link for the whole map).
"""
import math
import mmap
import os
import sys

import numpy as np


# ---------------------------------------------------------------------------------------------
# ring arithmetic on the host (healpix ring scheme; the kernels' own versions live in csrc/healpix.cuh)
# ---------------------------------------------------------------------------------------------
def ring_start(nside, ring):
    """first pixel of 1-based `ring` (healpix get_ring_info_small); ring == 4 nside gives npix"""
    if ring >= 4 * nside:
        return 12 * nside * nside
    if ring < nside:
        return 2 * ring * (ring - 1)
    if ring < 3 * nside:
        return 2 * nside * (nside - 1) + (ring - nside) * 4 * nside
    nr = 4 * nside - ring
    return 12 * nside * nside - 2 * nr * (nr + 1)


def ring_z(nside, ring):
    """z = cos(colatitude) of the pixel centres of `ring` (healpix ring2z)"""
    if ring < nside:
        return 1.0 - ring * ring * (4.0 / (12 * nside * nside))
    if ring <= 3 * nside:
        return (2 * nside - ring) * (2.0 / (3 * nside))
    r = 4 * nside - ring
    return r * r * (4.0 / (12 * nside * nside)) - 1.0


def ring_above(nside, z):
    """healpix ring_above: the ring whose centres lie at or just north of height z (0 .. 4 nside - 1)"""
    az = abs(z)
    if az <= 2.0 / 3.0:
        return int(nside * (2.0 - 1.5 * z))
    ir = int(nside * math.sqrt(3.0 * (1.0 - az)))
    return ir if z > 0 else 4 * nside - ir - 1


def owner_rings(nside, world_size, weights=None):
    """First ring of every rank's band plus the end marker 4 nside, aligned to ring boundaries: shares of the
    sphere's area proportional to `weights` (z = 1 - 2 F_g, F the cumulative share), equal shares when None.
    Depends on (nside, world_size, weights) only, so every rank and every spray agrees on it."""
    if weights is None:
        cum = [(g + 1.0) / world_size for g in range(world_size)]
    else:
        w = np.asarray(weights, dtype=np.float64)
        assert len(w) == world_size and np.all(w > 0), "one positive weight per rank"
        cum = list(np.cumsum(w / w.sum()))
    rings = [1]
    for g in range(1, world_size):
        r = ring_above(nside, 1.0 - 2.0 * cum[g - 1]) + 1
        rings.append(min(max(r, rings[-1]), 4 * nside))
    rings.append(4 * nside)
    return rings


class Band:
    """The part of the map one rank owns: rings [ring_lo, ring_hi) = pixels [pix_lo, pix_hi).  `weights` (one per
    rank, the same on every rank) sizes the bands, e.g. by each GPU's measured device->host bandwidth so that the
    read-back of the map finishes on all ranks together (parallel.readback_weights)."""

    def __init__(self, nside, rank, world_size, weights=None):
        self.nside, self.rank, self.world_size = int(nside), int(rank), int(world_size)
        self.weights = None if weights is None else tuple(float(x) for x in weights)
        self.rings = owner_rings(self.nside, self.world_size, self.weights)
        self.pix_bounds = [ring_start(self.nside, r) for r in self.rings]
        self.ring_lo, self.ring_hi = self.rings[rank], self.rings[rank + 1]
        self.pix_lo, self.pix_hi = self.pix_bounds[rank], self.pix_bounds[rank + 1]

    @property
    def npix(self):
        return self.pix_hi - self.pix_lo

    @property
    def whole(self):
        return self.world_size == 1


#: work (= halo count) ratio of consecutive chunks.  A chunk's device->host copy overlaps the next chunk's
#: compute, so what stays exposed at the end of a spray is the LAST chunk's copy: geometrically shrinking
#: chunks end on a small one, and a ratio >= t_copy / t_compute keeps every earlier copy hidden
#: (measured in round 1 on the 1e7-halo bench catalog: 10 chunks x 0.75 beat 8 equal chunks by 11 ms end to end).
CHUNK_RATIO = 0.75
MAX_CHUNKS = 10            # 12 with capped shares (several ranks)
MIN_CHUNK_HALOS = 50_000
#: Several ranks: the read-back is SLOWER than the painting (the host ingests ~110 GB/s from eight GPUs), so the copy
#: engine should start early and never idle: no chunk larger than this share -- the large chunks of the geometric series
#: are split into equal ones (8-GPU trace: a 25% chunk kept the copy engine idle for 7 of 70 ms).  With one rank the
#: painting is the slower side and the plain geometric series is 2 ms better end to end (fewer, larger launches).
MAX_CHUNK_SHARE = 0.12
#: share of the first chunk when there are three or more: the halo columns go host->device chunk by chunk on
#: their own stream, so painting starts as soon as this small lead chunk has arrived and every later upload hides
#: behind the previous chunk's kernels
LEAD_FRACTION = 0.04


class Plan:
    """Host-side schedule of one rank's share of a spray (no device work, no synchronisation).

    reach = (north, south) from Catalog.disc_reach: for the halos in colatitude order, north[i] is the smallest
    colatitude reached by any disc of halos i.. and south[i] the largest reached by any disc of halos ..i."""

    def __init__(self, band, reach, n_chunks=None, ratio=None):
        nside = band.nside
        self.band = band
        north, south = reach
        margin = 4.0 / nside               # more than two ring spacings: rounding of ring boundaries, centre from (x, y, z)
        n = len(north)
        if band.whole:
            lo, hi = 0, n
        else:
            # colatitudes of the band's first and last ring; a halo whose disc stays `margin` clear of them cannot touch it
            t_top = math.acos(max(-1.0, min(1.0, ring_z(nside, band.ring_lo))))
            t_bot = math.acos(max(-1.0, min(1.0, ring_z(nside, max(band.ring_hi - 1, band.ring_lo)))))
            lo = int(np.searchsorted(south, t_top - margin, side="left")) if band.rank > 0 else 0
            hi = int(np.searchsorted(north, t_bot + margin, side="right")) if band.rank < band.world_size - 1 else n
            hi = max(hi, lo)
            if band.npix == 0:
                hi = lo
        self.halo_lo, self.halo_hi = lo, hi
        m = hi - lo
        capped = not band.whole
        if n_chunks is None:
            n_chunks = max(1, min(MAX_CHUNKS + (2 if capped else 0), m // MIN_CHUNK_HALOS))
        ratio = CHUNK_RATIO if ratio is None else ratio
        if n_chunks >= 3 and not capped:
            f = np.power(float(ratio), np.arange(n_chunks - 1))
            f = np.concatenate([[LEAD_FRACTION], (1.0 - LEAD_FRACTION) * f / f.sum()])
        elif n_chunks >= 3:
            # lead chunk, equal body chunks of at most MAX_CHUNK_SHARE, then a tail shrinking by `ratio`
            n_tail = min(n_chunks - 2, 5)
            tail = MAX_CHUNK_SHARE * np.power(float(ratio), np.arange(1, n_tail + 1))
            n_body = n_chunks - 1 - n_tail
            body_total = 1.0 - LEAD_FRACTION - tail.sum()
            if body_total <= 0 or body_total / n_body > 2.5 * MAX_CHUNK_SHARE:     # too few chunks for that shape
                f = np.power(float(ratio), np.arange(n_chunks - 1))
                f = np.concatenate([[LEAD_FRACTION], (1.0 - LEAD_FRACTION) * f / f.sum()])
            else:
                f = np.concatenate([[LEAD_FRACTION], np.full(n_body, body_total / n_body), tail])
        else:
            f = np.power(float(ratio), np.arange(n_chunks))
        cuts = lo + np.round(np.cumsum(f / f.sum()) * m).astype(np.int64)
        cuts[-1] = hi
        #: (halo_lo, halo_hi, done_ring): after the chunk's kernels, rings [ring_lo, done_ring) are final
        self.chunks = []
        a = lo
        for e in cuts:
            e = int(max(e, a))
            if e >= hi:
                done = band.ring_hi
            else:
                top = float(north[e]) - margin        # no halo from e on reaches farther north than this
                done = ring_above(nside, math.cos(min(math.pi, top))) if top > 0.0 else band.ring_lo
                done = min(max(done, band.ring_lo), band.ring_hi)
            if self.chunks and done < self.chunks[-1][2]:
                done = self.chunks[-1][2]
            self.chunks.append((a, e, done))
            a = e
        if not self.chunks:
            self.chunks = [(lo, hi, band.ring_hi)]
        self.chunks[-1] = (self.chunks[-1][0], hi, band.ring_hi)


# ---------------------------------------------------------------------------------------------
# page-locked host copy of the whole map, shared by the ranks of one node
# ---------------------------------------------------------------------------------------------
_PAGE = mmap.PAGESIZE


class SharedHostMap:
    """fp64 / fp32 [npix] in host memory that every rank of the node can write.

    world_size == 1: a torch pinned tensor.  world_size > 1: an anonymous memory file (memfd_create on rank 0,
    opened by the other ranks through /proc/<pid>/fd/<fd>; not bounded by the size of the /dev/shm mount),
    mapped shared by every rank; each rank page-locks (cudaHostRegister) the pages of ITS band so that its
    device->host copies are asynchronous DMA.  Construction is collective."""

    def __init__(self, npix, np_dtype, band, register=True):
        import torch
        self.npix = int(npix)
        self.dtype = np.dtype(np_dtype)
        self.nbytes = self.npix * self.dtype.itemsize
        self.band = band
        self._registered = None
        self._mm = None
        self._fd = None
        if band.world_size == 1:
            tdt = torch.float32 if self.dtype == np.float32 else torch.float64
            self.tensor = torch.empty(self.npix, dtype=tdt, pin_memory=bool(register and torch.cuda.is_available()))
            self.array = self.tensor.numpy()
            self._base_refs = sys.getrefcount(self.array)
            return
        import torch.distributed as dist
        info = [None]
        if band.rank == 0:
            self._fd = os.memfd_create("astropaint_b200_map")
            os.ftruncate(self._fd, max(self.nbytes, 1))
            info[0] = (os.getpid(), self._fd)
        dist.broadcast_object_list(info, src=0)
        if band.rank != 0:
            pid, fd = info[0]
            self._fd = os.open(f"/proc/{pid}/fd/{fd}", os.O_RDWR)
        self._mm = mmap.mmap(self._fd, max(self.nbytes, 1), mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        self.array = np.frombuffer(self._mm, dtype=self.dtype, count=self.npix)
        self.tensor = torch.from_numpy(self.array)
        dist.barrier()        # every rank has the file open before rank 0 may ever drop it
        if register and torch.cuda.is_available() and band.npix > 0:
            base = self.array.ctypes.data
            lo = base + band.pix_lo * self.dtype.itemsize
            hi = base + band.pix_hi * self.dtype.itemsize
            lo_al = lo - (lo % _PAGE)
            hi_al = min(base + ((self.nbytes + _PAGE - 1) // _PAGE) * _PAGE, ((hi + _PAGE - 1) // _PAGE) * _PAGE)
            from . import _capi
            _capi.check(_capi.lib().ap_host_register(_capi.c_ptr(lo_al), hi_al - lo_al))
            self._registered = lo_al
        self._base_refs = sys.getrefcount(self.array)

    def handed_out(self):
        """True while anything outside this object (the user's variable, a view, a tensor made from it) still
        references the host array: such a buffer must never be written again after Canvas.clean()."""
        return sys.getrefcount(self.array) > self._base_refs

    def band_view(self, rel_lo, rel_hi):
        """torch view of pixels [pix_lo + rel_lo, pix_lo + rel_hi) of the rank's own band"""
        p = self.band.pix_lo
        return self.tensor[p + rel_lo:p + rel_hi]

    def close(self):
        if self._registered is not None:
            from . import _capi
            try:
                _capi.lib().ap_host_unregister(_capi.c_ptr(self._registered))
            except Exception:
                pass
            self._registered = None
        # the mapping itself stays alive as long as numpy views of it do (np.frombuffer holds the mmap)
        if self._fd is not None:
            try:
                os.close(self._fd)
            except OSError:
                pass
            self._fd = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
