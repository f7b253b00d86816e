"""HEALPix map files: what `healpy.write_map` / `healpy.read_map` do for `Canvas.save_map_to_file` /
`load_map_from_file` (reference: astropaint/paint_bucket.py:1478-1558), without healpy or astropy.

The file layout is healpy 1.13's default: an empty primary HDU, then ONE binary-table extension holding the map as
a single column `TEMPERATURE` of big-endian float32 (healpy's `dtype=np.float32` default -- the reference calls
`hp.write_map(filename, self.pixels, overwrite=overwrite)`, so its files are single precision), 1024 pixels per table
row (`fits_IDL=True`), with the HEALPix keywords `PIXTYPE, ORDERING='RING', NSIDE, FIRSTPIX, LASTPIX, INDXSCHM,
OBJECT`.  Restated from the FITS standard and healpy's documented header; healpy is not installable here, so files
written by this module are checked by reading them back (and against the standard's block / card structure), not
against healpy itself.
"""
import os

import numpy as np

_BLOCK = 2880


def _card(key, value=None, comment=""):
    if value is None:
        return f"{key:<80}"[:80]
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20}"
    elif isinstance(value, float):
        v = f"{value:>20.14G}"
    else:
        v = f"'{str(value):<8}'"
        v = f"{v:<20}"
    text = f"{key:<8}= {v}"
    if comment:
        text += f" / {comment}"
    return f"{text:<80}"[:80]


def _header(cards):
    text = "".join(cards) + f"{'END':<80}"
    pad = (-len(text)) % _BLOCK
    return (text + " " * pad).encode("ascii")


def write_map(filename, m, overwrite=False, dtype=np.float32, column_name="TEMPERATURE", unit=""):
    """healpy.write_map(filename, m, nest=False, dtype=np.float32, fits_IDL=True, overwrite=overwrite)"""
    m = np.asarray(m)
    npix = m.size
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    if os.path.exists(filename) and not overwrite:
        raise OSError(f"File '{filename}' already exists.")       # astropy's message through healpy
    dtype = np.dtype(dtype)
    code = {"float32": "E", "float64": "D"}[dtype.name]
    per_row = 1024 if (npix > 1024 and npix % 1024 == 0) else 1
    rows = npix // per_row
    primary = _header([_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", 8, "array data type"),
                       _card("NAXIS", 0, "number of array dimensions"), _card("EXTEND", True)])
    ext = _header([
        _card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8, "array data type"),
        _card("NAXIS", 2, "number of array dimensions"), _card("NAXIS1", per_row * dtype.itemsize, "length of dimension 1"),
        _card("NAXIS2", rows, "length of dimension 2"), _card("PCOUNT", 0, "number of group parameters"),
        _card("GCOUNT", 1, "number of groups"), _card("TFIELDS", 1, "number of table fields"),
        _card("TTYPE1", column_name), _card("TFORM1", f"{per_row}{code}"), _card("TUNIT1", unit),
        _card("PIXTYPE", "HEALPIX", "HEALPIX pixelisation"),
        _card("ORDERING", "RING", "Pixel ordering scheme, either RING or NESTED"),
        _card("EXTNAME", "xtension", "name of this binary table extension"),
        _card("NSIDE", nside, "Resolution parameter of HEALPIX"), _card("FIRSTPIX", 0, "First pixel # (0 based)"),
        _card("LASTPIX", npix - 1, "Last pixel # (0 based)"),
        _card("INDXSCHM", "IMPLICIT", "Indexing: IMPLICIT or EXPLICIT"),
        _card("OBJECT", "FULLSKY", "Sky coverage, either FULLSKY or PARTIAL")])
    with open(filename, "wb") as f:
        f.write(primary)
        f.write(ext)
        chunk = 1 << 24
        for a in range(0, npix, chunk):      # big-endian, chunked: an nside 8192 map is 6.4 GB in fp64
            f.write(np.ascontiguousarray(m[a:a + chunk], dtype=dtype.newbyteorder(">")).tobytes())
        f.write(b"\0" * ((-npix * dtype.itemsize) % _BLOCK))


def _read_header(f):
    cards = {}
    while True:
        block = f.read(_BLOCK)
        if len(block) < _BLOCK:
            raise OSError("truncated FITS header")
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return cards
            if card[8:10] == "= ":
                val = card[10:].split(" /")[0].strip()
                if val.startswith("'"):
                    val = val[1:val.rindex("'")].strip() if val.count("'") >= 2 else val.strip("'").strip()
                cards[key] = val


def read_map(filename, dtype=np.float64):
    """healpy.read_map(filename): field 0 of the first binary-table extension as a RING-ordered array"""
    with open(filename, "rb") as f:
        primary = _read_header(f)
        naxis = int(primary.get("NAXIS", 0))
        size = abs(int(primary.get("BITPIX", 8))) // 8
        for k in range(1, naxis + 1):
            size *= int(primary[f"NAXIS{k}"])
        if naxis:
            f.seek((size + _BLOCK - 1) // _BLOCK * _BLOCK, os.SEEK_CUR)
        h = _read_header(f)
        if h.get("XTENSION") != "BINTABLE":
            raise OSError("no binary table extension found")
        if h.get("ORDERING", "RING").upper() != "RING":
            raise NotImplementedError("only RING-ordered map files are supported")
        tform = h["TFORM1"]
        code = tform[-1]
        per_row = int(tform[:-1] or 1)
        ftype = {"E": ">f4", "D": ">f8"}[code]
        rows = int(h["NAXIS2"])
        row_bytes = int(h["NAXIS1"])
        if int(h.get("TFIELDS", 1)) != 1 or row_bytes != per_row * np.dtype(ftype).itemsize:
            raise NotImplementedError("only single-column HEALPix map files are supported")
        data = np.fromfile(f, dtype=ftype, count=rows * per_row)
    if data.size != rows * per_row:
        raise OSError("truncated FITS data")
    return data.astype(dtype)
