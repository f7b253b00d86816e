"""Known-answer vectors from the published docstring examples of healpy (the third-party package that holds
the reference's pixel arithmetic: healpy==1.13.0, requirements.txt:15).  healpy itself is NOT installed here and
there is no network: the examples below are restated FROM MEMORY of the public API documentation of
healpy.pixelfunc (ang2pix, pix2ang, pix2vec, vec2pix, ring2nest, max_pixrad, nside2resol, nside2pixarea,
ud_grade) -- they could not be re-fetched here.  Their evidential weight comes from agreement: the oracle
(written independently from the HEALPix paper's ring-scheme formulas) reproduces all 28 integers and 40 floats,
which neither a mis-remembered example nor a wrong restatement would do by chance.  They pin the RING pixel numbering, the pixel-centre convention, the RING<->NEST map that ud_grade
goes through and ud_grade's averaging.  query_disc has no numeric docstring example; its two known answers at the end
of this file come from healpy's own TEST SUITE instead (see there).
Floats are quoted as printed (8 decimals unless the docstring shows repr precision).
"""
import numpy as np

PI = np.pi
ANG2PIX_16 = dict(theta=[PI / 2, PI / 4, PI / 2, 0, PI], phi=[0., PI / 4, PI / 2, 0, 0],
                  pix=[1440, 427, 1520, 0, 3068])
PIX2ANG_16 = dict(pix=[1440, 427, 1520, 0, 3068],
                  theta=[1.52911759, 0.78550497, 1.57079633, 0.05103658, 3.09055608],
                  phi=[0., 0.78539816, 1.61988371, 0.78539816, 0.78539816])
PIX2ANG_16_1440 = (1.5291175943723188, 0.0)                       # hp.pix2ang(16, 1440)
PIX2ANG_PIX11 = dict(nside=[1, 2, 4, 8],                           # hp.pix2ang([1, 2, 4, 8], 11)
                     theta=[2.30052398, 0.84106867, 0.41113786, 0.2044802],
                     phi=[5.49778714, 5.89048623, 5.89048623, 5.89048623])
PIX2VEC_16_1504 = (0.99879545620517241, 0.049067674327418015, 0.0)  # hp.pix2vec(16, 1504)
PIX2VEC_16 = dict(pix=[1440, 427], x=[0.99913157, 0.5000534], y=[0., 0.5000534], z=[0.04166667, 0.70703125])
PIX2VEC_PIX11 = dict(nside=[1, 2], x=[0.52704628, 0.68861915], y=[-0.52704628, -0.28523539],
                     z=[-0.66666667, 0.66666667])                  # hp.pix2vec([1, 2], 11)
VEC2PIX_16 = dict(x=[1, 0], y=[0, 1], z=[0, 0], pix=[1504, 1520])
RING2NEST_16_1504 = 1130
RING2NEST_2 = [3, 7, 11, 15, 2, 1, 6, 5, 10, 9]                    # hp.ring2nest(2, np.arange(10))
RING2NEST_PIX11 = dict(nside=[1, 2, 4, 8], nest=[11, 13, 61, 253])  # hp.ring2nest([1, 2, 4, 8], 11)
NSIDE2RESOL_128_ARCMIN = 27.483891294539248
NSIDE2PIXAREA_128 = 6.391586616190171e-05
MAX_PIXRAD_16 = 0.0660147614                                        # '%.10f' % hp.max_pixrad(16)
UD_GRADE_48_TO_1 = [5.5, 7.25, 9., 10.75, 21.75, 21.75, 23.75, 25.75, 36.5, 38.25, 40., 41.75]  # hp.ud_grade(np.arange(48.), 1)


# ---- hp.query_disc: the two known answers of healpy's own test-suite, healpy/test/test_query_disc.py, class
# TestQueryDisc (setUp: NSIDE = 8, vec = [0.17101007, 0.03015369, 0.98480775], radius = np.radians(6)):
#   test_not_inclusive:  query_disc(NSIDE, vec, radius, inclusive=False) == [4]
#   test_inclusive:      query_disc(NSIDE, vec, radius, inclusive=True)  == [0, 3, 4, 5, 11, 12, 13, 23]
# (both annotated there with the output of the HEALPix IDL routine `query_disc, 8, vec, 6, listpix, /DEG, NESTED=0
# [, /inclusive]`).  Restated FROM MEMORY like everything in this file (round 2): the oracle, written before these
# vectors were recalled, reproduces the single pixel and the eight-pixel list exactly -- the only anchor of the disc
# query (exclusive AND inclusive, fact = 4) to something the real healpy asserts.
QUERY_DISC_TEST = dict(nside=8, vec=[0.17101007, 0.03015369, 0.98480775], radius_deg=6.0,
                       exclusive=[4], inclusive=[0, 3, 4, 5, 11, 12, 13, 23])
# healpy.rotator.angdist docstring: hp.rotator.angdist([.2, 0], [.2, 1e-6]) -> array([1.98669331e-07])
ANGDIST_EXAMPLE = dict(dir1=[0.2, 0.0], dir2=[0.2, 1e-6], value=1.98669331e-07)
