"""CPU tests of the complete capteur of the oracle (output.f90:294-595): Monte Carlo photon maps of the image
step (loutput_mc), lonly_capt_interet and the lorigine tallies.  Run with -m "not gpu"."""
import numpy as np

from oracle.binding import Oracle

from helpers import small_problems

IMG = dict(letape_th=0, lmono=1, lmono0=1, loutput_mc=1, lambda_in=8, p_lambda_in=8, n_photons2=300, n_phot_lim=1.0e30,
           npix_x=32, npix_y=32, map_size=700.0, N_thet=4, N_phi=2)


def _P():
    return small_problems()["cyl2D"]()


def test_photon_map_conserves_the_detected_energy():
    P = _P()
    O = Oracle(P)
    t = O.run(n_threads=1, lsepar_pola=1, lsepar_contrib=1, **IMG)
    m = t.stokes_map
    assert m.shape == (32, 32, 4, 2, 8)
    assert t.stats[0] == 128 * 300 and t.stats[6] > 0
    # the map is wide enough for the whole model: every escaping packet lands in a pixel with its energy
    I = m[..., 0]
    assert I.sum() > 0
    # contributions add up to the total intensity, pixel by pixel
    assert np.allclose(m[..., 4:8].sum(axis=-1), I, rtol=1e-12, atol=1e-300)
    # unscattered star light stays in the four pixels around the map centre (the star sits on their common corner)
    star = m[..., 4]
    for it in range(4):
        for ip in range(2):
            nz = np.argwhere(star[:, :, it, ip] > 0)
            assert len(nz) <= 4 and all(15 <= a <= 16 and 15 <= b <= 16 for a, b in nz)
    # polarisation comes from scattered light only
    assert np.abs(m[..., 1]).sum() > 0
    # a narrow map loses the packets that fall outside (output.f90:430-435)
    small = O.run(n_threads=1, lsepar_pola=1, lsepar_contrib=1, **dict(IMG, map_size=100.0))
    assert small.stokes_map[..., 0].sum() < I.sum()


def test_left_right_symmetry_puts_half_a_photon_in_each_mirror_pixel():
    P = _P()
    O = Oracle(P)
    a = O.run(n_threads=1, lsepar_pola=1, **IMG)
    b = O.run(n_threads=1, lsepar_pola=1, l_sym_ima=1, **IMG)
    Ia, Ib = a.stokes_map[..., 0], b.stokes_map[..., 0]
    assert np.isclose(Ia.sum(), Ib.sum(), rtol=1e-12)
    # the symmetrised map is mirror-symmetric about the vertical axis of the image (pixel i <-> npix_x + 1 - i up
    # to the one-pixel shift of int()): compare coarse halves
    left, right = Ib[:16].sum(), Ib[16:].sum()
    assert abs(left - right) < 0.02 * Ib.sum()
    # U flips sign in the mirror pixel: the symmetrised U map integrates to ~0 where the plain one does not have to
    assert abs(b.stokes_map[..., 2].sum()) <= abs(np.abs(b.stokes_map[..., 2]).sum()) * 0.2 + 1e-12


def test_only_capt_interet_and_origin_tallies():
    P = _P()
    O = Oracle(P)
    kw = dict(letape_th=0, lmono=1, lambda_in=8, p_lambda_in=8, n_photons2=10 ** 9, n_phot_lim=100.0, N_thet=5)
    full = O.run(n_threads=1, lorigine=1, capt_interet=2, **kw)
    part = O.run(n_threads=1, lonly_capt_interet=1, capt_inf=2, capt_sup=3, **kw)
    # the packets are the same (chunks stop on n_phot_lim): bins 2..3 are kept, the others dropped
    assert np.array_equal(part.sed[:, 1:3], full.sed[:, 1:3])
    assert part.sed[:, 0].sum() == 0 and part.sed[:, 3:].sum() == 0 and full.sed[:, 0].sum() > 0
    # origin of the energy received in bin capt_interet: star + disk cells = the bin's energy
    got = full.star_origin[7] + full.disk_origin[7].sum()
    assert np.isclose(got, full.sed[7, 1].sum(), rtol=1e-12)
    assert full.star_origin[7] > 0


def test_packet_counts_per_cell():
    """xN_abs (radiation_field.f90:53,60): one count per crossing of a non-empty cell."""
    P = _P()
    O = Oracle(P)
    t = O.run(n_threads=1, n_photons2=50, lxN_abs=1)                      # thermal, library mode: one column
    assert t.xN_abs.shape == (P.n_cells, 1)
    assert 0 < t.xN_abs.sum() <= t.stats[1]
    assert np.array_equal(t.xN_abs[:, 0] > 0, t.xKJ_abs > 0)
    s = O.run(n_threads=1, xJ=True, letape_th=0, lmono=1, lambda_in=8, p_lambda_in=8, n_photons2=10 ** 9, n_phot_lim=50.0,
              lxJ_abs=1, lxN_abs=1)
    assert s.xN_abs.shape == (P.n_cells, P.n_lambda)
    assert s.xN_abs[:, 7].sum() > 0 and s.xN_abs.sum() == s.xN_abs[:, 7].sum()
    assert np.array_equal(s.xN_abs[:, 7] > 0, s.xJ_abs[:, 7] > 0)
