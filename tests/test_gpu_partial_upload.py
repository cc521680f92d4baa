"""Row-range transfers of ``p2p_process_image``: the library copies only the panorama rows the requested views
can touch.  The range must be exactly the oracle's (min / max tap row of the reference pitch maps, ref :114-175 +
cv2's 1/32-px quantiser) and the pixels must not change."""
import numpy as np
import pytest

from oracle import ref_port, synth

pytestmark = pytest.mark.gpu

GEOMS = [  # Wp, Hp, W, H, fov, pitches
    (1024, 512, 240, 136, 120, [30, 60, 90]),       # C2 scaled: top rows incl. the pole, bottom quarter untouched
    (1024, 512, 256, 256, 90, [90]),                # one horizon face: middle band only
    (1024, 512, 256, 256, 90, [0]),                 # north pole face
    (1024, 512, 256, 256, 90, [180]),               # south pole face: reaches the last row (clamp row needed)
    (1000, 500, 328, 200, 100, [45, 135, 150]),
    (1023, 77, 64, 40, 60, [80, 100]),
    (2048, 1024, 640, 480, 90, [90]),               # C1
]


def oracle_rows(W, H, fov, pitches, Wp, Hp):
    lo, hi = None, None
    for p in pitches:
        _, V = ref_port.pitch_mapping(W, H, fov, p, Wp, Hp)
        ok = ~np.isnan(V)
        iy = (np.rint(V[ok] * np.float32(32)).astype(np.int64)) >> 5
        lo = int(iy.min()) if lo is None else min(lo, int(iy.min()))
        hi = int(iy.max()) if hi is None else max(hi, int(iy.max()))
    return lo, min(hi + 1, Hp - 1)


@pytest.mark.parametrize("Wp,Hp,W,H,fov,pitches", GEOMS)
def test_row_range_equals_oracle(pkg, proj, Wp, Hp, W, H, fov, pitches):
    from oracle import svml_model

    if not svml_model.host_numpy_uses_svml():
        pytest.skip("bit comparison with NumPy's arccos needs the AVX-512 (SVML) code path on this host")
    consts = [pkg.pitch_constants(W, fov, p) for p in pitches]
    assert proj.view_row_range(consts, W, H, Wp, Hp) == oracle_rows(W, H, fov, pitches, Wp, Hp)


@pytest.mark.parametrize("Wp,Hp,W,H,fov,pitches", GEOMS)
def test_partial_upload_same_pixels(pkg, proj, Wp, Hp, W, H, fov, pitches):
    L = pkg._lib
    pano = synth.noise(Wp, Hp, 5)
    yaws = [0, 90, 180, 270]
    shifts = [pkg.yaw_table(Wp, y)[2] for y in yaws]
    if any(s is None for s in shifts):
        shifts = [0, Wp // 4, Wp // 2, (3 * Wp) // 4]
    consts = [pkg.pitch_constants(W, fov, p) for p in pitches]
    outs = {}
    for partial in (1, 0):
        proj.set_option(L.OPT_PARTIAL_UPLOAD, partial)
        out = np.empty((len(shifts), len(pitches), H, W, 3), np.uint8)
        with proj.slots(1) as (s,):
            # poison the slot first so stale rows of an earlier full upload can not hide a missing row
            proj.upload(s, np.full((Hp, Wp, 3), 255, np.uint8))
            proj.process_image(s, pano, shifts, consts, W, H, out)
            proj.sync(s)
        outs[partial] = out
    proj.set_option(L.OPT_PARTIAL_UPLOAD, 1)
    assert np.array_equal(outs[1], outs[0])
    first, last = proj.view_row_range(consts, W, H, Wp, Hp)
    assert 0 <= first <= last <= Hp - 1


def test_partial_slot_refuses_other_views(pkg, proj):
    Wp, Hp, W, H, fov = 1024, 512, 256, 256, 90
    pano = synth.noise(Wp, Hp, 6)
    horizon = [pkg.pitch_constants(W, fov, 90)]
    pole = [pkg.pitch_constants(W, fov, 0)]
    out = np.empty((1, 1, H, W, 3), np.uint8)
    with proj.slots(1) as (s,):
        proj.process_image(s, pano, [0], horizon, W, H, out)
        proj.sync(s)
        again = proj.project(s, [128], horizon, W, H)      # same geometry, other yaw: rows are there
        proj.sync(s)
        assert again.shape == (1, 1, H, W, 3)
        with pytest.raises(pkg.P2PError) as ei:
            proj.project(s, [0], pole, W, H)                # needs rows the partial upload never moved
        assert ei.value.code == -4
        with pytest.raises(pkg.P2PError):
            proj.download_pano(s, Wp, Hp)
        proj.upload(s, pano)                                # a full upload makes every view legal again
        proj.project(s, [0], pole, W, H)
        proj.sync(s)
