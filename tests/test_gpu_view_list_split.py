"""GPU tests of the flat view list (one launch for any (yaw, pitch) list, BASELINE configs[4] cube faces), output row
bands, panorama replication between contexts and the single-image / folder split over several GPUs.

The split paths are exercised on ONE device too: a device may be listed twice (``devices=[0, 0]``) - two slots of the
same context then play the two GPUs, the panorama is replicated by a device-to-device copy - so the logic is covered on a
single-GPU box; with >= 2 GPUs the same tests also run over distinct devices (peer copy over NVLink).
"""
import threading

import numpy as np
import pytest

from oracle import fixedpoint as fp
from oracle import svml_model, synth

pytestmark = pytest.mark.gpu

FACES = [(0, 90), (90, 90), (180, 90), (270, 90), (0, 0), (0, 180)]   # cube faces, pole pitches included


def _device_sets(pkg):
    n = pkg._lib.load().p2p_device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 4))))
    if n >= 8:
        sets.append(list(range(8)))
    return sets


def _flat(pkg, Wp, W, fov, views):
    shifts = [pkg.yaw_table(Wp, y)[2] for y, _ in views]
    consts = [pkg.pitch_constants(W, fov, p) for _, p in views]
    assert all(s is not None for s in shifts)
    return shifts, consts


# ------------------------------------------------------------------------------------------
# flat view list
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("noise", [True, False])
def test_cube_faces_one_launch_bit_exact_against_oracle(pkg, proj, noise):
    """The six cube faces (four yaws at pitch 90 + the two poles) are ONE kernel launch and equal the oracle bit for bit
    (the per-view product API needed three launches)."""
    Wp, Hp, W, H, fov = 1024, 512, 256, 256, 90
    pano = synth.noise(Wp, Hp, 5) if noise else synth.smooth(Wp, Hp, 5)
    shifts, consts = _flat(pkg, Wp, W, fov, FACES)
    with proj.slots(1) as (s,):
        proj.upload(s, pano)
        proj.sync(s)
        n0 = proj.launches
        out = proj.project_list(s, shifts, consts, W, H)
        proj.sync(s)
        assert proj.launches - n0 == 1
    strict = svml_model.host_numpy_uses_svml()   # else the reference itself differs in the last ulp on this host
    for i, (yaw, pitch) in enumerate(FACES):
        want = fp.project_view_single_pass(pano, yaw, pitch, W, H, fov)
        if strict:
            assert np.array_equal(out[i], want), (yaw, pitch)
        else:
            assert (out[i] == want).all(axis=-1).mean() >= 0.93, (yaw, pitch)


def test_view_list_equals_product_api_and_any_order(pkg, proj):
    """Any order, repeated views, more than four yaws per pitch, odd group sizes: same bytes as the yaw x pitch API."""
    Wp, Hp, W, H, fov = 2048, 1024, 320, 200, 110
    pano = synth.noise(Wp, Hp, 6)
    yaws, pitches = [0, 90, 180, 270, 45, 135], [30, 60, 90, 150]
    with proj.slots(1) as (s,):
        proj.upload(s, pano)
        ref = proj.project(s, [pkg.yaw_table(Wp, y)[2] for y in yaws], [pkg.pitch_constants(W, fov, p) for p in pitches],
                           W, H)
        proj.sync(s)
        rng = np.random.default_rng(0)
        views = [(yaws[k], pitches[j]) for k in range(len(yaws)) for j in range(len(pitches))]
        order = list(rng.permutation(len(views))) + [3, 3, 7]          # shuffled, with repeats
        shifts, consts = _flat(pkg, Wp, W, fov, [views[i] for i in order])
        out = proj.project_list(s, shifts, consts, W, H)
        proj.sync(s)
    for n, i in enumerate(order):
        k, j = divmod(i, len(pitches))
        assert np.array_equal(out[n], ref[k, j]), (n, i)


@pytest.mark.parametrize("W,H", [(256, 136), (250, 100), (8, 3), (64, 1)])
def test_row_bands_tile_the_view(pkg, proj, W, H):
    """Bands written by separate calls assemble the full views; rows outside a band are not touched.  W = 250 takes the
    generic kernels (whole views rendered, band cut by the copy)."""
    from p2p_b200 import shard

    Wp, Hp, fov = 1024, 512, 100
    pano = synth.noise(Wp, Hp, 7)
    views = [(0, 45), (90, 45), (180, 90), (0, 135), (270, 90)]
    shifts, consts = _flat(pkg, Wp, W, fov, views)
    with proj.slots(1) as (s,):
        proj.upload(s, pano)
        full = proj.project_list(s, shifts, consts, W, H)
        proj.sync(s)
        for world in (1, 2, 3, 5):
            out = np.full((len(views), H, W, 3), 0xAB, np.uint8)
            for r in range(world):
                lo, hi = shard.shard_rows(H, r, world)
                before = out.copy()
                proj.project_list(s, shifts, consts, W, H, rows=(lo, hi), out=out)
                proj.sync(s)
                mask = np.ones(H, bool)
                mask[lo:hi] = False
                assert np.array_equal(out[:, mask], before[:, mask])     # nothing outside the band
            assert np.array_equal(out, full), world
        with pytest.raises(pkg.P2PError) as e:
            proj.project_list(s, shifts, consts, W, H, rows=(0, H + 1))
        assert e.value.code == -1


def test_view_list_on_partial_slot_and_device_output(pkg, proj):
    """A slot filled by ``process_image`` (row-range upload) serves view lists its rows cover and refuses others."""
    torch = pytest.importorskip("torch")
    Wp, Hp, W, H, fov = 2048, 1024, 256, 144, 90
    pano = synth.noise(Wp, Hp, 8)
    consts = [pkg.pitch_constants(W, fov, 40)]
    shifts = [0, Wp // 4]
    with proj.slots(1) as (s,):
        out = np.empty((2, 1, H, W, 3), np.uint8)
        proj.process_image(s, pano, shifts, consts, W, H, out)
        proj.sync(s)
        again = proj.project_list(s, shifts, consts * 2, W, H)
        proj.sync(s)
        assert np.array_equal(again, out[:, 0])
        with pytest.raises(pkg.P2PError) as e:
            proj.project_list(s, [0], [pkg.pitch_constants(W, fov, 150)], W, H)
        assert e.value.code == -4
        # device-resident output
        d = torch.empty((2, H, W, 3), dtype=torch.uint8, device="cuda:0")
        proj.project_list(s, shifts, consts * 2, W, H, out_device_ptr=d.data_ptr())
        proj.sync(s)
        assert np.array_equal(d.cpu().numpy(), out[:, 0])


# ------------------------------------------------------------------------------------------
# replication between slots / contexts
# ------------------------------------------------------------------------------------------
def test_copy_pano_between_slots_and_contexts(pkg, proj):
    Wp, Hp = 1000, 300
    pano = synth.noise(Wp, Hp, 9)
    other = pkg.Projector(pkg._lib.load().p2p_device_count() - 1, n_slots=2)   # last device: a peer copy when N >= 2
    try:
        with proj.slots(2) as (a, b):
            proj.upload(a, pano)
            proj.copy_pano_from(b, proj, a)          # same context
            other.copy_pano_from(0, proj, a)         # another context (another device if there is one)
            proj.sync(b)
            other.sync(0)
            assert np.array_equal(proj.download_pano(b, Wp, Hp), pano)
            assert np.array_equal(other.download_pano(0, Wp, Hp), pano)
            consts = [pkg.pitch_constants(128, 90, 70)]
            v0 = proj.project(a, [10], consts, 128, 96)
            v1 = other.project(0, [10], consts, 128, 96)
            proj.sync(a)
            other.sync(0)
            assert np.array_equal(v0, v1)
            with pytest.raises(pkg.P2PError):
                proj.copy_pano_from(a, proj, a)
            with pytest.raises(pkg.P2PError) as e:
                other.copy_pano_from(0, other, 1)    # empty source slot
            assert e.value.code == -4
    finally:
        other.close()


# ------------------------------------------------------------------------------------------
# one image split over several GPUs; a folder sharded over GPUs (8f-4)
# ------------------------------------------------------------------------------------------
def test_single_image_split_pixels_identical(pkg):
    Wp, Hp, W, H, fov = 2048, 1024, 480, 270, 120
    pano = synth.noise(Wp, Hp, 10)
    yaws, pitches = [0, 90, 180, 270], [30, 60, 90]
    one = [pkg.process_yaw_and_pitchs(pano, y, pitches, W, H, fov) for y in yaws]
    want = fp.project_view_single_pass(pano, 90, 60, W, H, fov)
    if svml_model.host_numpy_uses_svml():
        assert np.array_equal(one[1][1], want)
    else:
        assert (one[1][1] == want).all(axis=-1).mean() >= 0.93
    from p2p_b200 import panorama_to_plane_pitch as front

    for devs in _device_sets(pkg):
        out = front._project(pkg.get_projector(), pano, yaws, pitches, W, H, fov, devices=devs)
        for k in range(len(yaws)):
            for j in range(len(pitches)):
                assert np.array_equal(out[k, j], one[k][j]), (devs, k, j)
    # module-level switch: the reference-named entry points split every image
    try:
        pkg.set_devices(_device_sets(pkg)[-1])
        got = pkg.process_yaw_and_pitchs(pano, 180, pitches, W, H, fov)
        assert all(np.array_equal(a, b) for a, b in zip(got, one[2]))
    finally:
        pkg.set_devices(None)


@pytest.mark.parametrize("fmt", ["png", "jpg"])
def test_single_image_split_files_identical(pkg, tmp_path, fmt):
    cv2 = pytest.importorskip("cv2")
    Wp, Hp, W, H, fov = 1024, 512, 200, 120, 100
    yaws, pitches = [0, 90, 180], [60, 120]
    src = tmp_path / "in" / "pano.jpg"
    src.parent.mkdir()
    assert cv2.imwrite(str(src), synth.smooth(Wp, Hp, 3))
    ref_dir = tmp_path / "ref"
    pkg.main(str(src), str(ref_dir), yaws, pitches, W, H, num_workers=2, output_format=fmt, fov_deg=fov)
    names = sorted(p.name for p in ref_dir.iterdir())
    assert len(names) == len(yaws) * len(pitches)
    for i, devs in enumerate(_device_sets(pkg)):
        out = tmp_path / f"split{i}"
        pkg.main(str(src), str(out), yaws, pitches, W, H, num_workers=2, output_format=fmt, fov_deg=fov, devices=devs)
        assert sorted(p.name for p in out.iterdir()) == names
        for n in names:
            assert (out / n).read_bytes() == (ref_dir / n).read_bytes(), (devs, n)
    # the CLI flag
    out = tmp_path / "cli"
    pkg.cli(["--input_path", str(src), "--output_path", str(out), "--FOV", str(fov), "--output_width", str(W),
             "--output_height", str(H), "--yaw_angles", *map(str, yaws), "--pitch_angles", *map(str, pitches),
             "--output_format", fmt, "--num_workers", "2", "--devices", "0", "0"])
    for n in names:
        assert (out / n).read_bytes() == (ref_dir / n).read_bytes(), n


def test_folder_sharded_over_devices_identical(pkg, tmp_path):
    """``main(devices=...)`` on a folder (ref :320-341 walks the files one after the other): same files as one device."""
    cv2 = pytest.importorskip("cv2")
    src = tmp_path / "in"
    src.mkdir()
    for i in range(7):
        ext = "jpg" if i % 2 else "png"
        assert cv2.imwrite(str(src / f"p{i}.{ext}"), synth.smooth(1024 if i % 3 else 2048, 512 if i % 3 else 1024, i))
    W, H, fov, yaws, pitches = 200, 120, 100, [0, 90], [60, 120]
    ref_dir = tmp_path / "ref"
    pkg.main(str(src), str(ref_dir), yaws, pitches, W, H, num_workers=3, output_format="jpg", fov_deg=fov)
    names = sorted(p.name for p in ref_dir.iterdir())
    assert len(names) == 7 * 4
    n_dev = pkg._lib.load().p2p_device_count()
    for i, devs in enumerate([[0, 0]] + ([list(range(min(n_dev, 8)))] if n_dev >= 2 else [])):
        out = tmp_path / f"d{i}"
        pkg.main(str(src), str(out), yaws, pitches, W, H, num_workers=3, output_format="jpg", fov_deg=fov, devices=devs)
        assert sorted(p.name for p in out.iterdir()) == names
        for n in names:
            assert (out / n).read_bytes() == (ref_dir / n).read_bytes(), (devs, n)


# ------------------------------------------------------------------------------------------
# the last error is per calling thread (unlocked std::string shared by the worker threads)
# ------------------------------------------------------------------------------------------
def test_last_error_is_per_thread(pkg, proj):
    lib = pkg._lib.load()
    consts = pkg.Projector._consts_array([pkg.pitch_constants(64, 90, 90)])
    msgs = {}

    def worker(name, slot):
        import ctypes as C

        sh = (C.c_int32 * 1)(0)
        out = np.empty(64 * 64 * 3, np.uint8)
        for _ in range(200):
            rc = lib.p2p_project_views(proj.ctx, slot, 1, sh, 1, consts, 64, 64, out.ctypes.data, 0)
            assert rc < 0
            msgs[name] = lib.p2p_last_error(proj.ctx).decode()
            assert msgs[name] == ("bad slot" if slot == 99 else "slot holds no panorama"), msgs[name]

    fresh = pkg.Projector(0, n_slots=1)
    try:
        t = [threading.Thread(target=worker, args=("a", 99)), threading.Thread(target=worker, args=("b", 99))]
        for x in t:
            x.start()
        for x in t:
            x.join()
        assert msgs == {"a": "bad slot", "b": "bad slot"}
        # a thread that never failed on this context sees an empty message, not another thread's
        seen = []
        th = threading.Thread(target=lambda: seen.append(lib.p2p_last_error(fresh.ctx).decode()))
        th.start()
        th.join()
        assert seen == [""]
    finally:
        fresh.close()


def test_panorama_assembled_from_pieces(pkg, proj):
    """``p2p_upload_pano_rows`` / ``p2p_copy_pano_rows``: pieces that touch extend the rows a slot holds (the all-gather of
    an image split over GPUs), a piece elsewhere replaces them; projecting needs the rows the views touch."""
    Wp, Hp, W, H, fov = 1024, 512, 128, 96, 90
    pano = synth.noise(Wp, Hp, 12)
    consts = [pkg.pitch_constants(W, fov, 90)]
    with proj.slots(3) as (a, b, c):
        proj.upload(a, pano)
        want = proj.project(a, [0], consts, W, H)
        proj.sync(a)
        # pieces in any contiguous order
        proj.upload_rows(b, pano, 200, 300)
        proj.upload_rows(b, pano, 300, Hp)
        with pytest.raises(pkg.P2PError) as e:     # rows 0 .. 199 are still missing
            proj.project(b, [0], consts, W, H)
        assert e.value.code == -4
        proj.upload_rows(b, pano, 0, 200)
        got = proj.project(b, [0], consts, W, H)
        proj.sync(b)
        assert np.array_equal(got, want)
        assert np.array_equal(proj.download_pano(b, Wp, Hp), pano)
        # all-gather between two slots
        proj.upload_rows(b, synth.noise(512, 256, 1), 0, 10)   # another image: slot b forgets the rows it held
        proj.upload_rows(b, pano, 0, 256)                      # -> rows 0 .. 255 of `pano` only
        with pytest.raises(pkg.P2PError) as e:
            proj.project(b, [0], consts, W, H)
        assert e.value.code == -4
        proj.upload_rows(c, pano, 256, Hp)
        proj.copy_pano_rows_from(b, proj, c, 256, Hp + 1)
        proj.copy_pano_rows_from(c, proj, b, 0, 256)
        for s in (b, c):
            got = proj.project(s, [0], consts, W, H)
            proj.sync(s)
            assert np.array_equal(got, want)
        proj.upload_rows(a, synth.noise(512, 256, 1), 0, 10)   # another size: slot a now holds rows 0 .. 9 of that image
        with pytest.raises(pkg.P2PError) as e:                 # the source does not hold these rows
            proj.copy_pano_rows_from(b, proj, a, 5, 50)
        assert e.value.code == -4
        with pytest.raises(pkg.P2PError):
            proj.upload_rows(a, pano, 10, Hp + 1)
    # the front-end helper over devices (a device may be listed twice)
    for devs in _device_sets(pkg):
        projs = [pkg.get_projector(d) for d in devs]
        import contextlib

        with contextlib.ExitStack() as st:
            slots = [st.enter_context(p.slots(1))[0] for p in projs]
            keep = pkg.scatter_upload(projs, slots, pano)
            for p, s in zip(projs, slots):
                got = p.project(s, [0], consts, W, H)
                p.sync(s)
                assert np.array_equal(got, want), devs
                assert np.array_equal(p.download_pano(s, Wp, Hp), pano), devs
            del keep


@pytest.mark.parametrize("W,H", [(8, 5), (64, 9), (72, 16), (264, 20), (520, 11), (1928, 8)])
def test_row_segment_lengths_identical(pkg, W, H):
    """The row-segment kernel writes 32-bit words everywhere except at the two ends of a warp's segment (byte stores by
    lanes 0 / 31): every segment length - one chunk per warp up to the whole row - must give the per-pixel kernel's bytes."""
    L = pkg._lib
    Wp, Hp, fov = 2048, 1024, 110
    pano = synth.noise(Wp, Hp, 13)
    yaws, pitches = [0, 90, 180, 270, 45], [35, 90, 160]   # five yaws: a group of four and a group of one per pitch
    p = pkg.Projector(0, n_slots=2)
    try:
        p.set_option(L.OPT_MIRROR, 0)
        want = p.project_image(pano, yaws, pitches, W, H, fov).copy()
        p.set_option(L.OPT_MIRROR, 2)
        for seg in (1, 2, 3, 4, 5, 7, 64):
            p.set_option(L.OPT_SEG_CHUNKS, seg)
            got = p.project_image(pano, yaws, pitches, W, H, fov)
            assert np.array_equal(got, want), seg
    finally:
        p.close()


def test_integration_md_multi_gpu_snippet_runs(pkg):
    """The view-list / split snippet printed in INTEGRATION.md, executed verbatim (two contexts on device 0 play two GPUs;
    with >= 2 devices the second context lives on device 1)."""
    import ctypes as C
    import re
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    blocks = re.findall(r"```python\n(.*?)```", (root / "INTEGRATION.md").read_text(), flags=re.S)
    stub = next(b for b in blocks if "def process_yaw_and_pitchs" in b and "C.CDLL" in b)
    snippet = next(b for b in blocks if "p2p_project_view_list.argtypes" in b)
    ns = {}
    exec(compile(stub.replace('C.CDLL("libp2p_b200.so")', f'C.CDLL(r"{pkg._lib.LIB_PATH}")'), "INTEGRATION.md", "exec"), ns)
    lib = ns["_lib"]
    n_dev = lib.p2p_device_count()
    ctxs = []
    for r in range(2):
        c = C.c_void_p()
        assert lib.p2p_create(min(r, n_dev - 1), 2, C.byref(c)) == 0
        ctxs.append(c)
    Wp, Hp, W, H, fov = 1024, 512, 128, 96, 90
    pano = synth.noise(Wp, Hp, 14)
    views = FACES
    shifts_l, consts_l = _flat(pkg, Wp, W, fov, views)
    shifts = (C.c_int32 * len(views))(*shifts_l)
    consts = (ns["PitchConsts"] * len(views))()
    for i, (f, c, s) in enumerate(consts_l):
        consts[i].f, consts[i].c, consts[i].s = f, c, s
    out = np.zeros((len(views), H, W, 3), np.uint8)

    def ck(rc):
        assert rc == 0, lib.p2p_last_error(ctxs[0]).decode()

    lib.p2p_destroy.argtypes = [C.c_void_p]
    ns.update(ctxs=ctxs, n=2, pano=pano, Wp=Wp, Hp=Hp, n_views=len(views), shifts=shifts, consts=consts, W=W, H=H, out=out,
              ck=ck)
    try:
        exec(compile(snippet, "INTEGRATION.md", "exec"), ns)
        proj = pkg.get_projector()
        with proj.slots(1) as (s,):
            proj.upload(s, pano)
            want = proj.project_list(s, shifts_l, consts_l, W, H)
            proj.sync(s)
        assert np.array_equal(out, want)
    finally:
        for c in ctxs:
            lib.p2p_destroy(c)
