"""The drop-in `wass_stereo` executable: CLI contract on CPU (usage, --genconfig, config errors, exit codes;
SURVEY.md §8b) and a full workdir run on the GPU."""
import os
import subprocess
import numpy as np
import pytest
from helpers import ROOT

EXE = os.path.join(ROOT, "wass_b200", "bin", "wass_stereo")


def run(args, cwd=None, env=None):
    return subprocess.run([EXE] + args, capture_output=True, text=True, cwd=cwd, env=env)


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(EXE):
        from wass_b200 import build
        build.build()


def test_no_arguments_is_usage_and_exit_0():
    r = run([])
    assert r.returncode == 0          # drivers use this as a liveness probe (cli/wasscli/wasscli.py:66-71)
    assert "wass_stereo  v." in r.stdout and "Usage:" in r.stdout
    assert "wass_stereo [--genconfig] <config_file> <workdir> [--measure] [--rectify-only]" in r.stdout


def test_genconfig_format(tmp_path):
    r = run(["--genconfig"], cwd=tmp_path)
    assert r.returncode == 0
    txt = (tmp_path / "stereo_config.txt").read_text()
    # incfg.hpp:436-450: "# desc\n# \n[#]KEY=value\n\n", keys sorted, defaults commented out
    assert "# Stereo match window size\n# \n#WINSIZE=13\n\n" in txt
    assert '#LEFT_MASK_IMAGE="none"\n' in txt and "#SAVE_COMPRESSED=true\n" in txt and "#ZGAP_PERCENTILE=99\n" in txt
    assert "#SAVE_INPUT_SCALE=0.3\n" in txt and "#PLANE_REFINE_XMIN=-9999\n" in txt and "#MAX_DISPARITY=640\n" in txt
    keys = [l.lstrip("#").split("=")[0] for l in txt.splitlines() if "=" in l and not l.startswith("# ")]
    assert keys == sorted(keys) and len(keys) >= 47
    for k in ("RANDOM_SEED", "MIN_TRIANGULATED_POINTS", "DISABLE_AUTO_LEFT_RIGHT", "PLANE_RANSAC_ROUNDS", "DENSE_P2_MULT",
              "TRIANG_MIN_ANGLE", "PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE", "DISCARD_BURNED_AREAS", "DISPARITY_OFFSET"):
        assert k in keys


def test_bad_invocations(tmp_path):
    assert run(["a"]).returncode == 255                          # -1
    assert run(["cfg.txt", str(tmp_path / "missing_wd")]).returncode == 255
    wd = tmp_path / "wd"; wd.mkdir()
    assert run([str(tmp_path / "nocfg.txt"), str(wd)]).returncode == 255
    cfg = tmp_path / "cfg.txt"
    cfg.write_text("NOT_A_KEY=3\n")
    r = run([str(cfg), str(wd)])
    assert r.returncode == 255 and "Unexpected key: NOT_A_KEY" in r.stdout
    cfg.write_text("SAVE_AS_PLY=yes\n")
    r = run([str(cfg), str(wd)])
    assert r.returncode == 255 and 'Unable to parse yes to "true" or "false"' in r.stdout
    assert (wd / "wass_stereo_log.txt").exists()


def test_config_parsing_rules(tmp_path):
    # spaces stripped outside quotes, '#' comments, CRLF, non-default keys un-commented in the saved copy
    wd = tmp_path / "wd"; wd.mkdir()
    cfg = tmp_path / "cfg.txt"
    cfg.write_text("# comment\r\n  WINSIZE = 11 \r\nLEFT_MASK_IMAGE = \"my mask.png\"\r\n\r\nSAVE_AS_PLY=true\n")
    r = run([str(cfg), str(wd)])
    assert r.returncode == 255                # no input data in the workdir (or no GPU) -> fails after the config stage
    saved = (wd / "stereo_config.txt").read_text()
    assert "\nWINSIZE=11\n" in saved and '\nLEFT_MASK_IMAGE="my mask.png"\n' in saved and "\nSAVE_AS_PLY=true\n" in saved
    assert "#MAX_DISPARITY=640\n" in saved


@pytest.mark.gpu
def test_full_workdir_run(tmp_path):
    from wass_b200 import synth, workdir, capi
    from oracle import pipeline as op
    W, H, D = 640, 480, 64
    right, left, dtrue = synth.make_pair(W, H, D, seed=1, d0=8.0)
    c = synth.make_calibration(W, H)
    wd = tmp_path / "000000_wd"
    # cam0 = left, cam1 = right, X1 = X0 + T with T.x > 0 (no auto-swap)
    workdir.write_workdir(str(wd), left, right, c["K0"], c["K1"], c["R"], c["T"])
    cfg = tmp_path / "stereo_config.txt"
    workdir.write_config(str(cfg), MAX_DISPARITY=D, RANDOM_SEED=7, SAVE_AS_PLY=True, PLANE_RANSAC_ROUNDS=60)
    r = run([str(cfg), str(wd)])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    for tok in ("[P|10|100]", "[P|20|100]", "[P|40|100]", "[P|60|100]", "[P|80|100]", "[P|90|100]", "[P|100|100]", "All done."):
        assert tok in r.stdout
    for f in ("mesh_cam.xyzC", "plane.txt", "P0cam.txt", "P1cam.txt", "Cam0_poseR.txt", "Cam0_poseT.txt", "Cam1_poseR.txt",
              "Cam1_poseT.txt", "wass_stereo_log.txt", "stereo_config.txt", "mesh.ply", "plane_refinement_inliers.xyz",
              "00000000_s.png", "K0_small.txt", "scale.txt"):
        assert (wd / f).exists(), f
    # the diagnostic JPEGs the reference writes (wass_stereo.cpp:833, 854, 1001-1017, 1911-1926; PovMesh.cpp:982): present,
    # decodable by an independent decoder, of the reference's sizes, and showing what they should
    import cv2
    rh, rw = cv2.imread(str(wd / "disparity_stereo_ouput.jpg"), cv2.IMREAD_UNCHANGED).shape        # the common ROI
    assert H - 4 <= rh <= H and W - 4 <= rw <= W
    shapes = {"stereo.jpg": (H, 2 * W, 3), "stereo_input.jpg": (2 * rh, rw + D, 1), "disparity_final_scaled.jpg": (H, W, 1),
              "disparity_coverage.jpg": (H // 2, W // 2, 3),
              "graph_components.jpg": (int(np.rint(rh * 0.5)), int(np.rint(rw * 0.5)), 3)}      # cv::resize: cvRound(n * 0.5)
    for f, shp in shapes.items():
        im = cv2.imread(str(wd / f), cv2.IMREAD_UNCHANGED)
        assert im is not None, f
        assert im.shape == (shp[:2] if shp[2] == 1 else shp), (f, im.shape)
    st = cv2.imread(str(wd / "stereo.jpg"), cv2.IMREAD_GRAYSCALE).astype(int)
    assert np.abs(st[5:15, 10:W - 10] - left[5:15, 10:W - 10].astype(int)).mean() < 3        # left | right side by side
    assert np.abs(st[5:15, W + 10:2 * W - 10] - right[5:15, 10:W - 10].astype(int)).mean() < 3
    gc = cv2.imread(str(wd / "graph_components.jpg"))
    assert (gc[..., 1] > 128).mean() > 0.5                                                  # the kept component, in green
    dj = cv2.imread(str(wd / "disparity_final_scaled.jpg"), cv2.IMREAD_GRAYSCALE)
    assert dj[H // 2, W // 2 - 50:W // 2 + 50].mean() > 20 and dj.min() < 8                    # disparities inside, holes rendered dark
    P1 = workdir.load_matrix_txt(str(wd / "P1cam.txt"))
    assert np.allclose(P1, c["K1"] @ np.hstack([c["R"], c["T"].reshape(3, 1)]))
    assert (wd / "P0cam.txt").read_text().count("\n") == 2 and "e+" in (wd / "P0cam.txt").read_text()
    plane = np.array([float(x) for x in (wd / "plane.txt").read_text().split()])
    assert plane.shape == (4,) and abs(np.linalg.norm(plane[:3]) - 1) < 1e-9 and plane[2] > 0
    pts = workdir.load_camera_mesh(str(wd / "mesh_cam.xyzC"))
    assert pts.shape[0] > 0.5 * W * H
    dist = np.abs(pts @ plane[:3] + plane[3])
    assert dist.max() < 1.5 + 1e-2                               # PLANE_MAX_DISTANCE crop
    # same frame through the library in-process: same point count and plane
    h = capi.Handle(0)
    cal = capi.stereo_rectify(c["K0"], c["K1"], c["R"], c["T"], W, H)
    lr = h.rectify_image(left, c["K0"], cal["R1"], cal["P1"])
    rr = h.rectify_image(right, c["K1"], cal["R2"], cal["P2"])
    ymin = max(cal["roi1"][1], cal["roi2"][1]); ymax = min(cal["roi1"][1] + cal["roi1"][3], cal["roi2"][1] + cal["roi2"][3])
    wroi = min(cal["roi1"][2], cal["roi2"][2])
    rl = (cal["roi1"][0], ymin, wroi, ymax - ymin); rrr = (cal["roi2"][0], ymin, wroi, ymax - ymin)
    h.dense_stereo(lr[rl[1]:rl[1] + rl[3], rl[0]:rl[0] + rl[2]].copy(), rr[rrr[1]:rrr[1] + rrr[3], rrr[0]:rrr[0] + rrr[2]].copy(),
                   capi.dense_params(MAX_DISPARITY=D))
    calib = dict(K0=c["K0"], K1=c["K1"], R=c["R"], T=c["T"], R1=cal["R1"], R2=cal["R2"], P1=cal["P1"], P2=cal["P2"], roi_left=rl, roi_right=rrr)
    n = h.triangulate_from_dense(left, right, calib, (H, W))
    assert ("%d valid points found" % n) in r.stdout
    h.close()
    # depth against the generator's ground truth
    z_true = W / dtrue
    # (the exported cloud is cropped to PLANE_MAX_DISTANCE around the fitted plane, so compare ranges, not medians)
    assert z_true.min() * 0.9 < np.percentile(pts[:, 2], 1) and np.percentile(pts[:, 2], 99) < z_true.max() * 1.1


@pytest.mark.gpu
def test_debug_images_can_be_switched_off(tmp_path):
    from wass_b200 import synth, workdir
    W, H, D = 320, 240, 48
    right, left, _ = synth.make_pair(W, H, D, seed=5, d0=8.0)
    c = synth.make_calibration(W, H)
    wd = tmp_path / "000000_wd"
    workdir.write_workdir(str(wd), left, right, c["K0"], c["K1"], c["R"], c["T"])
    cfg = tmp_path / "stereo_config.txt"
    workdir.write_config(str(cfg), MAX_DISPARITY=D, RANDOM_SEED=7, PLANE_RANSAC_ROUNDS=60, SAVE_DEBUG_IMAGES=False)
    r = run([str(cfg), str(wd)])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert (wd / "mesh_cam.xyzC").exists() and not [f for f in os.listdir(wd) if f.endswith(".jpg")]


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [0.5, 1.5])
def test_workdir_run_with_dense_scale(tmp_path, scale):
    """DENSE_SCALE != 1 through the executable (wass_stereo.cpp:783-797, 903-928, 1172-1180): the matcher runs on resized
    crops and the disparity comes back in full-resolution pixels, so the fitted plane must agree with the DENSE_SCALE=1 run."""
    from wass_b200 import synth, workdir
    W, H, D = 640, 480, 64
    right, left, _ = synth.make_pair(W, H, D, seed=2, d0=8.0)
    c = synth.make_calibration(W, H)
    planes = []
    for s in (1.0, scale):
        wd = tmp_path / ("wd_%g" % s)
        workdir.write_workdir(str(wd), left, right, c["K0"], c["K1"], c["R"], c["T"])
        cfg = tmp_path / ("cfg_%g.txt" % s)
        # at scale s the disparities are s times as large: keep the search range covering them
        workdir.write_config(str(cfg), MAX_DISPARITY=D if s <= 1 else 2 * D, RANDOM_SEED=7, PLANE_RANSAC_ROUNDS=60, DENSE_SCALE=s)
        r = run([str(cfg), str(wd)])
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        planes.append(np.array([float(x) for x in (wd / "plane.txt").read_text().split()]))
        pts = workdir.load_camera_mesh(str(wd / "mesh_cam.xyzC"))
        assert pts.shape[0] > 0.3 * W * H
    assert np.abs(planes[0][:3] @ planes[1][:3]) > 0.999            # same normal within ~2.5 degrees
    assert abs(planes[0][3] - planes[1][3]) < 0.05 * abs(planes[0][3])


@pytest.mark.gpu
@pytest.mark.parametrize("angle", [0.0, 4.0])
def test_workdir_run_with_custom_rectifier(tmp_path, angle):
    """USE_CUSTOM_STEREORECTIFY through the executable (wass_stereo.cpp:496-529, stereorectify.cpp): homographies written,
    points triangulated through their inverses; the fitted plane agrees with the cv::stereoRectify run of the same frame."""
    from wass_b200 import synth, workdir
    W, H, D = 640, 480, 64
    right, left, _ = synth.make_pair(W, H, D, seed=4, d0=8.0)
    c = synth.make_calibration(W, H)
    planes = []
    for custom in (False, True):
        wd = tmp_path / ("wd_%d" % custom)
        workdir.write_workdir(str(wd), left, right, c["K0"], c["K1"], c["R"], c["T"])
        cfg = tmp_path / ("cfg_%d.txt" % custom)
        workdir.write_config(str(cfg), MAX_DISPARITY=D, RANDOM_SEED=7, PLANE_RANSAC_ROUNDS=60, USE_CUSTOM_STEREORECTIFY=custom,
                             RECTIFY_ANGLE=angle)
        r = run([str(cfg), str(wd)])
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        if custom:
            assert "Using WASS custom stereorectify" in r.stdout
            for f in ("H0_rect.txt", "H1_rect.txt"):
                Hm = workdir.load_matrix_txt(str(wd / f))
                assert Hm.shape == (3, 3) and abs(np.linalg.det(Hm) - 1) < 1e-9
        planes.append(np.array([float(x) for x in (wd / "plane.txt").read_text().split()]))
        pts = workdir.load_camera_mesh(str(wd / "mesh_cam.xyzC"))
        assert pts.shape[0] > 0.3 * W * H
    # (a rotated rectifying plane crops a different part of the rippled synthetic surface, and RANSAC samples other triples:
    # the normal moves by up to ~3 degrees, the offset by a few %)
    assert np.abs(planes[0][:3] @ planes[1][:3]) > 0.998
    assert abs(planes[0][3] - planes[1][3]) < 0.1 * abs(planes[0][3])


@pytest.mark.gpu
@pytest.mark.parametrize("rounds,thr,expect_plane", [(60, 1.0, True), (25, 1e-7, False)])
def test_workdir_run_matches_oracle_pipeline(tmp_path, rounds, thr, expect_plane):
    """The same workdir, configuration and RANDOM_SEED through the drop-in executable and through the CPU oracle
    (oracle/sgbm_oracle.c pinned to cv2, oracle/pipeline.py pinned to the reference's own PovMesh.cpp / triangulate.hpp):
    plane.txt within 1e-6, mesh_cam.xyzC the same points to +-1 LSB, mesh.ply the same floats; also the soft RANSAC failure
    (plane.txt = nan, mesh still written with the best hypothesis, wass_stereo.cpp:2101-2107)."""
    import struct
    from wass_b200 import synth, workdir
    from oracle import sgbm, pipeline as op
    W, H, D, seed = 320, 240, 48, 11
    right, left, _ = synth.make_pair(W, H, D, seed=3, d0=8.0)
    left = left.copy(); left[60:70, 100:120] = 255             # burned pixels in the original left image
    c = synth.make_calibration(W, H)
    wd = tmp_path / "000000_wd"
    workdir.write_workdir(str(wd), left, right, c["K0"], c["K1"], c["R"], c["T"])
    cfg = tmp_path / "stereo_config.txt"
    workdir.write_config(str(cfg), MAX_DISPARITY=D, RANDOM_SEED=seed, SAVE_AS_PLY=True, PLANE_RANSAC_ROUNDS=rounds,
                         PLANE_RANSAC_THRESHOLD=thr)
    r = run([str(cfg), str(wd)])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]

    # ---- the oracle on the same inputs (the rig is already rectified: cv::stereoRectify gives R1 = R2 = I, full-frame ROIs)
    cal = op.rectified_calibration_identity(c["K0"], c["T"], W, H)
    rl, rr = cal["roi_left"], cal["roi_right"]
    lc = left[rl[1]:rl[1] + rl[3], rl[0]:rl[0] + rl[2]]
    rc = right[rr[1]:rr[1] + rr[3], rr[0]:rr[0] + rr[2]]
    i1, i2 = synth.pad_for_sgbm(rc, lc, D)
    d16 = sgbm.compute(i1, i2, sgbm.wass_params(D, mode=0))["disp"][:, D:D + rr[2]]      # MODE_SGBM: the executable's default
    disp = np.zeros((H, W), np.float32)
    disp[rr[1]:rr[1] + rr[3], rr[0]:rr[0] + rr[2]] = op.postprocess_disparity(d16, 1, D)
    tri = op.triangulate(disp, cal, left, right)
    assert ("%d valid points found" % tri["n"]) in r.stdout
    valid, p3d = tri["valid"], tri["p3d"]
    zgap = op.zgap_percentile(valid, p3d[..., 2], 99.0)
    comp = op.biggest_component(valid, p3d[..., 2], zgap)
    triples = op.ransac_draw_triples(op.LibcRand(seed), rr[2], rr[3], rounds)
    ok, plane, _ = op.ransac_find_plane(comp, p3d, triples, thr)
    assert ok == expect_plane
    mask = comp
    if ok:
        mask = op.crop_plane(comp, p3d, plane, thr)
        plane, _ = op.refine_plane(mask, p3d)
        mask = op.crop_plane(mask, p3d, plane, 1.5)
    ref_xyzc = op.xyz_compressed_bytes(mask, p3d, plane)

    # ---- plane.txt
    txt = (wd / "plane.txt").read_text().split()
    if ok:
        got = np.array([float(x) for x in txt])
        assert np.abs(got - plane).max() <= 1e-6
    else:
        assert txt == ["nan"] * 4
    # ---- mesh_cam.xyzC: same count, header within 1e-6 relative, quantised points +-1 LSB
    buf = (wd / "mesh_cam.xyzC").read_bytes()
    n = struct.unpack_from("<I", buf, 0)[0]
    assert n == struct.unpack_from("<I", ref_xyzc, 0)[0] == int(mask.sum())
    hg, hr = np.array(struct.unpack_from("<18d", buf, 4)), np.array(struct.unpack_from("<18d", ref_xyzc, 4))
    assert np.allclose(hg, hr, rtol=1e-6, atol=1e-9)
    qg = np.frombuffer(buf, "<u2", 3 * n, 148).astype(np.int64)
    qr = np.frombuffer(ref_xyzc, "<u2", 3 * n, 148).astype(np.int64)
    assert np.abs(qg - qr).max() <= 1
    # ---- mesh.ply (PovMesh.cpp:463-517): header and the 15-byte records, points in grid order, grey from the right image
    ply = (wd / "mesh.ply").read_bytes()
    head, body = ply.split(b"end_header\n", 1)
    assert head.decode().split("\n")[:3] == ["ply", "format binary_little_endian 1.0", "element vertex %d" % n]
    rec = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "u1", 3)]))
    assert rec.shape[0] == n
    assert np.abs(rec["p"] - p3d[mask].astype(np.float32)).max() <= 1e-5 * np.abs(p3d[mask]).max()
    assert np.array_equal(rec["c"], np.repeat(tri["color"][mask][:, None], 3, axis=1))


@pytest.mark.gpu
def test_batch_mode_equals_one_process_per_frame(tmp_path):
    """`wass_stereo --batch <config> <wd>...` (one process, one warm arena, one batched matcher run per --batch-size frames,
    PNG decode of the next frames on a second thread) leaves in every workdir exactly what one process per frame leaves
    (cli/wasscli/wasscli.py:326-346), and writes the planes in frame order."""
    from wass_b200 import synth, workdir
    W, H, D, n = 320, 240, 48, 5
    c = synth.make_calibration(W, H)
    cfg = tmp_path / "stereo_config.txt"
    workdir.write_config(str(cfg), MAX_DISPARITY=D, RANDOM_SEED=5, PLANE_RANSAC_ROUNDS=40, SGM_FULL_8PATH=True)
    single, batch = [], []
    for i in range(n):
        right, left, _ = synth.make_pair(W, H, D, seed=20 + i, d0=8.0 + i)
        for tag, lst in (("s", single), ("b", batch)):
            wd = tmp_path / ("%s_%06d_wd" % (tag, i))
            workdir.write_workdir(str(wd), left, right, c["K0"], c["K1"], c["R"], c["T"])
            lst.append(wd)
    (tmp_path / "b_000003_wd" / "undistorted" / "00000001.png").unlink()       # one broken frame must not stop the others
    (tmp_path / "s_000003_wd" / "undistorted" / "00000001.png").unlink()
    for wd in single:
        r = run([str(cfg), str(wd)])
        assert (r.returncode == 0) == (wd.name != "s_000003_wd")
    planes_out = tmp_path / "planes.txt"
    r = run(["--batch", "--batch-size", "2", "--planes-out", str(planes_out), str(cfg)] + [str(w) for w in batch])
    assert r.returncode == 255 and "[batch] FAILED" in r.stdout and "1 failed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    rows = [l.split() for l in planes_out.read_text().splitlines()]
    assert len(rows) == n and rows[3] == ["nan"] * 4
    for i, (ws, wb) in enumerate(zip(single, batch)):
        if i == 3:
            assert not (wb / "mesh_cam.xyzC").exists()
            continue
        assert (ws / "plane.txt").read_text() == (wb / "plane.txt").read_text()
        assert (ws / "mesh_cam.xyzC").read_bytes() == (wb / "mesh_cam.xyzC").read_bytes()
        assert (ws / "stereo_config.txt").read_text() == (wb / "stereo_config.txt").read_text()
        for f in ("P0cam.txt", "P1cam.txt", "Cam1_poseT.txt", "plane_refinement_inliers.xyz", "00000000_s.png"):
            assert (ws / f).read_bytes() == (wb / f).read_bytes(), f
        assert "All done." in (wb / "wass_stereo_log.txt").read_text()
        assert np.allclose([float(v) for v in rows[i]], [float(v) for v in (wb / "plane.txt").read_text().split()], rtol=0, atol=1e-15)
    mean = [float(v) for v in r.stdout.split("frames: ")[1].split()[:4]]
    assert np.allclose(mean, np.nanmean(np.array(rows, float), axis=0), atol=1e-14)
