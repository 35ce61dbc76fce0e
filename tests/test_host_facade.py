"""The C++ drop-in layer (ro_map_b200/host): nerf::NerfManagerOffline / NerfManagerOnline / NeRF over the C ABI."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def host_lib():
    from ro_map_b200 import build
    return build.build_host()


def test_facade_exports_reference_interface(host_lib):
    """Every member function the reference's clients call is exported with the reference's name."""
    syms = subprocess.run(["nm", "-DC", "--defined-only", str(host_lib)], capture_output=True, text=True, check=True).stdout
    wanted = [
        "nerf::NerfManagerOffline::NerfManagerOffline(", "nerf::NerfManagerOffline::Init()", "nerf::NerfManagerOffline::ReadDataset()",
        "nerf::NerfManagerOffline::CreateNeRF(", "nerf::NerfManagerOffline::WaitThreadsEnd()", "nerf::NerfManagerOffline::GetNeRF(int)",
        "nerf::NerfManagerOffline::GetAllNeRF()", "nerf::NerfManagerOffline::GetAllTwc()", "nerf::NerfManagerOffline::GetIntrinsics(",
        "nerf::NerfManagerOnline::NerfManagerOnline(", "nerf::NerfManagerOnline::Init()", "nerf::NerfManagerOnline::DatasetInit(",
        "nerf::NerfManagerOnline::NewFrameToDataset(", "nerf::NerfManagerOnline::UpdateDataset(", "nerf::NerfManagerOnline::CreateNeRF(",
        "nerf::NerfManagerOnline::GetFrameIdx(double)", "nerf::NerfManagerOnline::UpdateNeRFBbox(", "nerf::NerfManagerOnline::DrawMesh(",
        "nerf::NerfManagerOnline::WaitThreadsEnd()", "nerf::NerfManagerOnline::RenderNeRFsTest(",
        "nerf::NeRF::GetFrameIdAndBBox()", "nerf::NeRF::GetObjTow()", "nerf::NeRF::GetBoundingBox()", "nerf::NeRF::DrawCPUMesh()",
        "nerf::NeRF::TrainOffline(int)", "nerf::NeRF::TrainOnline()", "nerf::NeRF::UpdateFrameBBox(", "nerf::NeRF::RenderTestImg(",
    ]
    missing = [w for w in wanted if w not in syms]
    assert not missing, missing
    # no CUDA runtime and no test-oracle dependency in the facade itself: it talks to the core through the C ABI only
    needed = subprocess.run(["readelf", "-d", str(host_lib)], capture_output=True, text=True, check=True).stdout
    assert "libmon_b200.so" in needed and "libcudart" not in needed and "oracle" not in needed


def test_headless_driver_usage_message():
    from ro_map_b200 import build
    build.build_host()
    p = subprocess.run([str(build.HOST_BIN)], capture_output=True, text=True)
    assert p.returncode == 1 and "usage: offline_nerf" in p.stderr


@pytest.mark.gpu
def test_offline_nerf_on_disk_sequence(tmp_path, host_lib):
    """End to end through the reference's on-disk schema: PNG keyframes + config.yaml + obj_offline/k.txt -> the
    headless OfflineNeRF -> trained objects, logged losses, rendered test view (PNG)."""
    import cv2
    from ro_map_b200 import build, synthetic as syn
    seq = syn.make_sequence(n_frames=8, n_objects=2, seed=1337, H=240, W=240, K=(333.333, 333.333, 120.0, 120.0))
    syn.write_sequence(seq, str(tmp_path / "seq"))
    cfg = ROOT / "ro_map_b200" / "configs" / "base.json"
    p = subprocess.run([str(build.HOST_BIN), str(cfg), str(tmp_path / "seq"), "1", "2", "2"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("object ")]
    assert len(lines) == 2, p.stdout
    for l in lines:
        tok = l.split()
        step, loss = int(tok[tok.index("step") + 1]), float(tok[tok.index("loss") + 1])
        assert step == 1000 and np.isfinite(loss) and loss < 0.05, l
    # per-Train_Step log lines like the reference's (nerf_model.cu:1661-1662)
    assert sum("train_time:" in l and "Step:" in l for l in p.stdout.splitlines()) == 4
    img = cv2.imread(str(tmp_path / "output" / "0" / "test_img" / "view0.png"), cv2.IMREAD_COLOR)
    dep = cv2.imread(str(tmp_path / "output" / "0" / "test_depth" / "view0.png"), cv2.IMREAD_UNCHANGED)
    fid, x, y, h, w = seq.objects[0].boxes[0]
    assert img is not None and img.shape == (h, w, 3) and dep.dtype == np.uint16 and dep.shape == (h, w)
    # the rendered view resembles the keyframe inside the object's mask (PSNR > 18 dB after 1000 iterations)
    gt = seq.rgb[fid][y:y + h, x:x + w, ::-1].astype(np.float32)
    m = (seq.instance[fid][y:y + h, x:x + w] == seq.objects[0].instance_id) & (dep > 0)
    assert m.mean() > 0.1
    mse = ((img.astype(np.float32) - gt)[m] ** 2).mean() / 255.0 ** 2
    assert -10 * np.log10(mse) > 18.0, -10 * np.log10(mse)
    # GenerateMesh + SaveMesh: ./output/<id>.ply after training and obj.ply next to the test views, reference PLY layout
    for ply in (tmp_path / "output" / "0.ply", tmp_path / "output" / "0" / "obj.ply"):
        txt = ply.read_text().splitlines()
        assert txt[0] == "ply" and txt[1] == "format ascii 1.0"
        nv = int(next(l for l in txt if l.startswith("element vertex")).split()[-1])
        nf = int(next(l for l in txt if l.startswith("element face")).split()[-1])
        body = txt[txt.index("end_header") + 1:]
        assert nv > 20 and nf > 20 and len(body) == nv + nf
        v = np.array([l.split() for l in body[:nv]], dtype=np.float64)
        f = np.array([l.split() for l in body[nv:]], dtype=np.int64)
        assert v.shape[1] == 9 and (f[:, 0] == 3).all() and f[:, 1:].max() < nv
        half = 1.1 * seq.objects[0].half
        assert (np.abs(v[:, :3]) <= half + 1e-3).all()                       # vertices inside the object box
        # like the reference, the vertex array is padded to a multiple of 128 with unreferenced zero vertices (marching_cubes.cu:499)
        used = np.unique(f[:, 1:])
        assert nv % 128 == 0 and nv - len(used) < 128 and not v[np.setdiff1d(np.arange(nv), used), :6].any()
        # unit normals (3 printed decimals); a model trained on 8 views has a few degenerate slivers whose normal is ~0
        assert (np.abs(np.linalg.norm(v[used, 3:6], axis=1) - 1.0) <= 5e-3).mean() >= 0.98
        # the surface spans the object (orientation and manifoldness are checked on an analytic field below; a model
        # trained on 8 views keeps floaters at the box faces, so no orientation statistic here)
        assert (np.abs(v[:, :3]).max(0) > 0.25 * seq.objects[0].half).all()


@pytest.mark.gpu
def test_online_manager_replay(tmp_path, host_lib):
    """NerfManagerOnline driven like the SLAM frontend drives it (DatasetInit, NewFrameToDataset per keyframe with BGR
    cv::Mat-style buffers, CreateNeRF on first sight, UpdateNeRFBbox(train_step=1) per observation, WaitThreadsEnd,
    RenderNeRFsTest): objects start training once they have more than 10 boxes (nerf.cu:222) and end trained."""
    import cv2
    from ro_map_b200 import build, synthetic as syn
    seq = syn.make_sequence(n_frames=18, n_objects=2, seed=7, H=200, W=200, K=(277.7775, 277.7775, 100.0, 100.0))
    syn.write_sequence(seq, str(tmp_path / "seq"))
    cfg = ROOT / "ro_map_b200" / "configs" / "base.json"
    p = subprocess.run([str(build.REPLAY_BIN), str(cfg), str(tmp_path / "seq"), "1", "100", "2", str(tmp_path / "out")],
                       capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("object ")]
    assert len(lines) == 2, p.stdout
    for l, obj in zip(lines, seq.objects):
        tok = l.split()
        boxes, step, loss = int(tok[tok.index("boxes") + 1]), int(tok[tok.index("step") + 1]), float(tok[tok.index("loss") + 1])
        assert boxes == len(obj.boxes)
        # one Train_Step_Online(100) per box update after the 10th box, plus the final one (nerf.cu:222-246); box updates
        # that arrive while a step is running are merged, so the count is bounded rather than exact
        assert 100 <= step <= 100 * (boxes - 10 + 1) and step % 100 == 0, l
        assert np.isfinite(loss) and loss < 0.2, l
    assert "ingest_ms_per_keyframe" in p.stdout
    img = cv2.imread(str(tmp_path / "out" / "0" / "test_img" / "view0.png"), cv2.IMREAD_COLOR)
    assert img is not None and img.shape[2] == 3
    # the reference's full output set (nerf.cu:255-404): test.txt / train.txt schemas, 60-view turn-table video, obj.ply
    out0 = tmp_path / "out" / "0"
    test_lines = (out0 / "test.txt").read_text().splitlines()
    assert test_lines[0].startswith("#stamp  box.x  box.y  box.h  box.w  tx  ty  tz  qx  qy  qz  qw")
    tok = test_lines[1].split()
    first = seq.objects[0].boxes[0]
    assert tok[0] == "view0" and [int(v) for v in tok[1:5]] == [int(v) for v in first[1:5]] and len(tok) == 12
    q = np.array(tok[8:12], dtype=np.float64)
    assert np.isclose(np.linalg.norm(q), 1.0, atol=1e-4)
    # object-centric camera position = ObjTow * Twc translation
    Toc = np.asarray(seq.objects[0].Tow, dtype=np.float64).reshape(4, 4) @ np.asarray(seq.poses[first[0]], dtype=np.float64).reshape(4, 4)
    assert np.allclose(np.array(tok[5:8], dtype=np.float64), Toc[:3, 3], atol=3e-4)
    train_lines = (out0 / "train.txt").read_text().splitlines()
    assert train_lines[0] == "#class Bbox" and len(train_lines[1].split()) == 4
    assert len(train_lines) == 3 + len(seq.objects[0].boxes) and all(len(l.split()) == 12 for l in train_lines[3:])
    vids = sorted((out0 / "video_img").glob("*.png"), key=lambda p: int(p.stem))
    assert [int(p.stem) for p in vids] == list(range(60)) and len(list((out0 / "video_depth").glob("*.png"))) == 60
    v0 = cv2.imread(str(vids[0]), cv2.IMREAD_COLOR)
    d0 = cv2.imread(str(out0 / "video_depth" / "0.png"), cv2.IMREAD_UNCHANGED)
    assert v0.shape == (seq.H // 2, seq.W // 2, 3) and d0.dtype == np.uint16 and d0.shape == v0.shape[:2]
    # the turn-table camera looks at the object: the centre of (nearly) every view is covered (depth > 0)
    cover = [cv2.imread(str(out0 / "video_depth" / f"{i}.png"), cv2.IMREAD_UNCHANGED)[seq.H // 4, seq.W // 4] > 0 for i in range(0, 60, 6)]
    assert sum(cover) >= 6, cover
    assert (out0 / "obj.ply").read_text().startswith("ply")


def test_mesh_extraction_on_analytic_field(tmp_path):
    """ro_map_b200/host/mesh.h (GenerateMesh/TransCPUMesh/SaveMesh replacement) on an analytic sphere field, the two
    C-ABI calls stubbed (tests/host/mesh_check.cpp): closed 2-manifold with consistent winding (Euler characteristic 2),
    vertices on the iso-surface, outward 1-ring normals, colours = logistic(rgb logits), reference PLY layout and vertex padding;
    then white noise (every marching-cubes configuration): consistent winding everywhere; the classic table the reference uses is
    known to leave cracks on faces whose corners alternate, identically in the reference (golden comparison: test_golden_romap.py)."""
    exe = tmp_path / "mesh_check"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT / 'ro_map_b200' / 'host'}", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tests' / 'host'}",
                    str(ROOT / "tests" / "host" / "mesh_check.cpp"), "-o", str(exe)], check=True)
    for res, r_tol in ((64, 1e-3), (17, 1e-2)):
        ply = tmp_path / f"s{res}.ply"
        out = subprocess.run([str(exe), str(res), str(ply)], capture_output=True, text=True, check=True).stdout.split()
        fact = {out[i]: float(out[i + 1]) for i in range(0, len(out), 2)}
        assert fact["verts"] > 100 and fact["bad_edges"] == 0 and fact["euler"] == 2, fact
        assert fact["faces"] == 2 * fact["verts"] - 4                           # closed triangle mesh of genus 0
        assert fact["max_r_err"] < r_tol and fact["min_normal_dot"] > 0.95 and fact["bad_colors"] == 0, fact
        # the reference's winding (geometric normals into the dense side) encloses the sphere's volume with a negative sign;
        # the table respects the cube's rotations
        assert abs(-fact["volume"] - 4.0 / 3.0 * np.pi * 0.3 ** 3) < (2e-3 if res == 64 else 4e-3), fact
        assert fact["asymmetric_cases"] == 0
        txt = ply.read_text().splitlines()
        hdr = txt[:txt.index("end_header") + 1]
        assert hdr[0] == "ply" and hdr[1] == "format ascii 1.0" and hdr[-2] == "property list uchar int vertex_index"
        assert [l for l in hdr if l.startswith("property")][:9] == [f"property float {c}" for c in ("x", "y", "z", "nx", "ny", "nz")] + \
            [f"property uchar {c}" for c in ("red", "green", "blue")]
        assert len(txt) - len(hdr) == int(fact["padded_verts"] + fact["faces"])
        # the reference pads the vertex array to a multiple of 128 with zero vertices (marching_cubes.cu:499); every configuration has triangles
        assert fact["bad_padding"] == 0 and fact["unreferenced"] == 0 and fact["cases"] == 254, fact
    # white noise with an empty border: all 256 cell configurations incl. faces whose corners alternate.  The published table
    # resolves such a face differently for a configuration and its complement, so a few percent of the edges are open there
    # (exactly as in the reference's mesh: the triangle sets are identical, tests/test_golden_romap.py); every vertex is used
    out = subprocess.run([str(exe), "40", str(tmp_path / "noise.ply"), "random"], capture_output=True, text=True, check=True).stdout.split()
    fact = {out[i]: float(out[i + 1]) for i in range(0, len(out), 2)}
    assert fact["verts"] > 50000 and fact["bad_padding"] == 0 and fact["unreferenced"] == 0, fact
    assert fact["bad_edges"] <= 0.08 * fact["edges"], fact


def test_pose_math_turntable_and_quaternion(tmp_path):
    """ro_map_b200/host/pose_math.h: GenerateToc's turn-table poses (nerf_model.cu:2186-2205) are rigid, sit at
    (r, theta, phi) and look at the object origin; rotation -> quaternion agrees with scipy on both branches (trace > 0
    and largest-diagonal); Toc = ObjTow * Twc composes column-major like Eigen."""
    from scipy.spatial.transform import Rotation
    exe = tmp_path / "pose_check"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT / 'ro_map_b200' / 'host'}", str(ROOT / "tests" / "host" / "pose_check.cpp"), "-o", str(exe)], check=True)
    run = lambda *a: np.array(subprocess.run([str(exe), *map(str, a)], capture_output=True, text=True, check=True).stdout.split(), dtype=np.float64)
    for theta in (6.0, 90.0, 186.0, 360.0):
        T = run("toc", theta, 30.0, 1.5).reshape(4, 4).T          # printed column-major
        Rm, t = T[:3, :3], T[:3, 3]
        assert np.allclose(Rm.T @ Rm, np.eye(3), atol=1e-6) and np.isclose(np.linalg.det(Rm), 1.0, atol=1e-6)
        assert np.isclose(np.linalg.norm(t), 1.5, atol=1e-6) and np.isclose(t[2], 1.5 * np.sin(np.pi / 6), atol=1e-6)
        assert np.allclose(np.arctan2(t[1], t[0]) % (2 * np.pi), np.deg2rad(theta) % (2 * np.pi), atol=1e-5) or np.isclose(theta, 360.0)
        assert np.allclose(Rm[:, 2], -t / np.linalg.norm(t), atol=1e-6)       # optical axis points at the origin
        assert abs(Rm[2, 0]) < 1e-7                                            # camera x axis stays horizontal
        assert np.allclose(T[3], [0, 0, 0, 1])
    rng = np.random.default_rng(3)
    rots = list(Rotation.random(20, random_state=4)) + [Rotation.from_euler("xyz", e) for e in ((3.1, 0.01, 0.02), (0.01, 3.1, 0.0), (0.02, 0.0, 3.1))]
    for r in rots:
        T = np.eye(4); T[:3, :3] = r.as_matrix(); T[:3, 3] = rng.normal(size=3)
        q = run("quat", *T.T.reshape(-1))
        ref = r.as_quat()
        assert min(np.abs(q - ref).max(), np.abs(q + ref).max()) < 2e-6, (q, ref)
        B = np.eye(4); B[:3, :3] = Rotation.random(random_state=5).as_matrix(); B[:3, 3] = (0.3, -0.2, 0.9)
        C = run("mul", *T.T.reshape(-1), *B.T.reshape(-1)).reshape(4, 4).T
        assert np.allclose(C, T @ B, atol=1e-5)


def test_turntable_poses_against_reference_generate_toc(tmp_path):
    """ro_map_b200/host/pose_math.h turntable_toc against the reference's own NeRF_Model::GenerateToc (nerf_model.cu:2186-2205) run by
    oracle/ref/make_golden_toc.py: all 60 poses of RenderVideo's turn-table (theta accumulated in float like :1841) at three radii.
    Both sides evaluate sin / cos of float angles with the host libm and normalise in float: equal to a few float ulps."""
    gold = np.load(ROOT / "tests" / "golden" / "romap_toc_golden.npz")
    exe = tmp_path / "pose_check"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT / 'ro_map_b200' / 'host'}", str(ROOT / "tests" / "host" / "pose_check.cpp"), "-o", str(exe)], check=True)
    worst = 0.0
    for i, r in enumerate(gold["radii"]):
        for j, th in enumerate(gold["thetas"]):
            out = subprocess.run([str(exe), "toc", repr(float(th)), repr(float(gold["phi"])), repr(float(r))], capture_output=True, text=True, check=True).stdout.split()
            T = np.array(out, dtype=np.float64)
            worst = max(worst, float(np.abs(T - gold["toc"][i, j]).max()))
    assert worst <= 4e-7 * float(gold["radii"].max()), worst


def test_draw_cpu_mesh_issues_the_reference_gl_calls(tmp_path):
    """nerf::NeRF::DrawCPUMesh (nerf.cu:484-507) against a recording stand-in for <GL/gl.h> (tests/host/gl_stub): client arrays for
    positions / normals / u8 colours, one glDrawElements over the index list; nothing is drawn before a mesh exists or while the
    training thread holds the mesh (try-lock).  The shipped libMON.so of this image is headless (no GL header -> no-op)."""
    from ro_map_b200 import build
    build.build()
    exe = tmp_path / "draw_check"
    subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", f"-I{ROOT / 'tests' / 'host' / 'gl_stub'}", f"-I{ROOT / 'include'}", f"-I{ROOT / 'ro_map_b200' / 'host'}",
                    str(ROOT / "ro_map_b200" / "host" / "nerf_host.cpp"), str(ROOT / "tests" / "host" / "draw_check.cpp"),
                    f"-L{ROOT / 'ro_map_b200'}", "-lmon_b200", "-lz", f"-Wl,-rpath,{ROOT / 'ro_map_b200'}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("-- ")
    sect = {s.splitlines()[0]: s.splitlines()[1:] for s in out if s.strip()}
    drawn = ["enable 8074", "enable 8075", "enable 8076", "vertex 3 1406 0 verts", "normal 1406 0 normals", "color 3 1401 0 colors",
             "draw 4 3 1405 indices", "disable 8074", "disable 8075", "disable 8076"]
    assert sect["no mesh yet"] == [] and sect["mesh being updated by the training thread"] == []
    assert sect["mesh present"] == drawn and sect["manager entry point"] == drawn
    # the library built for this image has no OpenGL dependency
    needed = subprocess.run(["readelf", "-d", str(ROOT / "ro_map_b200" / "libMON.so")], capture_output=True, text=True).stdout
    assert "libGL" not in needed
