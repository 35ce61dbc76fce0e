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
        assert nv > 100 and nf > 100 and len(body) == nv + nf
        v = np.array([l.split() for l in body[:nv]], dtype=np.float64)
        f = np.array([l.split() for l in body[nv:]], dtype=np.int64)
        assert v.shape[1] == 9 and (f[:, 0] == 3).all() and f[:, 1:].max() < nv
        half = 1.1 * seq.objects[0].half
        assert (np.abs(v[:, :3]) <= half + 1e-4).all()                       # vertices inside the object box
        assert np.allclose(np.linalg.norm(v[:, 3:6], axis=1), 1.0, atol=5e-3)
        # the surface spans the object (orientation and manifoldness are checked on an analytic field below; a model
        # trained on 8 views keeps floaters at the box faces, so no orientation statistic here)
        assert (np.abs(v[:, :3]).max(0) > 0.6 * seq.objects[0].half).all()


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


def test_mesh_extraction_on_analytic_field(tmp_path):
    """ro_map_b200/host/mesh.h (GenerateMesh/TransCPUMesh/SaveMesh replacement) on an analytic sphere field, the two
    C-ABI calls stubbed (tests/host/mesh_check.cpp): closed 2-manifold with consistent winding (Euler characteristic 2),
    vertices on the iso-surface, outward 1-ring normals, colours = logistic(rgb logits), reference PLY layout."""
    exe = tmp_path / "mesh_check"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT / 'ro_map_b200' / 'host'}", f"-I{ROOT / 'include'}",
                    str(ROOT / "tests" / "host" / "mesh_check.cpp"), "-o", str(exe)], check=True)
    for res, r_tol in ((64, 1e-3), (17, 1e-2)):
        ply = tmp_path / f"s{res}.ply"
        out = subprocess.run([str(exe), str(res), str(ply)], capture_output=True, text=True, check=True).stdout.split()
        fact = {out[i]: float(out[i + 1]) for i in range(0, len(out), 2)}
        assert fact["verts"] > 100 and fact["bad_edges"] == 0 and fact["euler"] == 2, fact
        assert fact["faces"] == 2 * fact["verts"] - 4                           # closed triangle mesh of genus 0
        assert fact["max_r_err"] < r_tol and fact["min_normal_dot"] > 0.95 and fact["bad_colors"] == 0, fact
        txt = ply.read_text().splitlines()
        hdr = txt[:txt.index("end_header") + 1]
        assert hdr[0] == "ply" and hdr[1] == "format ascii 1.0" and hdr[-2] == "property list uchar int vertex_index"
        assert [l for l in hdr if l.startswith("property")][:9] == [f"property float {c}" for c in ("x", "y", "z", "nx", "ny", "nz")] + \
            [f"property uchar {c}" for c in ("red", "green", "blue")]
        assert len(txt) - len(hdr) == int(fact["verts"] + fact["faces"])
