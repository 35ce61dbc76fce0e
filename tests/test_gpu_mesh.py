"""GPU marching cubes (ro_map_b200/csrc/kernels_mesh.cu, mon_object_extract_mesh / mon_mesh_from_lattice) against its CPU
statement (tests/host/mesh_cpu.h) element for element, and against the reference's own marching_cubes.cu output
(tests/golden/romap_mesh_golden.npz, the assertions of tests/test_golden_romap.py)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import test_golden_romap as tg  # noqa: E402

mg = tg.mg
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def core():
    from ro_map_b200 import build, core
    build.build()
    if core.device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return core


def _cpu_mesh(tmp_path, name, lattice, lo, hi, thresh):
    import subprocess
    exe = tmp_path / "mesh_golden"
    if not exe.exists():
        subprocess.run(["g++", "-O2", "-std=c++17", f"-I{tg.ROOT / 'ro_map_b200' / 'host'}", f"-I{tg.ROOT / 'include'}", f"-I{tg.ROOT / 'tests' / 'host'}",
                        str(tg.ROOT / "tests" / "host" / "mesh_golden.cpp"), "-o", str(exe)], check=True)
    lat = tmp_path / f"{name}.f32"
    np.ascontiguousarray(lattice, np.float32).tofile(lat)
    out = subprocess.run([str(exe), str(lat), str(lattice.shape[0]), repr(float(thresh)), *[repr(float(v)) for v in lo], *[repr(float(v)) for v in hi],
                          str(tmp_path / name)], capture_output=True, text=True, check=True).stdout.split()
    return (int(out[0]), np.fromfile(str(tmp_path / name) + ".verts", np.float32).reshape(-1, 3),
            np.fromfile(str(tmp_path / name) + ".normals", np.float32).reshape(-1, 3), np.fromfile(str(tmp_path / name) + ".indices", np.uint32))


@pytest.mark.parametrize("kind,res", mg.MESH_CASES)
def test_gpu_marching_cubes_on_the_reference_lattices(core, tmp_path, kind, res):
    """a sphere, white noise (ambiguous faces) and every one of the 256 cell configurations: vertices and indices of the GPU
    kernels equal the CPU statement's element for element (both use the lattice order), unit normals up to the order of the float
    atomics; and the reference's own output (vertex sets bit-identical, triangles identical incl. winding and per-cell order)."""
    lat = mg.mesh_lattice(kind, res)
    lo, hi = mg.MESH_BOX
    m = core.mesh_from_lattice(lat, lo, hi, 2.0)
    n_surface, v, n, idx = _cpu_mesh(tmp_path, kind, lat, lo, hi, 2.0)
    assert m["n_surface"] == n_surface and m["verts"].shape == v.shape and len(m["verts"]) % 128 == 0
    assert np.array_equal(m["verts"], v)
    assert np.array_equal(m["indices"], idx)
    assert np.abs(m["normals"] - n).max() < 1e-4
    assert not m["colors"].any()
    tg.check_mesh_against_reference(kind, m["n_surface"], m["verts"], m["normals"], m["indices"])


def test_gpu_marching_cubes_sizes(core):
    """empty and full lattices (no surface), the smallest lattice, and one whose size is not a multiple of the scan's block"""
    lo, hi = np.array([-1, -1, -1], np.float32), np.array([1, 1, 1], np.float32)
    for fill in (0.0, 5.0):
        m = core.mesh_from_lattice(np.full((9, 9, 9), fill, np.float32), lo, hi, 2.0)
        assert m["n_surface"] == 0 and len(m["indices"]) == 0 and len(m["verts"]) == 0
    two = np.zeros((2, 2, 2), np.float32)
    two[0, 0, 0] = 4.0
    m = core.mesh_from_lattice(two, lo, hi, 2.0)
    assert m["n_surface"] == 3 and len(m["indices"]) == 3 and len(m["verts"]) == 128
    assert np.allclose(np.sort(np.abs(m["verts"][:3]).sum(axis=1)), [2.0, 2.0, 2.0])     # three edge midpoints next to the corner (-1,-1,-1)
    rng = np.random.default_rng(5)
    lat = rng.uniform(0, 4, (37, 37, 37)).astype(np.float32)
    m = core.mesh_from_lattice(lat, lo, hi, 2.0)
    assert m["n_surface"] > 30000 and m["indices"].max() == m["n_surface"] - 1 and len(np.unique(m["indices"])) == m["n_surface"]


def test_object_mesh_matches_the_cpu_statement(core, small_seq, tmp_path):
    """GenerateMesh on a trained object: mon_object_extract_mesh (lattice, marching cubes, normals, colours on the GPU) against the
    CPU statement fed with the same density lattice (mon_object_density_grid) and the same network queries."""
    seq, obj = small_seq, small_seq.objects[0]
    ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
    for i in range(len(seq.poses)):
        ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
    bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
    g = core.NerfObject(ds, core.default_config(rays_per_batch=1024), obj.Tow, bmin, bmax, obj.instance_id)
    g.set_bboxes(obj.boxes)
    g.train(600)
    res = 48
    sigma = g.density_grid((res, res, res))
    thresh = float(np.percentile(sigma, 85))
    m = g.extract_mesh(res, thresh)
    n_surface, v, n, idx = _cpu_mesh(tmp_path, "object", sigma, bmin, bmax, thresh)
    assert n_surface > 200
    assert m["n_surface"] == n_surface and np.array_equal(m["verts"], v) and np.array_equal(m["indices"], idx)
    assert np.abs(m["normals"] - n).max() < 1e-4
    # colours: logistic(rgb logits) * 255 truncated, at WarpPoint(vertex), padding vertices included
    unit = (m["verts"] - bmin.astype(np.float32)) / (bmax - bmin).astype(np.float32)
    rgb = 1.0 / (1.0 + np.exp(-g.query_points(unit)[:, :3].astype(np.float32)))
    want = np.clip(rgb * np.float32(255.0), 0, 255).astype(np.uint8)
    assert np.abs(m["colors"].astype(int) - want.astype(int)).max() <= 1
    assert (m["colors"] == want).mean() > 0.99
    g.close()
