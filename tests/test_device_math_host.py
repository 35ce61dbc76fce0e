"""The kernels' shared arithmetic header (ro_map_b200/csrc/mon_device.cuh) compiled for the HOST and held against the CPU
oracle bit for bit — and, for the rays of the render window of tests/golden/romap_golden.npz, against what the reference's
own GenerateRenderRays produced on a B200 — bit for bit as well.

The header writes every rounding-relevant expression with explicit _rn intrinsics (IEEE single operations that nvcc neither
fuses nor reorders); tests/host/device_math_check.cpp restates those intrinsics one to one for g++ (-ffp-contract=off), so a
change of the operation order in the header shows up here, on the CPU, before any GPU run.  The GPU tests then only have to
confirm that the device executes the same IEEE operations."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle" / "ref"))
import make_golden_romap as mg  # noqa: E402
import test_golden_romap as tg  # noqa: E402


def _hex(a):
    return [f"{int(v):08x}" for v in np.ascontiguousarray(a, np.float32).reshape(-1).view(np.uint32)]


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("devmath") / "device_math_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-attributes", f"-I{ROOT / 'ro_map_b200' / 'csrc'}", f"-I{ROOT / 'include'}",
                    "-I/usr/local/cuda/include", str(ROOT / "tests" / "host" / "device_math_check.cpp"), "-o", str(out)], check=True)
    return out


@pytest.mark.parametrize("tag", tg.TAGS)
def test_device_header_matches_oracle_and_reference(exe, oracle, tag):
    gold = np.load(tg.GOLD)
    seq = mg.make_scene()
    k = next(c[1] for c in mg.CASES if c[0] == tag)
    obj = seq.objects[k]
    bmin, bmax = (-1.1 * obj.half).astype(np.float32), (1.1 * obj.half).astype(np.float32)
    fid, x, y, h, w = [int(v) for v in gold[tag + "r_box"]]
    box = (fid, max(0, x - 3), max(0, y - 3), h + 6, w + 6)                       # a little larger than the golden window
    Twc, Tow = np.asarray(seq.poses[fid], np.float32), np.asarray(obj.Tow, np.float32)
    args = [str(v) for v in box[1:]] + _hex(seq.K) + _hex(Twc.T) + _hex(Tow.T) + _hex(bmin) + _hex(bmax)   # .T: column-major
    lines = subprocess.run([str(exe), *args], capture_output=True, text=True, check=True).stdout.splitlines()
    n = box[3] * box[4]
    assert len(lines) == n
    hit = np.array([int(l.split()[0]) for l in lines], np.int32)
    rays = np.array([[int(t, 16) for t in l.split()[1:10]] for l in lines], np.uint32).view(np.float32)

    # 1. rays and in-box flags: bit-identical to the oracle's GenerateRenderRays restatement
    rays_o, hit_o = oracle.render_rays(box, seq.poses[fid], seq.K, obj.Tow, bmin, bmax)
    assert np.array_equal(hit, hit_o) and 0 < hit.sum() < n
    assert np.array_equal(rays[hit == 1].view(np.uint32), rays_o[hit == 1].view(np.uint32))

    # 2. the golden window (the reference's own kernel on a B200) is a sub-window of this one
    sub = np.array([(yy - box[2]) * box[4] + (xx - box[1]) for yy in range(y, y + h) for xx in range(x, x + w)])
    g_hit, g_rays = gold[tag + "r_in_box"], gold[tag + "r_rays"]
    assert np.array_equal(hit[sub], g_hit)
    assert tg.bits_equal(rays[sub][g_hit == 1], g_rays[g_hit == 1])               # bit-identical to the reference's kernel

    # 3. sampling, unit-cube warp, grid cell / fraction and hash index of the hit rays: the oracle's A3 / A4 restatements
    rest = np.array([[int(t, 16) for t in l.split()[10:17]] for l, hh in zip(lines, hit) if hh], np.uint32).view(np.float32)
    index = np.array([int(l.split()[17]) for l, hh in zip(lines, hit) if hh], np.uint32)
    dt = np.full((int(hit.sum()), 64), 0.625, np.float32)
    pts_o, t_o = oracle.sample_points(rays_o[hit == 1], 64, bmin, bmax, dt)
    pts_o, t_o = pts_o.reshape(-1, 64, 3)[:, 5], t_o.reshape(-1, 64)[:, 5]
    assert np.array_equal(rest[:, 0].view(np.uint32), t_o.view(np.uint32))
    assert np.array_equal(rest[:, 1:4].view(np.uint32), np.ascontiguousarray(pts_o).view(np.uint32))
    idx_o, w_o = oracle.encode_corners(oracle.default_config(), pts_o)
    assert np.array_equal(index, idx_o[:, 9, 5])                                  # corner (x+1, y, z+1) of level 9 (hashed, scale 8191)
    fr = rest[:, 4:7]
    w7 = ((fr[:, 0] * fr[:, 1]).astype(np.float32) * fr[:, 2]).astype(np.float32)  # weight of corner (1,1,1): fx * fy * fz in this order
    assert np.array_equal(w7.view(np.uint32), np.ascontiguousarray(w_o[:, 9, 7]).view(np.uint32))
