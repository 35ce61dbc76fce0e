"""Generates tests/golden/romap_golden.npz and romap_mesh_golden.npz by RUNNING RO-MAP's own glue kernels on a GPU: oracle/_ref/libmon_ref.so built
with ROMAP_GENUINE compiles the reference's nerf_model.cu from where it lies (unmodified), so GenerateRays,
fill_rollover_rays, GenerateInputPoints, VolumeRender, VolumeRenderGradient_No_Compacted, SumLoss, GenerateRenderRays,
GenerateRenderInputPoints and VolumeRender_Render below are the reference's kernels, fed by the reference's tiny-cuda-nn.

    python oracle/ref/make_golden_romap.py [out.npz] [--mesh-only]          # needs a CUDA device; run under gpurun

Rows A1-A3, A6, A7, A14 of SURVEY.md section 8a.  Every stage's INPUT is stored next to its OUTPUT (the network output that
feeds the compositing is the reference's own), so tests/test_golden_romap.py can hold each oracle stage against the
reference in isolation, on the CPU.  The scene (tests/conftest.py's `small_seq`) and the injected random numbers are
regenerated bit-for-bit by the test from the seeds below; a SHA-256 of the scene guards that.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

SMALL = dict(H=200, W=200, K=(277.7775, 277.7775, 100.0, 100.0))   # = tests/conftest.py
S = 32                       # training samples per ray (nerf_model.h: mnSampleNum)
S2 = 64                      # render samples per ray (mnRenderSampleNum)
CASES = [                    # (tag, object index, rays per batch, use depth, warm-up iterations, seed of the injected randoms)
    ("a_", 0, 256, True, 600, 4242),
    ("b_", 1, 128, False, 40, 4243),
]
RENDER_HW = (12, 16)         # rendered window (h, w), straddling the left edge of the object's 2-D box


def make_scene():
    from ro_map_b200 import synthetic as syn
    return syn.make_sequence(n_frames=6, n_objects=2, seed=1337, **SMALL)


def scene_sha(seq) -> str:
    h = hashlib.sha256()
    for a in list(seq.rgb) + list(seq.instance) + list(seq.depth) + list(seq.poses):
        h.update(np.ascontiguousarray(a).tobytes())
    for o in seq.objects:
        h.update(np.ascontiguousarray(o.Tow, np.float32).tobytes() + np.ascontiguousarray(o.half, np.float32).tobytes())
        h.update(np.array(o.boxes, np.int64).tobytes())
    return h.hexdigest()


def injected(seed: int, R: int):
    """(sample_xy [R,2], rand_colors [R,3], rand_dt [R,S]) in (0,1], like curandGenerateUniform."""
    rng = np.random.default_rng(seed)
    return tuple((1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32) for shape in ((R, 2), (R, 3), (R, S)))


def render_window(obj):
    """(FrameId, x, y, h, w): RENDER_HW pixels at mid height of the object's first 2-D box, starting 5 px left of it."""
    fid, x, y, h, w = [int(v) for v in obj.boxes[0]]
    return (fid, max(0, x - 5), y + h // 2, RENDER_HW[0], RENDER_HW[1])


def render_dt(seed: int, n_rays: int):
    rng = np.random.default_rng(seed + 1000)
    return (1.0 - rng.random((n_rays, S2), dtype=np.float32)).astype(np.float32)


MESH_BOX = (np.array([-1.0, -2.0, -0.5], np.float32), np.array([1.0, 2.0, 0.5], np.float32))   # object box of the mesh cases
MESH_CASES = [("sphere", 24), ("noise", 14), ("cases", 22)]


def mesh_lattice(kind: str, res: int) -> np.ndarray:
    """[z][y][x] density lattice: an off-centre sphere of radius 0.3 (unit-cube metric) around the threshold 2.0, white
    noise in [0,4) with an empty border (ambiguous faces included; 252 of the 254 surface configurations occur), or "cases":
    each of the 256 cell configurations once as an isolated cell (7 x 7 x 6 cells, one empty lattice plane between them)."""
    if kind == "cases":
        d = np.full((res, res, res), 1.0, np.float32)
        corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]   # mask bit -> (dx, dy, dz)
        rng = np.random.default_rng(778)
        for mask in range(256):
            cx, cy, cz = 3 * (mask % 7), 3 * ((mask // 7) % 7), 3 * (mask // 49)
            for b, (dx, dy, dz) in enumerate(corners):
                if (mask >> b) & 1:
                    d[cz + dz, cy + dy, cx + dx] = np.float32(2.25 + 1.5 * rng.random())     # off-centre crossings
        return d
    if kind == "sphere":
        ax = (np.arange(res, dtype=np.float32) / np.float32(res - 1)).astype(np.float32)
        z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
        r = np.sqrt((x - np.float32(0.5)) ** 2 + (y - np.float32(0.45)) ** 2 + (z - np.float32(0.55)) ** 2, dtype=np.float32)
        return (np.float32(2.0) + np.float32(40.0) * (np.float32(0.3) - r)).astype(np.float32)
    rng = np.random.default_rng(777)
    d = np.zeros((res, res, res), np.float32)
    d[1:-1, 1:-1, 1:-1] = (4.0 * rng.random((res - 2,) * 3, dtype=np.float32)).astype(np.float32)
    return d


TRAJ = dict(obj=0, R=256, iters=30, seed=9000, stride=997)   # trajectory case: Train_Step's loop for 30 iterations from the initial weights


def traj_randoms(it: int, seq, obj, bmin, bmax):
    """Injected randoms of trajectory iteration `it`, made independent of the reference's slot race: one background colour and
    one jitter row for all slots, and pixels re-drawn (oracle as the judge) until every slot's ray survives, so that the batch
    needs no roll-over padding.  Deterministic in (seed, it, scene)."""
    from oracle import mon_oracle as orc
    R = TRAJ["R"]
    rng = np.random.default_rng(TRAJ["seed"] + it)
    u = lambda shape: (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)   # noqa: E731
    col = np.repeat(u((1, 3)), R, axis=0)
    dt = np.repeat(u((1, S)), R, axis=0)
    frames = orc.Frames(seq.rgb, seq.instance, seq.depth, seq.poses)
    sxy = np.zeros((R, 2), np.float32)
    for i in range(R):
        box = obj.boxes[i % len(obj.boxes)]
        for _ in range(500):
            cand = u((1, 2))
            if orc.generate_rays(1, [box], frames, seq.H, seq.W, seq.K, obj.Tow, bmin, bmax, obj.instance_id, True, cand, col[:1])[0] == 1:
                break
        else:
            raise RuntimeError("no surviving pixel")
        sxy[i] = cand[0]
    return sxy, col, dt


def make_traj(out_path: str):
    from ref_binding import RefLib, RefModel
    seq = make_scene()
    obj = seq.objects[TRAJ["obj"]]
    bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
    m = RefModel(1, 1337, RefLib())
    assert m.is_genuine()
    m.scene(seq.rgb, seq.instance, seq.depth, seq.poses, seq.H, seq.W, seq.K, obj.boxes, obj.Tow, bmin, bmax, obj.instance_id, True, TRAJ["R"])
    losses = []
    for it in range(TRAJ["iters"]):
        _, _, loss, n_in = m.train(1, traj_randoms(it, seq, obj, bmin, bmax))
        assert n_in == TRAJ["R"], (it, n_in)
        losses.append(loss)
    gold = {"scene_sha256": np.array(scene_sha(seq)), "loss": np.array(losses, np.float32),
            "master_sample": m.get(0)[:: TRAJ["stride"]].copy(), "ema_sample": m.get(2)[:: TRAJ["stride"]].copy(), "master_mlp": m.get(0)[:3072].copy()}
    m.close()
    np.savez_compressed(out_path, **gold)
    print("trajectory", [round(float(v), 5) for v in losses[:3]], "...", [round(float(v), 5) for v in losses[-3:]], "wrote", out_path, flush=True)


def f16_bits(a: np.ndarray) -> np.ndarray:
    h = a.astype(np.float16)
    assert np.array_equal(h.astype(np.float32), a), "value is not fp16-representable"
    return h.view(np.uint16)


def main(out_path: str, mesh_only: bool = False):
    from ref_binding import RefLib, RefModel
    lib = RefLib()
    seq = make_scene()
    gold = {"scene_sha256": np.array(scene_sha(seq))}
    for tag, k, R, use_depth, warm, seed in ([] if mesh_only else CASES):
        obj = seq.objects[k]
        bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
        m = RefModel(1, 1337, lib)
        assert m.is_genuine(), "libmon_ref.so was built without ROMAP_GENUINE"
        m.scene(seq.rgb, seq.instance, seq.depth, seq.poses, seq.H, seq.W, seq.K, obj.boxes, obj.Tow, bmin, bmax, obj.instance_id, use_depth, R)
        m.train(warm)                                         # cuRAND-driven warm-up so that densities and colours are not flat
        sxy, col, dt = injected(seed, R)
        _, _, loss, n_in = m.train(1, (sxy, col, dt))
        N = R * S
        gold[tag + "n_in"] = np.array(n_in)
        gold[tag + "loss"] = np.array(loss, np.float32)
        gold[tag + "rays"] = m.last(0, R * 9).reshape(R, 9)
        gold[tag + "ray_instance"] = m.last(12, R).astype(np.uint8)
        gold[tag + "target"] = m.last(10, R * 3).reshape(R, 3)
        gold[tag + "target_depth"] = m.last(11, R)
        gold[tag + "points"] = m.last(1, N * 3).reshape(N, 3)
        gold[tag + "dist"] = m.last(2, N)
        out = m.last(4, N * 16).reshape(N, 16)
        gold[tag + "out_bits"] = f16_bits(out)                # all 16 padded columns: the reference's kernels index with the padded width
        gold[tag + "rgb_rays"] = m.last(5, R * 3).reshape(R, 3)
        gold[tag + "depth_rays"] = m.last(6, R)
        gold[tag + "mask_rays"] = m.last(7, R)
        dout = m.last(8, N * 16).reshape(N, 16)
        assert not dout[:, 4:].any(), "dL/dout of the padding columns must stay zero"
        gold[tag + "dout_bits"] = f16_bits(dout[:, :4])
        gold[tag + "loss_rays"] = m.last(13, R)
        # Render (EMA weights) of a small window
        box = render_window(obj)
        n_rays = box[3] * box[4]
        rdt = render_dt(seed, n_rays)
        r = m.render(box, seq.poses[box[0]], rdt)
        gold[tag + "r_box"] = np.array(box, np.uint32)
        gold[tag + "r_rays"] = r["rays"]
        gold[tag + "r_in_box"] = r["in_box"]
        gold[tag + "r_points"] = r["points"]
        gold[tag + "r_dist"] = r["dist"]
        if np.array_equal(r["out4"].astype(np.float16).astype(np.float32), r["out4"]):
            gold[tag + "r_out4_bits"] = f16_bits(r["out4"])   # tiny-cuda-nn computes in fp16 and widens: store the bits
        else:
            gold[tag + "r_out4_f32"] = r["out4"]
        gold[tag + "r_rgb"], gold[tag + "r_depth"], gold[tag + "r_mask"] = r["rgb"], r["depth"], r["mask"]
        print(tag, "n_in", n_in, "loss", loss, "opaque training rays", int((gold[tag + "mask_rays"] > 0.5).sum()), "render hits", int(r["in_box"].sum()),
              "render opaque", int(r["mask"].sum()), "early-stopped rays", int((np.abs(dout[:, :4]).reshape(R, S, 4).sum(-1)[:, -1] == 0).sum()), flush=True)
        m.close()
        np.savez_compressed(out_path, **gold)                  # after every case: a later failure keeps the earlier cases
    if not mesh_only:
        print("wrote", out_path, Path(out_path).stat().st_size, "bytes")
    # the reference's marching cubes (marching_cubes.cu, compiled in place) on two lattices -> a second, small fixture
    mesh = {}
    m = RefModel(1, 1337, lib)
    obj = seq.objects[0]
    m.scene(seq.rgb[:1], seq.instance[:1], seq.depth[:1], seq.poses[:1], seq.H, seq.W, seq.K, obj.boxes[:1], obj.Tow, MESH_BOX[0], MESH_BOX[1], obj.instance_id, True, 128)
    for kind, res in MESH_CASES:
        v, n, idx = m.marching_cubes(mesh_lattice(kind, res), 2.0)
        mesh[kind + "_verts"], mesh[kind + "_normals"], mesh[kind + "_indices"] = v, n, idx
        print("mesh", kind, res, "verts (padded)", len(v), "referenced", len(np.unique(idx)), "triangles", len(idx) // 3, flush=True)
    m.close()
    mesh_path = str(Path(out_path).with_name("romap_mesh_golden.npz"))
    np.savez_compressed(mesh_path, **mesh)
    print("wrote", mesh_path, Path(mesh_path).stat().st_size, "bytes")


if __name__ == "__main__":
    if "--traj" in sys.argv:
        make_traj(str(ROOT / "gpurun_out" / "romap_traj_golden.npz"))
        sys.exit(0)
    args = [a for a in sys.argv[1:] if a != "--mesh-only"]
    main(args[0] if args else str(ROOT / "gpurun_out" / "romap_golden.npz"), mesh_only="--mesh-only" in sys.argv)
