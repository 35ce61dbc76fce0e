"""Synthetic stand-in for RO-MAP's 'room' sequence (the Drive-hosted original is not available offline).

Same camera model and on-disk schema as the reference reads:
800x800 frames, fx=fy=1111.11, cx=cy=400, camera-to-world poses on a hemisphere looking at the
origin, analytic objects (boxes, spheres, ellipsoids) standing on the z=0 floor, per-frame RGB u8,
u8 instance mask (0 = background, k = object class/instance id) and metric z-depth; per object the
object-to-world pose, half extents and the projected 2-D box (x, y, h, w) in every observing frame
(MON/Core/src/nerf.cu:58-118).  Everything is ray-cast with numpy; deterministic for a given seed.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

H = W = 800
FX = FY = 1111.11
CX = CY = 400.0


@dataclass
class SynthObject:
    instance_id: int            # also the "class" column of obj_offline/k.txt
    kind: str                   # "box" | "sphere" | "ellipsoid"
    Two: np.ndarray             # 4x4 object-to-world
    half: np.ndarray            # half extents a1 a2 a3 (object frame)
    color: np.ndarray           # base RGB in [0,1]
    boxes: list = field(default_factory=list)  # (FrameId, x, y, h, w)

    @property
    def Tow(self) -> np.ndarray:
        return np.linalg.inv(self.Two).astype(np.float32)


@dataclass
class SynthSequence:
    K: tuple
    H: int
    W: int
    poses: list                 # 4x4 camera-to-world per frame
    rgb: list                   # HxWx3 u8 per frame
    instance: list              # HxW u8
    depth: list                 # HxW f32 z-depth (0 where nothing was hit)
    objects: list


def _rotz(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def make_objects(n_objects: int, seed: int = 1337) -> list[SynthObject]:
    rng = np.random.default_rng(seed)
    kinds = ["box", "sphere", "ellipsoid"]
    palette = np.array([[0.85, 0.25, 0.2], [0.2, 0.6, 0.85], [0.3, 0.75, 0.35], [0.9, 0.75, 0.2],
                        [0.65, 0.35, 0.8], [0.95, 0.5, 0.15], [0.25, 0.8, 0.75], [0.8, 0.4, 0.55]])
    objs = []
    ring = 0.0 if n_objects == 1 else 0.75
    for k in range(n_objects):
        kind = kinds[k % 3]
        half = rng.uniform(0.12, 0.3, size=3)
        if kind == "sphere":
            half[:] = half[0]
        ang = 2 * math.pi * k / max(1, n_objects) + 0.3
        centre = np.array([ring * math.cos(ang), ring * math.sin(ang), half[2]])
        Two = np.eye(4)
        Two[:3, :3] = _rotz(rng.uniform(0, math.pi))
        Two[:3, 3] = centre
        objs.append(SynthObject(instance_id=k + 1, kind=kind, Two=Two.astype(np.float32), half=half.astype(np.float32),
                                color=palette[k % len(palette)].astype(np.float32)))
    return objs


def make_poses(n_frames: int, seed: int = 1337) -> list[np.ndarray]:
    """Camera-to-world, OpenCV axes (x right, y down, z forward), hemisphere r in [2,3], looking at the origin."""
    rng = np.random.default_rng(seed + 1)
    poses = []
    for i in range(n_frames):
        r = rng.uniform(2.0, 3.0)
        az = 2 * math.pi * (i / max(1, n_frames)) * 3.0 + rng.uniform(-0.05, 0.05)
        el = math.radians(rng.uniform(20.0, 65.0))
        eye = np.array([r * math.cos(el) * math.cos(az), r * math.cos(el) * math.sin(az), r * math.sin(el)])
        fwd = -eye / np.linalg.norm(eye)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        T = np.eye(4)
        T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, eye
        poses.append(T.astype(np.float32))
    return poses


def _hit_object(obj: SynthObject, o_w: np.ndarray, d_w: np.ndarray):
    """Returns (t, normal_obj) arrays; t = inf where missed. o_w [3], d_w [N,3] unit directions."""
    Tow = np.linalg.inv(obj.Two.astype(np.float64))
    o = Tow[:3, :3] @ o_w + Tow[:3, 3]
    d = d_w @ Tow[:3, :3].T
    a = obj.half.astype(np.float64)
    N = d.shape[0]
    t_hit = np.full(N, np.inf)
    nrm = np.zeros((N, 3))
    if obj.kind == "box":
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-a - o) / d
            t2 = (a - o) / d
        tn = np.minimum(t1, t2)
        tf = np.maximum(t1, t2)
        t0 = tn.max(axis=1)
        t1m = tf.min(axis=1)
        ok = (t0 <= t1m) & (t1m > 0)
        t = np.where(t0 > 0, t0, t1m)
        t_hit[ok] = t[ok]
        ax = tn.argmax(axis=1)
        nrm[np.arange(N), ax] = -np.sign(d[np.arange(N), ax])
    else:
        os_, ds_ = o / a, d / a
        A = (ds_ * ds_).sum(axis=1)
        B = 2 * (ds_ * os_).sum(axis=1)
        Cc = (os_ * os_).sum() - 1.0
        disc = B * B - 4 * A * Cc
        ok = disc >= 0
        sq = np.sqrt(np.where(ok, disc, 0))
        t = (-B - sq) / (2 * A)
        ok &= t > 0
        t_hit[ok] = t[ok]
        p = o + d * np.where(ok, t, 0)[:, None]
        n = p / (a * a)
        nrm = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
    return t_hit, nrm, o, d


def render_frame(objects: list[SynthObject], pose: np.ndarray, H: int = H, W: int = W, K=(FX, FY, CX, CY)):
    fx, fy, cx, cy = K
    ys, xs = np.mgrid[0:H, 0:W]
    dirs = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs, dtype=np.float64)], axis=-1).reshape(-1, 3)
    znorm = np.linalg.norm(dirs, axis=1)
    d_c = dirs / znorm[:, None]
    R, eye = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64)
    d_w = d_c @ R.T
    N = d_w.shape[0]
    best_t = np.full(N, np.inf)
    inst = np.zeros(N, np.uint8)
    rgb = np.zeros((N, 3))
    light = np.array([0.4, 0.3, 0.85])
    light /= np.linalg.norm(light)
    pix = np.arange(N).reshape(H, W)
    for obj in objects:
        # only the pixels inside the object's projected 3-D box can hit it
        bb = project_box(obj, pose, H, W, K, margin=3, min_size=1, clip_hi=(W, H))
        if bb is None:
            sel = np.arange(N)
        else:
            x0, y0, bh, bw = bb
            sel = pix[y0:y0 + bh + 1, x0:x0 + bw + 1].reshape(-1)
        t, n_o, o_o, d_o = _hit_object(obj, eye, d_w[sel])
        closer = t < best_t[sel]
        if not closer.any():
            continue
        n_w = n_o @ obj.Two[:3, :3].astype(np.float64).T
        shade = 0.35 + 0.65 * np.clip(n_w @ light, 0, 1)
        p_o = o_o + d_o * np.where(np.isfinite(t), t, 0)[:, None]
        stripes = 0.85 + 0.15 * np.sign(np.sin(18.0 * p_o[:, 0]) * np.sin(18.0 * p_o[:, 1] + 1.0) * np.sin(18.0 * p_o[:, 2] + 2.0))
        col = obj.color[None, :] * (shade * stripes)[:, None]
        hit = sel[closer]
        rgb[hit] = col[closer]
        inst[hit] = obj.instance_id
        best_t[hit] = t[closer]
    # floor z = 0 (checker) and sky, both instance 0
    with np.errstate(divide="ignore", invalid="ignore"):
        tf = -eye[2] / d_w[:, 2]
    floor = (tf > 0) & (tf < best_t) & (inst == 0)
    pf = eye + d_w * np.where(floor, tf, 0)[:, None]
    chk = ((np.floor(pf[:, 0] * 2) + np.floor(pf[:, 1] * 2)) % 2)
    fcol = np.where(chk[:, None] > 0, np.array([0.75, 0.72, 0.68]), np.array([0.45, 0.43, 0.4]))
    bg = inst == 0
    rgb[bg] = np.array([0.55, 0.7, 0.9])
    rgb[floor] = fcol[floor]
    depth = np.where(inst > 0, best_t / znorm, 0.0)  # z-depth in the camera frame (ray length / |dir|)
    rgb_u8 = np.clip(np.rint(rgb * 255.0), 0, 255).astype(np.uint8).reshape(H, W, 3)
    return rgb_u8, inst.reshape(H, W), depth.astype(np.float32).reshape(H, W)


def project_box(obj: SynthObject, pose: np.ndarray, H: int = H, W: int = W, K=(FX, FY, CX, CY), margin: int = 4,
                min_size: int = 8, clip_hi=None):
    """2-D box (x, y, h, w) of the object's 3-D box in the frame, clipped to the image; None if not visible."""
    fx, fy, cx, cy = K
    a = obj.half.astype(np.float64)
    corners = np.array([[sx * a[0], sy * a[1], sz * a[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    pw = corners @ obj.Two[:3, :3].astype(np.float64).T + obj.Two[:3, 3].astype(np.float64)
    Tcw = np.linalg.inv(pose.astype(np.float64))
    pc = pw @ Tcw[:3, :3].T + Tcw[:3, 3]
    if (pc[:, 2] <= 0.05).any():
        return None
    u = fx * pc[:, 0] / pc[:, 2] + cx
    v = fy * pc[:, 1] / pc[:, 2] + cy
    x0, x1 = int(math.floor(u.min())) - margin, int(math.ceil(u.max())) + margin
    y0, y1 = int(math.floor(v.min())) - margin, int(math.ceil(v.max())) + margin
    x0, y0 = max(x0, 0), max(y0, 0)
    hi_x, hi_y = clip_hi if clip_hi is not None else (W - 2, H - 2)
    x1, y1 = min(x1, hi_x), min(y1, hi_y)
    if x1 - x0 < min_size or y1 - y0 < min_size:
        return None
    return x0, y0, y1 - y0, x1 - x0  # x, y, h, w


def make_sequence(n_frames: int = 100, n_objects: int = 4, seed: int = 1337, H: int = H, W: int = W, K=(FX, FY, CX, CY)) -> SynthSequence:
    objects = make_objects(n_objects, seed)
    poses = make_poses(n_frames, seed)
    seq = SynthSequence(K=tuple(K), H=H, W=W, poses=poses, rgb=[], instance=[], depth=[], objects=objects)
    for fid, pose in enumerate(poses):
        rgb, inst, depth = render_frame(objects, pose, H, W, K)
        seq.rgb.append(rgb)
        seq.instance.append(inst)
        seq.depth.append(depth)
        for obj in objects:
            bb = project_box(obj, pose, H, W, K)
            if bb is not None and (inst == obj.instance_id).any():
                obj.boxes.append((fid, *bb))
    return seq


def _quat_from_R(R: np.ndarray):
    """(qx, qy, qz, qw) of a rotation matrix."""
    R = R.astype(np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        qw, qx, qy, qz = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        qw = (R[k, j] - R[j, k]) / s
        qx, qy, qz = q
    return qx, qy, qz, qw


def write_sequence(seq: SynthSequence, out_dir: str, depth_factor: float = 1.0 / 5000.0) -> None:
    """Writes the reference's on-disk schema: config.yaml, img.txt, groundtruth.txt, rgb/ depth/ instance/
    PNGs and obj_offline/k.txt (nerf_data.cu:31-112, nerf.cu:58-118).  Needs cv2 (present in this image)."""
    import cv2

    out = Path(out_dir)
    for sub in ("rgb", "depth", "instance", "obj_offline"):
        (out / sub).mkdir(parents=True, exist_ok=True)
    fx, fy, cx, cy = seq.K
    (out / "config.yaml").write_text(
        "%YAML:1.0\n"
        f"Camera.fx: {fx}\nCamera.fy: {fy}\nCamera.cx: {cx}\nCamera.cy: {cy}\n"
        f"Camera.H: {seq.H}\nCamera.W: {seq.W}\nDepthMapFactor: {depth_factor}\n")
    with open(out / "img.txt", "w") as f_img, open(out / "groundtruth.txt", "w") as f_gt:
        f_img.write("# timestamp filename\n")
        f_gt.write("# timestamp tx ty tz qx qy qz qw\n")
        for i, pose in enumerate(seq.poses):
            stamp = f"{i * 0.1:.6f}"
            name = f"{i:06d}.png"
            f_img.write(f"{stamp} {name}\n")
            qx, qy, qz, qw = _quat_from_R(pose[:3, :3])
            t = pose[:3, 3]
            f_gt.write(f"{stamp} {t[0]:.7f} {t[1]:.7f} {t[2]:.7f} {qx:.7f} {qy:.7f} {qz:.7f} {qw:.7f}\n")
            cv2.imwrite(str(out / "rgb" / name), seq.rgb[i][:, :, ::-1])
            cv2.imwrite(str(out / "instance" / name), seq.instance[i])
            d16 = np.clip(np.rint(seq.depth[i] / depth_factor), 0, 65535).astype(np.uint16)
            cv2.imwrite(str(out / "depth" / name), d16)
    for k, obj in enumerate(seq.objects):
        qx, qy, qz, qw = _quat_from_R(obj.Two[:3, :3])
        t = obj.Two[:3, 3]
        with open(out / "obj_offline" / f"{k}.txt", "w") as f:
            f.write("# class tx ty tz qx qy qz qw a1 a2 a3 ; then: timestamp x y h w\n")
            f.write(f"{obj.instance_id} {t[0]:.7f} {t[1]:.7f} {t[2]:.7f} {qx:.7f} {qy:.7f} {qz:.7f} {qw:.7f} "
                    f"{obj.half[0]:.7f} {obj.half[1]:.7f} {obj.half[2]:.7f}\n")
            for (fid, x, y, h, w) in obj.boxes:
                f.write(f"{fid * 0.1:.6f} {x} {y} {h} {w}\n")
