#!/usr/bin/env python
"""bench.py — train iterations/s per object of the Multi-Object-NeRF hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R] [--hidden-layers H]

A "step" is ONE training iteration of one object (GenerateBatch -> Step_No_Compacted -> optimizer_step,
MON/Core/src/nerf_model.cu:1630-1648) on one batch of R rays x 32 samples.  Workload = BASELINE.json
configs[1]: OfflineNeRF, 1 object per GPU, base.json (16-level hash grid, 64-wide MLP), synthetic 'room'
stand-in at 800x800.  With N GPUs object k trains on GPU k (objects are independent: no data-path collective,
weak scaling); `value` is the aggregate iterations/s over all objects.

value : K iterations timed on the device (CUDA events on the object's stream), keyframes already in HBM.
e2e   : the same through the C ABI from HOST buffers: keyframe upload (H2D) + box upload + K iterations +
        loss read-back (D2H), wall clock.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

S = 32
FRAMES = 30        # 30 keyframes x (1.92 MB rgb + 0.64 MB mask + 2.56 MB depth) = 154 MB > 126 MB L2
BYTES_ENC_PER_POINT = 512       # 16 levels x 8 corners x 2 features x 2 B
FLOPS_MLP_TRAIN_PER_POINT = {1: 18432, 2: 43008}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel, from the committed ncu --set full captures
# (profiles/r1t_ncu_kernels.txt: encode, scatter, fused MLP; profiles/r1d_ncu_mlp_optimizer.txt: optimizer) at R=4096,
# one hidden layer.  ncu flushes the caches before every replayed launch, so these are COLD-cache figures: in the
# running job the 46 MB of per-object state stays L2-resident between kernels.
TRAFFIC_NCU = {"encode": 5.42e6, "scatter": 12.77e6, "mlp_fused": 8.70e6, "optimizer": 50.3e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    # 250: past the start-up phase of a fresh object (the first ~100 iterations scatter a dense gradient and run up to
    # 2x slower, profiles/r1g_pdl_ab.txt); the offline job is 5000 iterations, so the steady state is what it pays for
    ap.add_argument("--warmup", type=int, default=250)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=4096, help="rays per batch (reference: 4096, nerf_model.h:173)")
    ap.add_argument("--hidden-layers", type=int, default=1, help="MLP hidden layers (reference base.json: 1)")
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--objects", type=int, default=0, help="objects in the job (default: one per GPU); object k trains on GPU k mod N")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "src": "fallback"}


def make_scene(n_objects: int, n_frames: int):
    from ro_map_b200 import synthetic as syn
    return syn.make_sequence(n_frames=n_frames, n_objects=n_objects, seed=1337)


def cpu_baseline(seq, obj, R, hidden, seconds):
    """The CPU oracle (a restatement of the reference's arithmetic, kind "port") on all host cores."""
    from oracle import mon_oracle as orc
    cores = os.cpu_count() or 1
    cfg = orc.default_config(n_hidden_layers=hidden)
    o = orc.OracleObject(cfg, R, S, obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id, True, n_threads=cores)
    frames = orc.Frames(seq.rgb, seq.instance, seq.depth, seq.poses)
    o.train_iter_rng(obj.boxes, frames, seq.H, seq.W, seq.K, 0)  # warm-up
    t0, n = time.perf_counter(), 0
    while True:
        o.train_iter_rng(obj.boxes, frames, seq.H, seq.W, seq.K, n + 1)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or n >= 200:
            break
    return {"value": n / dt, "unit": "iters/s", "cores": cores, "kind": "port",
            "sample": f"{n} full training iterations (R={R} rays x {S} samples, fp32 + software fp16 rounding) in {dt:.1f} s on {cores} threads"}


# ----------------------------------------------------------------------------- arms
def reference_core_gpu(args, seq, obj):
    """The reference Core on this box's GPU: unmodified vendored tiny-cuda-nn (hash grid, FullyFusedMLP, CUTLASS
    wgrad, Adam/EMA) driven through Train_Step's per-iteration call sequence incl. its 3 stream syncs and 3 cuRAND
    host calls; RO-MAP's glue kernels are the reference's own — oracle/ref/Makefile compiles nerf_model.cu unmodified,
    from where it lies, against stand-in headers for Eigen/OpenCV/GLEW (absent from this image); a library built without
    that (ref_is_genuine() == 0) runs them restated in the reference's launch shape and says so.  None when unavailable."""
    sys.path.insert(0, str(ROOT / "oracle" / "ref"))
    try:
        import ref_binding
        if not ref_binding.LIB_PATH.exists():
            return None
        import torch
        if not torch.cuda.is_available():
            return None
        m = ref_binding.RefModel(args.hidden_layers, 1337)
    except Exception as e:  # library missing / no device: fall back to the CPU port below
        print(f"[bench] reference Core on GPU unavailable: {e}", file=sys.stderr)
        return None
    m.scene(seq.rgb, seq.instance, seq.depth, seq.poses, seq.H, seq.W, seq.K, obj.boxes, obj.Tow, -1.1 * obj.half, 1.1 * obj.half,
            obj.instance_id, True, args.rays)
    m.train(max(args.warmup, 3))
    with ClockSampler(0) as clocks:
        dev_ms, wall_ms, loss, _ = m.train(args.steps)
    genuine = m.is_genuine()
    m.close()
    return {"device_ms": dev_ms, "wall_ms": wall_ms, "loss": loss, "clocks": clocks.summary(), "genuine": genuine}


def run_reference(args, rank, world):
    """Reference arm.  Rank 0 alone runs it; other ranks exit.
    The reference's implementation of this path is CUDA (tiny-cuda-nn): when oracle/_ref/libmon_ref.so travelled with
    the tree and a GPU is visible, the line's value is the reference Core measured on this box (kind "reference");
    the CPU restatement (oracle/, kind "port") is timed beside it on the host cores and becomes the line's value
    only when the reference library is unavailable."""
    if rank != 0:
        return
    seq = make_scene(1, args.frames)
    obj = seq.objects[0]
    ref = reference_core_gpu(args, seq, obj)
    base = cpu_baseline(seq, obj, args.rays, args.hidden_layers, max(2.0, min(args.cpu_seconds, 60.0)))
    if ref is not None:
        v = args.steps / (ref["wall_ms"] * 1e-3)   # the reference loop blocks on the host every iteration: wall clock IS its throughput
        line = {
            "impl": "reference", "metric": "train iters/sec per object", "value": v, "unit": "iters/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref["wall_ms"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 storage / f16 accumulate (tiny-cuda-nn)", "data": "synthetic",
            "config": workload_config(args, 1),
            "reference_kind": "reference Core on GPU: unmodified vendored tiny-cuda-nn (sm_100 build) + "
                              + ("RO-MAP's own nerf_model.cu kernels (compiled in place, unmodified)" if ref["genuine"] else "RO-MAP glue kernels restated in reference shape")
                              + ", Train_Step's call sequence; 1 object on 1 GPU",
            "device_ms_per_step": ref["device_ms"] / args.steps, "final_loss": ref["loss"], "clocks": ref["clocks"],
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64,
                    "note": "keyframes resident; per-iteration host syncs and cuRAND host calls included, as in Train_Step"},
        }
    else:
        v = base["value"]
        line = {
            "impl": "reference", "metric": "train iters/sec per object", "value": v, "unit": "iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 storage / f32 accumulate (CPU restatement)", "data": "synthetic",
            "config": workload_config(args, 1),
            "reference_kind": "CPU port (oracle/): the reference Core is CUDA-only and its library did not travel",
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    print(json.dumps(line))


def workload_config(args, n_objects):
    return {"workload": "OfflineNeRF 1 object per GPU, base.json (16-lvl hash 2^16x2 fp16, MLP 32-64%s-16pad), synthetic 'room' 800x800" % ("-64" if args.hidden_layers == 2 else ""),
            "rays_per_batch": args.rays, "samples_per_ray": S, "points_per_iter": args.rays * S, "n_hidden_layers": args.hidden_layers,
            "objects": n_objects, "keyframes": args.frames, "partition": "object k -> GPU k mod N, per-object streams (no collective)",
            "l2_policy": "keyframe set 154 MB > 126 MB L2; per-object state (42 MB) is L2-resident in steady state by design, as in production back-to-back iterations"}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from ro_map_b200 import build, core, partition

    distributed = world > 1
    if distributed:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build.build()
    if distributed:
        dist.barrier()
    if core.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    gpu = local_rank

    n_objects = args.objects if args.objects > 0 else world      # default: one object per GPU (weak scaling)
    seq = make_scene(n_objects, args.frames)
    mine = partition.assign_objects(n_objects, world)[rank]        # object k -> rank k % world, no collective on the data path
    R, K, Wm = args.rays, args.steps, args.warmup
    cfg = core.default_config(rays_per_batch=R, n_hidden_layers=args.hidden_layers)

    # keyframes in PINNED host memory (the C ABI then DMAs straight out of them, asynchronously)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    pinned = [(pin(seq.rgb[i]), pin(seq.instance[i]), pin(seq.depth[i])) for i in range(len(seq.poses))]

    def upload(ds):
        for i in range(len(seq.poses)):
            (_, rgb), (_, inst), (_, dep) = pinned[i]
            ds.add_frame(i, rgb, inst, dep, seq.poses[i])

    def make_objects(ds):
        objs = []
        for k in mine:
            o = seq.objects[k]
            n = core.NerfObject(ds, cfg, o.Tow, -1.1 * o.half, 1.1 * o.half, o.instance_id)
            n.set_bboxes(o.boxes)
            objs.append(n)
        return objs

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-timed: inputs resident.  Objects of one rank train concurrently on per-object streams.
    ds = core.Dataset(gpu, *seq.K, seq.H, seq.W, len(seq.poses), True)
    upload(ds)
    nerfs = make_objects(ds)
    for n in nerfs:
        n.train(max(Wm, 3))
    with ClockSampler(gpu) as clocks:
        barrier()
        l0 = sum(n.launch_count for n in nerfs)
        for n in nerfs:
            n.train_async(K)
        for n in nerfs:
            n.sync()
        barrier()
        ms = max([n.last_train_ms for n in nerfs], default=0.0)   # CUDA events on each object's stream
        launches = sum(n.launch_count for n in nerfs) - l0
        # keep the GPU under the same load while nvidia-smi samples (each sample is 100 ms; the timed region may be shorter)
        t_end = time.perf_counter() + 1.0
        while nerfs and time.perf_counter() < t_end:
            for n in nerfs:
                n.train_async(K)
            for n in nerfs:
                n.sync()
    ms_max = partition.reduce_max(ms, "cuda")
    launches = int(partition.reduce_sum(float(launches), "cuda"))
    loss = nerfs[0].train(1) if nerfs else float("nan")

    # ---------------- per-stage device times for the roofline (live, CUDA events between kernels), rank 0 / first object
    stages = nerfs[0].train_profiled(50) if nerfs else {}

    # ---------------- rendered rays/s (BASELINE.json's second metric): NeRF_Model::Render of one full 800x800 view per
    # object with the inference (EMA) weights, 64 samples per ray, through the C ABI into host buffers (D2H included)
    render = None
    if nerfs:
        full = (0, 0, 0, seq.H, seq.W)
        nerfs[0].render(full, seq.poses[0])                          # warm-up (workspace allocation)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_views = 3
        masks = []
        for v in range(n_views):
            for n in nerfs:
                masks.append(n.render(full, seq.poses[v % len(seq.poses)])[2])
        render_s = time.perf_counter() - t0
        covered = [float(m.mean()) for m in masks]
        # a second figure on the object's own 2-D box (what RenderTestImg renders): most of its rays hit the 3-D box
        ob = seq.objects[mine[0]].boxes[0]
        nerfs[0].render(ob, seq.poses[ob[0]])
        t0 = time.perf_counter()
        for _ in range(n_views):
            nerfs[0].render(ob, seq.poses[ob[0]])
        box_s = (time.perf_counter() - t0) / n_views
        render = {"rays_per_s": len(nerfs) * n_views * seq.H * seq.W / render_s, "samples_per_ray": int(cfg.render_samples_per_ray),
                  "view": f"{seq.W}x{seq.H}", "ms_per_view": 1e3 * render_s / (n_views * len(nerfs)), "opaque_fraction": float(np.mean(covered)),
                  "object_box": {"h_w": [int(ob[3]), int(ob[4])], "ms_per_view": 1e3 * box_s, "rays_per_s": ob[3] * ob[4] / box_s},
                  "region": "mon_object_render: rays (misses finished and dropped in the ray kernel) + 64 samples/ray + encode + MLP + compositing "
                            "+ D2H of the view, wall clock"}
    for n in nerfs:
        n.close()

    # ---------------- end to end from host buffers through the C ABI
    ds2 = core.Dataset(gpu, *seq.K, seq.H, seq.W, len(seq.poses), True)
    # warm the graphs with a throw-away frame set so that capture cost is not billed to the timed region
    upload(ds2)
    nerfs2 = make_objects(ds2)
    for n in nerfs2:
        n.train(max(Wm, 3))
    barrier()
    t0 = time.perf_counter()
    upload(ds2)                                  # H2D: every keyframe again, from pinned host arrays (async DMA)
    for k, n in zip(mine, nerfs2):
        n.set_bboxes(seq.objects[k].boxes)       # H2D: 20 B per box
    for n in nerfs2:
        n.train_async(K)
    losses_e2e = [n.train(0) for n in nerfs2]    # waits for the K iterations + 48-byte D2H (loss, step) per object
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if distributed:
        dist.barrier()
    e2e_s = partition.reduce_max(e2e_s, "cuda")
    px = seq.H * seq.W
    h2d = len(seq.poses) * (px * 3 + px + px * 4 + 96) + sum(len(seq.objects[k].boxes) for k in mine) * 20
    h2d = partition.reduce_sum(float(h2d), "cuda")
    for n in nerfs2:
        n.close()

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    N = R * S
    stage_roof = {}
    enc_bytes = BYTES_ENC_PER_POINT * N
    flops = FLOPS_MLP_TRAIN_PER_POINT[args.hidden_layers] * N
    P = 1911808 if args.hidden_layers == 1 else 1911808 + 4096
    for name, ms_k in stages.items():
        s_k = ms_k * 1e-3
        if name in ("encode", "scatter"):
            ach = enc_bytes / s_k / 1e9
            stage_roof[name] = {"ms": ms_k, "bound": "hbm", "achieved": ach, "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
        elif name == "mlp_fused":
            ach = flops / s_k / 1e12
            stage_roof[name] = {"ms": ms_k, "bound": "tensor", "achieved": ach, "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"]}
        elif name == "optimizer":
            lo = 10 * P / s_k / 1e9
            stage_roof[name] = {"ms": ms_k, "bound": "hbm", "achieved": lo, "unit": "GB/s (10 B/param floor; 44 B touched)", "frac": lo / peaks["hbm_gbs"]}
        else:
            stage_roof[name] = {"ms": ms_k}
    dominant = max(("encode", "scatter", "mlp_fused", "optimizer"), key=lambda k: stages[k])
    d = stage_roof[dominant]
    roofline = {"kernel": dominant, "bound": d["bound"], "achieved": d["achieved"], "unit": d["unit"].split(" ")[0],
                "peak": peaks["hbm_gbs"] if d["bound"] == "hbm" else peaks["tflops_sustained"], "peak_source": peaks["src"] + (" (sustained)" if d["bound"] == "tensor" else ""),
                "frac": d["frac"], "traffic": TRAFFIC_NCU.get(dominant), "traffic_source": "profiles/ (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                "algorithmic_bytes_or_flops_per_launch": enc_bytes if d["bound"] == "hbm" and dominant != "optimizer" else (10 * P if dominant == "optimizer" else flops),
                "stage_ms_sum": sum(stages.values()), "stages": stage_roof,
                "note": "stage times come from a serial (un-forked) replay with a CUDA event between kernels; the production graph overlaps batch+points of iteration i+1 with scatter+optimizer of iteration i"}

    # the CPU baseline is timed on rank 0 of the single-GPU run only (it is a property of the host, not of N)
    base = cpu_baseline(seq, seq.objects[0], R, args.hidden_layers, args.cpu_seconds) if world == 1 else None

    iters_per_s = n_objects * K / (ms_max * 1e-3)
    line = {
        "metric": "train iters/sec per object", "value": iters_per_s, "unit": "iters/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 storage / f32 accumulate", "data": "synthetic", "config": workload_config(args, n_objects),
        "iters_per_s_per_object": iters_per_s / max(n_objects, 1),
        "rays_per_s": iters_per_s * R, "points_per_s": iters_per_s * N, "final_loss": loss,
        "render": render,
        "clocks": clocks.summary(),
        "e2e": {"value": n_objects * K / e2e_s, "unit": "iters/s", "h2d_bytes_per_step": h2d / (n_objects * K), "d2h_bytes_per_step": 48.0 / K,
                "seconds": e2e_s, "final_loss": losses_e2e[0] if losses_e2e else None,
                "region": "keyframe upload from pinned host memory + box upload + K iterations per object + loss read-back"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": base,
    }
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
