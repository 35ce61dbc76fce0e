"""GPU parity, live: the sm_100a path (through the C ABI) against the REFERENCE ITSELF on the same GPU.

oracle/_ref/libmon_ref.so holds the reference's own code compiled unmodified from where it lies (vendored tiny-cuda-nn +
RO-MAP's nerf_model.cu, see oracle/ref/Makefile); it travels to the GPU box with the tree.  Both sides get the same
keyframes, boxes, parameters and injected random numbers, run ONE training iteration (Train_Step's body) and render the
same window.  The reference compacts rays with an atomicAdd (slot order = a hardware race) while ours is ascending, and
the stratification jitter / background colour are indexed by SLOT — so the injected jitter and colour rows are made
identical for every slot and the pixels are chosen so that every ray survives (no roll-over padding), which makes every
result independent of the order; rays are matched by sorting.  Every measured difference is written to
gpurun_out/live_parity_vs_reference_nh*.json before the tolerances are applied.

Tolerances are fp16-sized: the reference accumulates the MLP in fp16 on mma.sync in a hardware-defined order, ours in fp32
in TMEM, and the grid gradients are fp16 atomics in nondeterministic order on both sides.
Skipped (not failed) only when the reference library did not travel; never used by the product.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle" / "ref"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_binding():
    import ref_binding as rb
    if not rb.LIB_PATH.exists():
        pytest.skip("oracle/_ref/libmon_ref.so is not in the tree (built by oracle/ref/Makefile where /root/reference exists)")
    return rb


@pytest.fixture(scope="module")
def core():
    from ro_map_b200 import build, core
    build.build()
    if core.device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return core


def _sorted_by_ray(rays):
    return np.lexsort((rays[:, 8], rays[:, 5], rays[:, 4], rays[:, 3]))


@pytest.mark.parametrize("n_hidden", [1, 2])
def test_one_iteration_and_render_against_the_reference(core, oracle, ref_binding, small_seq, n_hidden):
    run_live_parity(core, oracle, ref_binding, small_seq, small_seq.objects[0], 256, n_hidden, f"live_parity_vs_reference_nh{n_hidden}")


def run_live_parity(core, oracle, ref_binding, seq, obj, R, n_hidden, report_name, all_survive=True, warm_iters=300):
    """One Train_Step iteration + one Render, reference library vs the CUDA path, same inputs (see the module docstring).
    all_survive=False: every second slot dies (corner of the 2-D box outside the 3-D box) and the batch is padded by roll-over
    (fill_rollover_rays) on both sides; with n_in = R / 2 every surviving ray is used exactly twice whatever the compaction
    order, so loss and gradients remain comparable although the reference's slot order is an atomicAdd race."""
    import test_golden_romap as tg
    S, S2 = 32, 64
    bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
    ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
    for i in range(len(seq.poses)):
        ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
    g = core.NerfObject(ds, core.default_config(rays_per_batch=R, n_hidden_layers=n_hidden), obj.Tow, bmin, bmax, obj.instance_id, 1337)
    g.set_bboxes(obj.boxes)
    r = ref_binding.RefModel(n_hidden, 1337)
    assert r.is_genuine(), "the reference library must contain RO-MAP's own kernels"
    r.scene(seq.rgb, seq.instance, seq.depth, seq.poses, seq.H, seq.W, seq.K, obj.boxes, obj.Tow, bmin, bmax, obj.instance_id, True, R)
    report = {"n_hidden_layers": n_hidden, "rays_per_batch": R, "image": [seq.H, seq.W], "keyframes": len(seq.poses), "all_slots_survive": all_survive}

    # A12: parameter initialisation — the reference's Trainer(seed 1337) against ours, all 1.9 M values
    assert np.array_equal(r.get(0), g.state("master"))

    # a trained starting point, the same on both sides, with FRESH optimizer state on both (a scratch object trains; set_params
    # leaves Adam moments, step counters and the EMA untouched in both implementations)
    g0 = core.NerfObject(ds, g.cfg, obj.Tow, bmin, bmax, obj.instance_id, 1337)
    g0.set_bboxes(obj.boxes)
    g0.train(warm_iters)
    master = g0.state("master")
    g0.close()
    g.set_params(master)
    r.set_params(master)

    rng = np.random.default_rng(77)
    u = lambda shape: (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)   # noqa: E731  (0,1] like cuRAND
    col = np.repeat(u((1, 3)), R, axis=0)          # one background colour and one jitter row for every slot: results do not depend on the slot order
    dt = np.repeat(u((1, S)), R, axis=0)
    # every slot must survive (pixel not occluded, ray hits the object box): with n_in < R the batch is padded by repeating slots
    # 0 .. R-n_in-1, and WHICH rays those are is the reference's atomicAdd race — loss and gradients would then differ legitimately
    frames = oracle.Frames(seq.rgb, seq.instance, seq.depth, seq.poses)
    # all_survive: every slot's pixel is re-drawn until its ray survives.  Otherwise: exactly every second slot survives and the
    # others die, so that n_in = R / 2 and the roll-over padding (slot i >= n_in repeats slot i % n_in) uses every surviving ray
    # exactly twice WHATEVER the compaction order was: sums over the batch stay independent of the reference's race
    want = np.ones(R, bool) if all_survive else (np.arange(R) % 2 == 0)
    sxy = np.zeros((R, 2), np.float32)
    cand = u((R, 2))
    todo = np.arange(R)
    for _ in range(400):
        alive = np.zeros(len(todo), bool)
        for k, i in enumerate(todo):
            alive[k] = oracle.generate_rays(1, [obj.boxes[i % len(obj.boxes)]], frames, seq.H, seq.W, seq.K, obj.Tow, bmin, bmax, obj.instance_id, True,
                                            cand[i:i + 1], col[:1])[0] == 1
        ok = alive == want[todo]
        sxy[todo[ok]] = cand[todo[ok]]
        todo = todo[~ok]
        if len(todo) == 0:
            break
        cand[todo] = u((len(todo), 2))
        if not all_survive:
            # dead slots are rare under uniform sampling (the 2-D box hugs the 3-D box): aim at the corners of the rectangle
            dead = ~want[todo]
            cand[todo[dead]] = np.where(u((int(dead.sum()), 2)) < 0.5, 0.002, 0.998).astype(np.float32) * u((int(dead.sum()), 2)) ** 0.05
    else:
        pytest.fail(f"no pixel with the wanted fate found for {len(todo)} slots")
    _, _, loss_r, n_in_r = r.train(1, (sxy, col, dt))
    loss_g, n_in_g = g.train_injected(sxy, col, dt)
    assert n_in_g == n_in_r
    assert n_in_g == (R if all_survive else R // 2)
    n = n_in_g

    rays_g, rays_r = g.last("rays").reshape(R, 9), r.last(0, R * 9).reshape(R, 9)
    pg, pr = _sorted_by_ray(rays_g[:n]), _sorted_by_ray(rays_r[:n])
    assert tg.rays_close(rays_g[:n][pg], rays_r[:n][pr])                               # <= 8 ulp / 1.5e-6 (see test_golden_romap.rays_close)
    assert np.array_equal(g.last("ray_instance")[:n][pg], r.last(12, R)[:n][pr])
    assert np.array_equal(g.last("target").reshape(R, 3)[:n][pg], r.last(10, R * 3).reshape(R, 3)[:n][pr])   # pixels and the (constant) background colour

    if not all_survive:
        # fill_rollover_rays (nerf_model.cu:280-294): slot i >= n_in repeats slot i % n_in, on both sides
        for rays in (rays_g, rays_r):
            assert np.array_equal(rays[n:], rays[np.arange(n, R) % n])
    lim_l2, lim_cos, lim_frac = 0.15, 0.995, 0.97

    # ---- measure everything first (the report is written before any tolerance is applied), then assert ----------------------
    checks = []   # (name, measured, limit, ok)

    def check(name, measured, limit, ok=None):
        measured = float(measured)
        checks.append((name, measured, limit, bool(measured <= limit if ok is None else ok)))
        report[name] = measured

    # network output (fp16 logits): the reference accumulates in fp16, ours in fp32
    out_g = g.last("out").reshape(R, S, 4)[:n][pg]
    out_r = r.last(4, R * S * 16).reshape(R, S, 16)[:n][pr][:, :, :4]
    check("out_logits_mean_abs_diff", np.abs(out_g - out_r).mean(), 5e-2)                 # measured 1.4e-2 (trained logits reach |x| ~ 30: fp16 ulp 0.03)
    check("out_logits_p999_abs_diff", np.quantile(np.abs(out_g - out_r), 0.999), 1.5)   # measured 0.37 / 0.44
    rel = np.abs(out_g - out_r) / (1.0 + np.abs(out_r))
    report["out_logits_rel_quantiles_50_99_999"] = [float(np.quantile(rel, q)) for q in (0.5, 0.99, 0.999)]
    # compositing: per-ray colour / depth / opacity
    for name, which, w, lim_max, lim_mean in (("rgb_rays", 5, 3, 0.12, 3e-3), ("depth_rays", 6, 1, 0.6, 5e-3), ("mask_rays", 7, 1, 0.2, 2e-3)):   # measured max 0.038 / 0.19 / 0.06 (one ray), mean 6e-4 / 1e-3 / 3e-4
        a_, b_ = g.last(name).reshape(R, w)[:n][pg], r.last(which, R * w).reshape(R, w)[:n][pr]
        check(name + "_max_abs_diff", np.abs(a_ - b_).max(), lim_max)
        check(name + "_mean_abs_diff", np.abs(a_ - b_).mean(), lim_mean)
    lg, lr = g.last("loss")[:n][pg], r.last(13, R)[:n][pr]
    check("loss_rays_max_abs_diff", np.abs(lg - lr).max(), 0.25)                          # measured 0.073 (the same single ray)
    report["loss"] = [float(loss_g), float(loss_r)]
    check("loss_rel_diff", abs(loss_g - loss_r) / max(abs(loss_r), 1e-6), 8e-2)          # measured 7e-4 / 2.2e-2;        # SumLoss / R (R % 256 == 0)
    # dL/dout (fp16, loss scale 128)
    do_g = g.last("dout").reshape(R, S, 4)[:n][pg]
    do_r = r.last(8, R * S * 16).reshape(R, S, 16)[:n][pr][:, :, :4]
    check("dout_rel_l2", np.linalg.norm(do_g - do_r) / np.linalg.norm(do_r), 0.2)         # measured 0.039 / 0.062
    check("dout_support_disagreement", ((do_g != 0).any(-1) != (do_r != 0).any(-1)).mean(), 1e-2)   # measured 1.2e-3 / 2.7e-3

    # parameter gradients (loss-scaled fp16).  The reference's weight gradients are CUTLASS split-K GEMMs accumulating in fp16
    gg, gr = g.state("grad"), r.get(3)
    n_mlp = g.n_mlp
    blocks = [("W_in", 0, 64 * 32)] + [(f"W_h{i}", 64 * 32 + i * 4096, 64 * 32 + (i + 1) * 4096) for i in range(n_hidden - 1)] + [("W_out", n_mlp - 16 * 64, n_mlp)]
    for bname, lo, hi in blocks:
        a_, b_ = gg[lo:hi], gr[lo:hi]
        check(f"grad_{bname}_rel_l2", np.linalg.norm(a_ - b_) / max(np.linalg.norm(b_), 1e-30), lim_l2)
        check(f"grad_{bname}_max_abs_rel_to_scale", np.abs(a_ - b_).max() / max(np.abs(b_).max(), 1e-30), 0.3)
        cos = float(np.dot(a_, b_) / max(np.linalg.norm(a_) * np.linalg.norm(b_), 1e-30))
        check(f"grad_{bname}_cosine", cos, lim_cos, ok=cos >= lim_cos)                         # measured 0.9993 .. 0.99999 (rel. L2 0.4 - 3.8 %)
    sup = ((gg[n_mlp:] != 0) == (gr[n_mlp:] != 0)).mean()
    check("grad_grid_support_agreement", sup, 0.98, ok=sup >= 0.98)                      # measured 0.9939 / 0.9960 (fp16 underflow of ~0 sums)
    touched = (gr[n_mlp:] != 0) & (gg[n_mlp:] != 0)
    check("grad_grid_rel_l2", np.linalg.norm((gg[n_mlp:] - gr[n_mlp:])[touched]) / np.linalg.norm(gr[n_mlp:][touched]), lim_l2)   # measured 0.031 / 0.044
    gcos = float(np.dot(gg[n_mlp:], gr[n_mlp:]) / (np.linalg.norm(gg[n_mlp:]) * np.linalg.norm(gr[n_mlp:])))
    check("grad_grid_cosine", gcos, lim_cos, ok=gcos >= lim_cos)                              # measured 0.9990 / 0.9995

    # after the optimizer step: fp32 master weights.  The FIRST Adam step moves every touched parameter by lr * sign(gradient) = 1e-2
    # whatever the magnitude, so entries whose ~0 gradient differs in sign (or in being touched at all) end 1e-2 .. 2e-2 apart and
    # everything else agrees to rounding
    dm = np.abs(g.state("master") - r.get(0))
    frac = float((dm <= 1e-5).mean())
    check("master_fraction_within_1e-5", frac, lim_frac, ok=frac >= lim_frac)                    # measured 0.9928 / 0.9941
    check("master_max_abs_diff", dm.max(), 2.5e-2)
    frac_mlp = float((dm[:n_mlp] <= 1e-5).mean())
    check("master_mlp_fraction_within_1e-5", frac_mlp, 0.95, ok=frac_mlp >= 0.95)        # measured 0.9967 / 0.9876

    # Render (EMA weights after that one step) of a window across the object's edge, same injected jitter
    fid, x, y, h, w = [int(v) for v in obj.boxes[0]]
    box = (fid, max(0, x - 6), y + h // 3, 24, 32)
    jit = u((box[3] * box[4], S2))
    rr = r.render(box, seq.poses[fid], jit)
    rgb_g, dep_g, mask_g = g.render(box, seq.poses[fid], use_ema=True, rand_dt=jit)
    rgb_g, dep_g, mask_g = rgb_g.reshape(-1, 3), dep_g.reshape(-1), mask_g.reshape(-1)
    hit = rr["in_box"] == 1
    report["render_rays_hit_miss"] = [int(hit.sum()), int((~hit).sum())]
    miss_ok = bool((rgb_g[~hit] == 1.0).all() and not dep_g[~hit].any() and not mask_g[~hit].any())   # misses: white, depth 0, mask 0
    check("render_misses_white", 0.0 if miss_ok else 1.0, 0.0)
    same = mask_g == rr["mask"]
    check("render_mask_agreement", same.mean(), 0.98, ok=same.mean() >= 0.98)
    report["render_opaque_fraction"] = float(rr["mask"].mean())
    mse = float(((rgb_g - rr["rgb"])[same] ** 2).mean())
    psnr = float(-10 * np.log10(max(mse, 1e-12)))
    check("render_psnr_db", psnr, 40.0, ok=psnr >= 40.0)                                # between the two renders; measured 52.8 / 51.4 dB
    check("render_depth_mean_abs_diff", np.abs(dep_g - rr["depth"])[same].mean(), 5e-3)   # measured 3e-4

    # one view of RenderVideo (nerf_model.cu:1832-1991): the reference's GenerateRenderVideoRays with a turn-table pose from its own
    # GenerateToc against mon_object_render_object_centric, same EMA weights, same injected jitter; a strip across the image centre
    # (the turn-table camera looks at the object origin) wide enough to leave the object on both sides
    radius = float(6.0 * np.max(obj.half))
    for theta in (6.0, 132.0, 306.0):
        Toc = r.generate_toc(theta, 30.0, radius)
        vh, vw = 8, min(seq.W, 4 * (int(2.6 * seq.K[0] * float(np.max(obj.half)) / radius) // 4 + 8))
        vbox = (0, seq.W // 2 - vw // 2, seq.H // 2 - vh // 2, vh, vw)
        jit = u((vh * vw, S2))
        vr = r.render2(vbox, Toc, object_centric=True, rand_dt=jit, want_rays=True)
        v_rgb, v_dep, v_mask = g.render(vbox, Toc, use_ema=True, rand_dt=jit, object_centric=True)
        v_rgb, v_dep, v_mask = v_rgb.reshape(-1, 3), v_dep.reshape(-1), v_mask.reshape(-1)
        vhit = vr["in_box"] == 1
        tag = f"video_theta{int(theta)}"
        report[tag + "_rays_hit_miss"] = [int(vhit.sum()), int((~vhit).sum())]
        check(tag + "_misses_white", 0.0 if bool((v_rgb[~vhit] == 1.0).all() and not v_dep[~vhit].any() and not v_mask[~vhit].any()) else 1.0, 0.0)
        vsame = v_mask == vr["mask"]
        check(tag + "_mask_agreement", vsame.mean(), 0.98, ok=vsame.mean() >= 0.98)
        report[tag + "_opaque_fraction"] = float(vr["mask"].mean())
        vmse = float(((v_rgb - vr["rgb"])[vsame] ** 2).mean())
        vpsnr = float(-10 * np.log10(max(vmse, 1e-12)))
        check(tag + "_psnr_db", vpsnr, 40.0, ok=vpsnr >= 40.0)
        check(tag + "_depth_mean_abs_diff", np.abs(v_dep - vr["depth"])[vsame].mean(), 5e-3)
        assert 0 < vhit.sum() < vhit.size, (tag, int(vhit.sum()), vhit.size)

    report["failed"] = [c[0] for c in checks if not c[3]]
    out = ROOT / "gpurun_out"
    if out.is_dir():
        (out / f"{report_name}.json").write_text(json.dumps(report, indent=1))
    print(json.dumps(report))
    g.close()
    r.close()
    ds.close()
    assert 0 < hit.sum() < hit.size
    assert not report["failed"], [(c[0], c[1], c[2]) for c in checks if not c[3]]
