"""Known-answer tests that pin the CPU oracle from first principles (CPU only).

The reference ships no golden vectors for this path, so each test restates the
published algorithm independently in numpy / Python integers (pcg32, std::seed_seq, the coherent
prime hash, closed-form volume rendering, Adam, finite differences of the rendering loss) and
compares it with oracle/mon_oracle.cpp.  tests/test_golden_tcnn.py additionally pins rows A4, A5,
A8-A12 against outputs of the reference's own vendored tiny-cuda-nn run on a B200.
"""
import math

import numpy as np
import pytest

from conftest import uniform_open_closed

M64 = (1 << 64) - 1


# ------------------------------------------------------------------ A12: pcg32 / seed_seq / init
def pcg32_py(initstate, initseq=1):
    state, inc = 0, ((initseq << 1) | 1) & M64
    mult = 0x5851F42D4C957F2D

    def nxt():
        nonlocal state
        old = state
        state = (old * mult + inc) & M64
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF

    nxt()
    state = (state + initstate) & M64
    nxt()
    return nxt


def seed_seq_py(seeds, n):
    """std::seed_seq::generate, C++ standard [rand.util.seedseq] p8."""
    out = [0x8B8B8B8B] * n
    s = len(seeds)
    t = 11 if n >= 623 else 7 if n >= 68 else 5 if n >= 39 else 3 if n >= 7 else (n - 1) // 2
    p, q = (n - t) // 2, (n - t) // 2 + t
    m = max(s + 1, n)
    T = lambda x: (x ^ (x >> 27)) & 0xFFFFFFFF
    for k in range(m):
        r1 = (1664525 * T(out[k % n] ^ out[(k + p) % n] ^ out[(k - 1) % n])) & 0xFFFFFFFF
        r2 = (r1 + (s if k == 0 else (k % n + seeds[k - 1]) if k <= s else k % n)) & 0xFFFFFFFF
        out[(k + p) % n] = (out[(k + p) % n] + r1) & 0xFFFFFFFF
        out[(k + q) % n] = (out[(k + q) % n] + r2) & 0xFFFFFFFF
        out[k % n] = r2
    for k in range(m, m + n):
        r3 = (1566083941 * T((out[k % n] + out[(k + p) % n] + out[(k - 1) % n]) & 0xFFFFFFFF)) & 0xFFFFFFFF
        r4 = (r3 - k % n) & 0xFFFFFFFF
        out[(k + p) % n] ^= r3
        out[(k + q) % n] ^= r4
        out[k % n] = r4
    return out


def test_seed_seq(oracle):
    assert list(oracle.seed_seq_1(1337)) == seed_seq_py([1337], 2)
    assert list(oracle.seed_seq_1(7)) == seed_seq_py([7], 2)


def test_pcg32_stream_and_advance(oracle):
    nxt = pcg32_py(0xDEADBEEF)
    want = np.array([np.uint32((nxt() >> 9) | 0x3F800000).view(np.float32) - np.float32(1.0) for _ in range(64)], np.float32)
    got = oracle.pcg32_floats(0xDEADBEEF, 0, 64)
    assert np.array_equal(got, want)
    assert np.array_equal(oracle.pcg32_floats(0xDEADBEEF, 37, 10), want[37:47])  # advance(k) == k draws


def test_param_init(oracle):
    cfg = oracle.default_config()
    p = oracle.init_params(cfg, 1337)
    n_mlp = oracle.n_mlp_params(cfg)
    assert n_mlp == 64 * 32 + 16 * 64 and p.size == 1911808
    # MLP: xavier uniform from sequential draws of pcg32(seed_seq{1337}[0])
    nxt = pcg32_py(seed_seq_py([1337], 2)[0])
    f = lambda: np.float32(np.uint32((nxt() >> 9) | 0x3F800000).view(np.float32) - np.float32(1.0))
    a = np.float32(math.sqrt(np.float32(6.0) / np.float32(96.0)))
    want = np.array([f() * np.float32(2.0) * a - a for _ in range(8)], np.float32)
    assert np.array_equal(p[:8], want)
    assert np.abs(p[:2048]).max() <= math.sqrt(6 / 96) and np.abs(p[2048:3072]).max() <= math.sqrt(6 / 80)
    g = p[n_mlp:]
    assert np.abs(g).max() <= 1e-4 * (1 + 1e-6) and abs(g.mean()) < 1e-6 and g.std() == pytest.approx(2e-4 / math.sqrt(12), rel=0.01)
    # grid element j*T + t is draw (4t + j) after the 3072 MLP draws (random.h:66-92)
    T = ((g.size + 3) // 4 + 127) // 128 * 128
    draws = oracle.pcg32_floats(int(seed_seq_py([1337], 2)[0]), 3072, 8)
    for t_, j in [(0, 0), (0, 1), (1, 0), (1, 3)]:
        want = np.float32(draws[4 * t_ + j]) * np.float32(2e-4) + np.float32(-1e-4)
        assert abs(g[j * T + t_] - want) <= 1e-11


# ------------------------------------------------------------------ A4 geometry and indices
def test_grid_layout(oracle):
    cfg = oracle.default_config()
    off, sc, res, n = oracle.grid_layout(cfg)
    assert list(off[:4]) == [0, 4096, 36864, 102400] and off[16] == 954368 and n == 1908736
    assert list(res[:3]) == [16, 32, 64] and res[15] == 524288
    assert np.array_equal(sc, np.array([16 * 2 ** l - 1 for l in range(16)], np.float32))


def test_hash_indices(oracle):
    cfg = oracle.default_config()
    pts = np.array([[0.0, 0.0, 0.0], [0.5, 0.25, 0.75], [0.999, 0.001, 0.5]], np.float32)
    idx, w = oracle.encode_corners(cfg, pts)
    off, sc, res, _ = oracle.grid_layout(cfg)
    for pi, p in enumerate(pts):
        for l in (0, 1, 2, 7, 15):
            size = int(off[l + 1] - off[l])
            q = [np.float32(np.float32(p[d]) * sc[l]) + np.float32(0.5) for d in range(3)]  # no FMA difference on these values
            g = [int(math.floor(x)) for x in q]
            fr = [np.float32(x) - np.float32(gi) for x, gi in zip(q, g)]
            for c in range(8):
                cc = [g[d] + ((c >> d) & 1) for d in range(3)]
                if l < 2:
                    want = (cc[0] + cc[1] * int(res[l]) + cc[2] * int(res[l]) ** 2) % size
                else:
                    want = ((cc[0] * 1) ^ ((cc[1] * 2654435761) & 0xFFFFFFFF) ^ ((cc[2] * 805459861) & 0xFFFFFFFF)) % size
                assert idx[pi, l, c] == want, (pi, l, c)
                ww = np.float32(1.0)
                for d in range(3):
                    ww = np.float32(ww * (fr[d] if (c >> d) & 1 else np.float32(1.0) - fr[d]))
                assert w[pi, l, c] == pytest.approx(float(ww), abs=1e-6)
    assert np.allclose(w.sum(axis=2), 1.0, atol=1e-5)


def test_encode_constant_table(oracle):
    """A table holding one constant must interpolate to that constant (weights sum to 1, fp16 rounding only)."""
    cfg = oracle.default_config()
    n_grid = oracle.n_params(cfg) - oracle.n_mlp_params(cfg)
    grid = np.full(n_grid, oracle.f2h(np.array([0.5], np.float32))[0], np.uint16)
    pts = np.random.default_rng(0).random((64, 3), dtype=np.float32)
    enc = oracle.h2f(oracle.encode(cfg, grid, pts))
    assert np.abs(enc - 0.5).max() <= 2e-3


# ------------------------------------------------------------------ A1
def test_ray_box(oracle):
    bmin, bmax = [-1, -1, -1], [1, 1, 1]
    hit, t0, t1 = oracle.ray_intersect(bmin, bmax, [0, 0, -3], [0, 0, 1])
    assert hit and t0 == 2.0 and t1 == 4.0
    hit, t0, t1 = oracle.ray_intersect(bmin, bmax, [0, 0, 0], [1, 0, 0])       # origin inside: tmin < 0
    assert hit and t0 == -1.0 and t1 == 1.0
    assert not oracle.ray_intersect(bmin, bmax, [0, 3, -3], [0, 0, 1])[0]      # passes above
    hit, t0, t1 = oracle.ray_intersect(bmin, bmax, [3, 3, 3], [1, 1, 1])      # box behind the origin: the slab test reports it, with t < 0
    assert hit and t1 == -2.0
    hit, t0, t1 = oracle.ray_intersect(bmin, bmax, [-3, 0.5, 0.5], [1, 0, 0])  # zero direction components -> inf slabs
    assert hit and t0 == 2.0 and t1 == 4.0


# ------------------------------------------------------------------ A6 closed form
def test_volume_render_constant_density(oracle):
    R, S = 4, 32
    sigma, logit = 3.0, 0.7
    out = np.zeros((R * S, 16), np.float32)
    out[:, :3] = logit
    out[:, 3] = math.log(sigma)
    out_h = oracle.f2h(out)
    sig_h = math.exp(float(oracle.h2f(out_h[0, 3:4])[0]))
    col_h = 1 / (1 + math.exp(-float(oracle.h2f(out_h[0, 0:1])[0])))
    t = np.tile(np.linspace(0.05, 1.0, S, dtype=np.float32), R)
    bg = np.full((R, 3), 0.25, np.float32)
    rgb, dep, mask = oracle.volume_render(R, S, out_h, t, bg)
    T = math.exp(-sig_h * 1.0)  # first interval is measured from the origin: transmittance depends on t_last only
    assert mask == pytest.approx(1 - T, abs=1e-5)
    assert rgb == pytest.approx(col_h * (1 - T) + T * 0.25, abs=1e-5)


def test_volume_render_early_stop(oracle):
    R, S = 1, 32
    out = np.zeros((S, 16), np.float32)
    out[:, 3] = 8.0                      # huge density: T < 1e-4 after the first sample
    out[0, 0], out[1:, 0] = -5.0, 5.0    # only sample 0 may contribute colour
    t = np.linspace(0.1, 1.0, S, dtype=np.float32)
    rgb, dep, mask = oracle.volume_render(R, S, oracle.f2h(out), t, np.zeros((1, 3), np.float32))
    assert mask[0] == pytest.approx(1.0, abs=1e-6) and dep[0] == pytest.approx(0.1, abs=1e-4)
    assert rgb[0, 0] == pytest.approx(1 / (1 + math.exp(5.0)), abs=1e-4)


# ------------------------------------------------------------------ A7 by finite differences
def _loss_f64(logits, t, tgt, tgt_d, bg, is_obj):
    """Scalar whose gradient the reference hand-derives (nerf_model.cu:854-945), float64, no early stop."""
    rgb = 1 / (1 + np.exp(-logits[:, :3]))
    sig = np.exp(logits[:, 3])
    dt = np.diff(np.concatenate([[0.0], t]))
    alpha = 1 - np.exp(-sig * dt)
    Tn = np.concatenate([[1.0], np.cumprod(1 - alpha)])
    w = alpha * Tn[:-1]
    C = (w[:, None] * rgb).sum(0) + Tn[-1] * bg
    D = (w * t).sum()
    mask = 1 - Tn[-1]
    if is_obj:
        L = ((C - tgt) ** 2).sum() - 0.5 * mask
        if tgt_d > 0:
            L += 0.5 * abs(D - tgt_d)
        return L
    # background rays: no colour gradient through sigma, +0.5*mask, +0.01*sum(sigma); colour channels keep theirs
    return None


def test_loss_backward_object_ray_finite_differences(oracle):
    rng = np.random.default_rng(3)
    S, R = 32, 1
    logits = rng.normal(0, 1, (S, 4))
    logits[:, 3] = rng.normal(0.5, 0.8, S)
    lh = oracle.f2h(np.pad(logits, ((0, 0), (0, 12))).astype(np.float32))
    lf = oracle.h2f(lh)[:, :4].astype(np.float64)
    t = np.sort(rng.uniform(0.5, 1.5, S)).astype(np.float32)
    tgt, bg = rng.random(3).astype(np.float32), rng.random((1, 3)).astype(np.float32)
    tgt_d = np.float32(1.0)
    rgb, dep, mask = oracle.volume_render(R, S, lh, t, bg)
    assert mask[0] < 1.0 - 1e-3  # keeps every sample visited
    scale = 128.0
    dout, loss = oracle.loss_backward(R, S, scale, lh, t, np.array([1], np.uint8), tgt[None], np.array([tgt_d]), rgb, dep, mask)
    got = oracle.h2f(dout)[:, :4].astype(np.float64) / (scale / R)
    num = np.zeros((S, 4))
    eps = 1e-5
    for n in range(S):
        for c in range(4):
            a, b = lf.copy(), lf.copy()
            a[n, c] += eps
            b[n, c] -= eps
            num[n, c] = (_loss_f64(a, t.astype(np.float64), tgt, tgt_d, bg[0], True) - _loss_f64(b, t.astype(np.float64), tgt, tgt_d, bg[0], True)) / (2 * eps)
    assert np.allclose(got, num, rtol=5e-3, atol=2e-5)
    assert np.all(oracle.h2f(dout)[:, 4:] == 0)
    want_log = ((rgb[0] - tgt) ** 2).mean() + 0.5 * np.sign(dep[0] - tgt_d) * (dep[0] - tgt_d) + (1 - mask[0])
    assert loss[0] == pytest.approx(float(want_log), rel=1e-5)


def test_loss_backward_background_ray(oracle):
    rng = np.random.default_rng(4)
    S, R = 32, 1
    logits = np.pad(rng.normal(0, 0.5, (S, 4)), ((0, 0), (0, 12))).astype(np.float32)
    lh = oracle.f2h(logits)
    lf = oracle.h2f(lh).astype(np.float64)
    t = np.sort(rng.uniform(0.5, 1.5, S)).astype(np.float32)
    tgt, bg = rng.random(3).astype(np.float32), rng.random((1, 3)).astype(np.float32)
    rgb, dep, mask = oracle.volume_render(R, S, lh, t, bg)
    dout, loss = oracle.loss_backward(R, S, 128.0, lh, t, np.array([0], np.uint8), tgt[None], np.array([0.0]), rgb, dep, mask)
    got = oracle.h2f(dout).astype(np.float64) / 128.0
    sig = np.exp(lf[:, 3])
    dt = np.diff(np.concatenate([[0.0], t.astype(np.float64)]))
    want_sigma = sig * dt * 0.5 * (1 - mask[0]) + 0.01 * sig   # nerf_model.cu:938-940
    assert np.allclose(got[:, 3], want_sigma, rtol=3e-3, atol=1e-6)
    assert loss[0] == pytest.approx(float(((rgb[0] - tgt) ** 2).mean() + mask[0]), rel=1e-5)


# ------------------------------------------------------------------ A5 / A8 against float64 matrix algebra
def test_mlp_forward_backward_vs_float64(oracle):
    rng = np.random.default_rng(5)
    for n_hidden in (1, 2):
        cfg = oracle.default_config(n_hidden_layers=n_hidden)
        n_mlp = oracle.n_mlp_params(cfg)
        N = 96
        wh = oracle.f2h(rng.normal(0, 0.2, n_mlp).astype(np.float32))
        eh = oracle.f2h(rng.normal(0, 0.5, (N, 32)).astype(np.float32))
        hid, out = oracle.mlp_forward(cfg, wh, eh)
        w = oracle.h2f(wh).astype(np.float64)
        x = oracle.h2f(eh).astype(np.float64)
        mats = [w[:2048].reshape(64, 32)] + [w[2048 + i * 4096: 2048 + (i + 1) * 4096].reshape(64, 64) for i in range(n_hidden - 1)]
        Wout = w[n_mlp - 1024:].reshape(16, 64)
        a, acts = x, []
        for M in mats:
            a = np.maximum(a @ M.T, 0)
            a = oracle.h2f(oracle.f2h(a.astype(np.float32))).astype(np.float64)  # activations are stored fp16
            acts.append(a)
        want = a @ Wout.T
        assert np.allclose(oracle.h2f(out), want, rtol=2e-3, atol=2e-3)
        assert np.allclose(oracle.h2f(hid[-1]), acts[-1], rtol=2e-3, atol=2e-3)
        # backward with the oracle's own hidden activations
        dh = oracle.f2h(np.pad(rng.normal(0, 0.05, (N, 4)), ((0, 0), (0, 12))).astype(np.float32))
        d_enc, dW = oracle.mlp_backward(cfg, wh, eh, hid, dh, round_fp16=False)
        g = oracle.h2f(dh).astype(np.float64)
        H = [oracle.h2f(hid[i]).astype(np.float64) for i in range(n_hidden)]
        dW_out = g.T @ H[-1]
        gh = (g @ Wout) * (H[-1] > 0)
        assert np.allclose(dW[n_mlp - 1024:].reshape(16, 64), dW_out, rtol=1e-4, atol=1e-5)
        gh = oracle.h2f(oracle.f2h(gh.astype(np.float32))).astype(np.float64)
        for li in range(n_hidden - 1, 0, -1):
            dWl = gh.T @ H[li - 1]
            o = 2048 + (li - 1) * 4096
            assert np.allclose(dW[o:o + 4096].reshape(64, 64), dWl, rtol=1e-4, atol=1e-5)
            gh = (gh @ mats[li]) * (H[li - 1] > 0)
            gh = oracle.h2f(oracle.f2h(gh.astype(np.float32))).astype(np.float64)
        assert np.allclose(dW[:2048].reshape(64, 32), gh.T @ x, rtol=1e-4, atol=1e-5)
        assert np.allclose(oracle.h2f(d_enc), gh @ mats[0], rtol=2e-3, atol=1e-4)


# ------------------------------------------------------------------ A9 against a numpy scatter
def test_encode_backward_vs_numpy_scatter(oracle):
    cfg = oracle.default_config()
    rng = np.random.default_rng(6)
    N = 200
    pts = rng.random((N, 3), dtype=np.float32)
    d_enc = oracle.f2h(rng.normal(0, 1e-2, (N, 32)).astype(np.float32))
    grad = oracle.encode_backward(cfg, pts, d_enc, mode=1)
    idx, w = oracle.encode_corners(cfg, pts)
    off, _, _, n_grid = oracle.grid_layout(cfg)
    want = np.zeros(n_grid, np.float64)
    g = oracle.h2f(d_enc).astype(np.float64)
    for l in range(16):
        for f in range(2):
            contrib = oracle.h2f(oracle.f2h((g[:, 2 * l + f, None] * w[:, l, :]).astype(np.float32))).astype(np.float64)
            np.add.at(want, (off[l] + idx[:, l, :].astype(np.int64)) * 2 + f, contrib)
    assert np.allclose(grad, want, rtol=1e-5, atol=1e-9)
    # fp16 sequential accumulation (mode 0) stays within fp16 rounding of the fp32 sum
    grad16 = oracle.encode_backward(cfg, pts, d_enc, mode=0)
    assert np.allclose(grad16, grad, rtol=5e-3, atol=1e-6)


# ------------------------------------------------------------------ A10 / A11 toy
def test_adam_and_ema_first_steps(oracle):
    cfg = oracle.default_config()
    P, n_mlp = oracle.n_params(cfg), oracle.n_mlp_params(cfg)
    pf = np.zeros(P, np.float32)
    pf[:4] = [0.5, -0.25, 0.125, 1.0]
    pf[n_mlp:n_mlp + 3] = [1e-4, -1e-4, 5e-5]
    ph = oracle.f2h(pf)
    m, v, ps, ema = np.zeros(P, np.float32), np.zeros(P, np.float32), np.zeros(P, np.uint32), np.zeros(P, np.uint16)
    g = np.zeros(P, np.float32)
    g[:4] = [128.0, -64.0, 0.0, 1.0]              # loss-scaled gradients (scale 128)
    g[n_mlp:n_mlp + 3] = [12.8, 0.0, -0.128]
    before = pf.copy()
    oracle.optimizer_step(cfg, 1, g, pf, ph, m, v, ps, ema)
    # first Adam step moves every touched weight by lr * sign(g) (bias-corrected m / sqrt(v) = +-1)
    assert pf[0] == pytest.approx(before[0] - 1e-2, abs=1e-6)
    assert pf[1] == pytest.approx(before[1] + 1e-2, abs=1e-6)
    # MLP weight with zero data gradient still gets the L2 term 1e-6 * w -> also a full-size first step
    assert pf[2] == pytest.approx(before[2] - 1e-2, abs=1e-6) and ps[2] == 1
    # grid entries: zero gradient => untouched, no step increment (adam.h:75-79)
    assert pf[n_mlp + 1] == before[n_mlp + 1] and ps[n_mlp + 1] == 0 and m[n_mlp + 1] == 0
    assert pf[n_mlp] == pytest.approx(before[n_mlp] - 1e-2, abs=1e-6) and ps[n_mlp] == 1
    assert pf[n_mlp + 2] == pytest.approx(before[n_mlp + 2] + 1e-2, abs=1e-6)
    assert m[0] == pytest.approx(0.1 * (1.0 + 1e-6 * 0.5), rel=1e-6) and v[0] == pytest.approx(0.01 * 1.0, rel=1e-5)
    # EMA step 1: (0 * ... + w * 0.05) / (1 - 0.95) = w, stored fp16
    assert np.array_equal(ema[:4], oracle.f2h(oracle.h2f(ph[:4]) * np.float32(0.05) / np.float32(1 - 0.95)))
    # second step with the same gradient: m/sqrt(v) stays 1 -> another full step; EMA = (0.95*0.05*e1 + 0.05*w2)/(1-0.95^2)
    w1 = oracle.h2f(ph[:1])[0]
    oracle.optimizer_step(cfg, 2, g, pf, ph, m, v, ps, ema)
    assert pf[0] == pytest.approx(before[0] - 2e-2, abs=2e-6) and ps[0] == 2
    w2 = oracle.h2f(ph[:1])[0]
    want = (w1 * 0.95 * (1 - 0.95) + w2 * 0.05) / (1 - 0.95 ** 2)
    assert oracle.h2f(ema[:1])[0] == pytest.approx(want, rel=2e-3)


def test_lr_decay_schedule(oracle):
    cfg = oracle.default_config(decay_start=2, decay_interval=2, decay_base=0.5)
    P = oracle.n_params(cfg)
    steps = []
    for step in (1, 2, 3, 4, 5, 6):
        pf = np.zeros(P, np.float32)
        ph, m, v = oracle.f2h(pf), np.zeros(P, np.float32), np.zeros(P, np.float32)
        ps, ema = np.zeros(P, np.uint32), np.zeros(P, np.uint16)
        g = np.zeros(P, np.float32)
        g[0] = 128.0
        oracle.optimizer_step(cfg, step, g, pf, ph, m, v, ps, ema)
        steps.append(-pf[0] / 1e-2)
    # decay fires when the nested step (before increment) >= decay_start: factor 0.5^(floor((s-1-2)/2)+1)
    assert steps == pytest.approx([1, 1, 0.5, 0.5, 0.25, 0.25], rel=1e-5)


# ------------------------------------------------------------------ A2/A3 on the small sequence
def test_generate_rays_and_samples(oracle, small_seq):
    seq, obj = small_seq, small_seq.objects[0]
    frames = oracle.Frames(seq.rgb, seq.instance, seq.depth, seq.poses)
    R, S = 512, 32
    rng = np.random.default_rng(1)
    sxy, col = uniform_open_closed(rng, (R, 2)), uniform_open_closed(rng, (R, 3))
    bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
    n_in, rays, rinst, tgt, tgtd = oracle.generate_rays(R, obj.boxes, frames, seq.H, seq.W, seq.K, obj.Tow, bmin, bmax,
                                                        obj.instance_id, True, sxy, col)
    assert 0 < n_in <= R
    d = rays[:, 3:6]
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    assert np.all(rays[:, 8] > rays[:, 7]) and np.all(rays[:, 7] >= 0)
    assert np.array_equal(rays[n_in:], rays[np.arange(n_in, R) % n_in])            # roll-over padding
    assert np.array_equal(tgt[n_in:], tgt[np.arange(n_in, R) % n_in])
    bgr = rinst[:n_in] == 0
    assert np.array_equal(tgt[:n_in][bgr], col[:n_in][bgr])                          # background target = RandColors[slot]
    assert np.all(tgtd[:n_in][bgr] == 0) and np.all(tgtd[:n_in][~bgr] > 0)
    # entry/exit points lie on the box surface
    for k in (0, n_in // 2, n_in - 1):
        o, dd, t0, t1 = rays[k, :3], rays[k, 3:6], rays[k, 7], rays[k, 8]
        for tt in (t0, t1):
            p = o + tt * dd
            assert np.all(p >= bmin - 1e-4) and np.all(p <= bmax + 1e-4)
            assert np.min(np.minimum(np.abs(p - bmin), np.abs(p - bmax))) < 1e-4
    dt = uniform_open_closed(rng, (R, S))
    pts, t = oracle.sample_points(rays, S, bmin, bmax, dt)
    t = t.reshape(R, S)
    assert np.all(np.diff(t, axis=1) > -1e-7) and np.all(t[:, 0] >= rays[:, 7]) and np.all(t[:, -1] <= rays[:, 8] + 1e-5)
    assert pts.min() > -1e-3 and pts.max() < 1 + 1e-3
