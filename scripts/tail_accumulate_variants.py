"""Timing + self-consistency of the two populate variants that were written without GPU access
(non-affine tail, accumulate_weights) on the C2 flow.  Prints ONE JSON line.  ``bench.py`` runs
this in a SUBPROCESS (its own CUDA context) after its own measurements, so that a fault in these
not-yet-hardware-verified paths cannot disturb the headline numbers; it can also be run alone:
    python scripts/tail_accumulate_variants.py [rows_per_turn]
No oracle here (scripts are product-side): the checks are internal -- the general engine with
identity maps must reproduce the fused affine path, a sigmoid map must keep every row inside its
bounds, and the accumulated pool must come out of the rows that were drawn.
"""

import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def timed(fn, reps=5, warm=2):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def main():
    import torch

    from nessai_b200 import _lib
    from nessai_b200.flowmodel import B200FlowModel
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import GeneralPopulateEngine, PopulateEngine

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    g = np.load(os.path.join(REPO, "tests", "golden", "c2_realnvp_mlp.npz"))
    cfg = json.loads(str(g["flow_config"]))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    D = cfg["n_inputs"]
    fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
    fm.initialise()
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    fm.model.eval()
    names = [f"x{i}" for i in range(D)]
    dtype = get_dtype(names)
    scale, shift = np.full(D, 1.5), np.linspace(-0.3, 0.3, D)
    lo, hi = np.full(D, -10.0), np.full(D, 10.0)
    lpc, radius = -D * np.log(20.0), 4.9
    out = {"rows_per_turn": n}

    aff = PopulateEngine(fm, names, dtype)
    aff.seed = 1234
    aff.configure(scale, shift, lo, hi, lpc, radius)
    aff._ensure(n, n, False)
    out["affine_draw_ms"] = timed(lambda: aff.draw_turn(n))
    out["affine_accept_ms"] = timed(lambda: aff.accept_turn(n, 0))

    # (1) identity maps through draw + tail + float64-row accept == the fused affine path
    gen = GeneralPopulateEngine(fm, names, dtype)
    gen.seed = 1234
    gen.configure(np.zeros(D, dtype=np.int32), scale, shift, lo, hi, lpc, radius)
    gen._ensure(n, n, False)
    gen.draw_turn(n)
    aff.draw_turn(n)
    la, lb = gen.d_logw[:n], aff.d_logw[:n]
    same_nan = bool(torch.equal(torch.isnan(la), torch.isnan(lb)))
    ok = ~torch.isnan(la) & ~torch.isnan(lb)
    out["identity_tail_vs_fused"] = {
        "same_dropped_rows": same_nan,
        "max_abs_dlogw": float((la[ok] - lb[ok]).abs().max()),
        "max_abs_dx": float((gen.physical_x(n) - aff.physical_x(n)).abs().max()),
        "stats_equal": bool(torch.equal(gen.d_stats, aff.d_stats)),
    }
    ca, cb = gen.accept_turn(n, 0).clone(), aff.accept_turn(n, 0).clone()
    out["identity_tail_vs_fused"]["same_accept_counts"] = bool(torch.equal(ca, cb))
    nb = int(ca[1]) * dtype.itemsize
    out["identity_tail_vs_fused"]["records_max_abs_diff"] = float(
        (gen.d_rows[:nb].view(torch.int32).view(-1, dtype.itemsize // 4)[:, : 2 * D].contiguous().view(torch.float64)
         - aff.d_rows[:nb].view(torch.int32).view(-1, dtype.itemsize // 4)[:, : 2 * D].contiguous().view(torch.float64)
         ).abs().max()) if nb else 0.0

    # (2) logit on every other parameter, rescale-to-bounds on the rest
    kind = (np.arange(D) % 2).astype(np.int32)
    s2 = np.where(kind == 1, 20.0, 1.5)
    t2 = np.where(kind == 1, -10.0, shift)
    gen.configure(kind, s2, t2, lo, hi, lpc, radius)
    out["general_draw_plus_tail_ms"] = timed(lambda: gen.draw_turn(n))
    out["general_accept_x64_ms"] = timed(lambda: gen.accept_turn(n, 0))
    tail_ms = timed(lambda: gen._after_draw(n))  # the tail kernel alone (re-applied to the same rows)
    out["tail_kernel_ms"] = tail_ms
    out["tail_kernel_gbs"] = n * (4 * D + 8 + 8 * D + 16) / (tail_ms * 1e-3) / 1e9
    gen.draw_turn(n)
    x = gen.physical_x(n)
    okr = ~torch.isnan(gen.d_logw[:n])
    out["general_valid_fraction"] = float(okr.float().mean())
    out["general_rows_in_bounds"] = bool(((x[okr] >= -10) & (x[okr] <= 10)).all())
    rows, p, a = gen.run(50_000, n, max_samples=50 * n)
    out["general_populate"] = {"rows": int(len(rows)), "n_proposed": int(p), "n_accepted": int(a)}

    # (2b) every kind once, against the same maps written with torch float64 ops on the device
    # (an independent evaluation: torch.sigmoid / atan2 / hypot / special.ndtr / ndtri ...)
    kind = np.array([0, 1, 2, 3, 4, 5, 6, 7, 10, 9, 8, 0, 1, 2, 3, 0][:D], dtype=np.int32)
    sc = np.array([1.5, 20.0, -4.0, 2.0, 1.2, 20.0, 0.9, 0.5, 1.0, 1.0, 1.0, 1.0, 20.0, 3.0, 1.0, 1.0][:D])
    sh = np.array([0.2, -10.0, 9.0, -1.0, 0.1, -10.0, 0.4, 0.0, 0.0, 0.0, 0.0, 0.0, -10.0, -9.0, 0.0, 0.0][:D])
    pa = np.array([1.0, 1.0, 1.0, 1.0, 0.3, 1.0, 0.2, 1.0, 1.0, 1.0, 1.0, 1.0, 1.7, 1.0, 0.6, 1.0][:D])
    pb = np.array([0.0, 0.0, 0.0, 0.0, 1.5, 0.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, -0.3, 0.0, 0.2, 0.0][:D])
    src = np.stack([np.arange(D)] * 2, axis=1).astype(np.int32)
    if D >= 11:
        src[7] = src[8] = (8, 7)
        src[9] = src[10] = (9, 10)
        gen.configure(kind, sc, sh, np.full(D, -np.inf), np.full(D, np.inf), lpc, radius, pre_scale=pa, pre_shift=pb,
                      src=src)
        gen.draw_turn(n)
        xp = gen.d_xp[:n].to(torch.float64)
        t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).cuda()  # noqa: E731
        u = xp[:, torch.from_numpy(src[:, 0].astype(np.int64)).cuda()] * t(pa) + t(pb)
        u1 = xp[:, torch.from_numpy(src[:, 1].astype(np.int64)).cuda()]
        cols, logj = [], torch.zeros(n, dtype=torch.float64, device="cuda")
        for d in range(D):
            k, v = int(kind[d]), u[:, d]
            if k == 1:
                h = torch.sigmoid(v)
                logj += torch.log(h) + torch.log1p(-h)
            elif k == 2:
                h = v.abs()
            elif k == 3:
                h, logj = torch.exp(v), logj + v
            elif k == 4:
                h = torch.log(v)
                logj -= h
            elif k == 5:
                h, logj = torch.special.ndtr(v), logj - 0.9189385332046727 - 0.5 * v * v
            elif k == 6:
                h = torch.special.ndtri(v)
                logj += 0.9189385332046727 + 0.5 * h * h
            elif k in (7, 8):
                h = torch.atan2(u1[:, d], xp[:, int(src[d, 0])])
                h = torch.where((h < 0) & (k == 8), h + 2 * np.pi, h)
            elif k in (9, 10):
                h = torch.hypot(xp[:, int(src[d, 0])], u1[:, d])
                logj -= torch.log(h)
            else:
                h = v
            if k < 7:
                logj += float(np.log(abs(sc[d])) + np.log(abs(pa[d])))
            cols.append(h * sc[d] + sh[d])
        x_t = torch.stack(cols, dim=1)
        x_k = gen.physical_x(n)
        fin = torch.isfinite(x_t) & torch.isfinite(x_k)
        rel = ((x_k - x_t).abs() / (1.0 + x_t.abs()))[fin]
        okk = ~torch.isnan(gen.d_logw[:n])
        # log q of the flow alone is not kept after the tail: compare the weights' dependence on
        # log|J| through differences between the kernel's log w and lpc + chi prior - (flow log q - logj)
        out["tail_vs_torch_float64"] = {
            "kinds": kind.tolist(), "max_rel_dx": float(rel.max()), "finite_fraction": float(fin.float().mean()),
            "same_nonfinite_pattern": bool(torch.equal(torch.isfinite(x_t), torch.isfinite(x_k))),
            "valid_fraction": float(okk.float().mean()),
        }

    # (3) accumulate_weights
    import time

    aff.run_accumulate(20_000, n, max_samples=40 * n)  # warm-up (allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows, p, a = aff.run_accumulate(20_000, n, max_samples=40 * n)
    dt = time.perf_counter() - t0
    info = aff.last_accumulate
    lw = aff.d_logw[: len(info["draw_offsets"]) * info["stride"]]
    okw = ~torch.isnan(lw)
    out["accumulate"] = {
        "rows": int(len(rows)), "n_proposed": int(p), "n_accepted": int(a), "turns": len(info["draw_offsets"]),
        "rejection_steps": len(info["rejects"]), "wall_ms": 1e3 * dt, "proposed_rows_per_s": p / dt,
        "n_expected_last": info["n_expected"][-1],
        "n_expected_check": float(torch.exp(lw[okw] - lw[okw].max()).sum()),
    }
    # accumulate over the non-affine tail with identity maps == the affine accumulating loop
    gen.configure(np.zeros(D, dtype=np.int32), scale, shift, lo, hi, lpc, radius)
    aff.seed = gen.seed = 777
    aff._turn_rows = gen._turn_rows = 0
    ra, pa, aa = aff.run_accumulate(20_000, n, max_samples=4 * n)
    rg, pg, ag = gen.run_accumulate(20_000, n, max_samples=4 * n)
    same = len(ra) == len(rg) and aa == ag
    out["accumulate_identity_tail_vs_affine"] = {
        "same_turns": pa == pg, "n_accepted": [int(aa), int(ag)],
        "n_expected_rel_diff": float(abs(aff.last_accumulate["n_expected"][-1] - gen.last_accumulate["n_expected"][-1])
                                     / aff.last_accumulate["n_expected"][-1]),
        "records_max_abs_diff": float(max(np.abs(ra[nm] - rg[nm]).max() for nm in names)) if same and len(ra) else None,
    }

    # the sum-exp reduction alone, over 8e6 rows of weights (HBM-bound: 8 B/row)
    m = min(8 * n, aff.d_logw.shape[0])
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    se_ms = timed(lambda: aff._call(lib.nb200_sum_exp, [aff.d_logw.data_ptr(), m, aff.d_stats.data_ptr(),
                                                         aff.d_partials.data_ptr(), aff._N_PARTIALS, st], "nb200_sum_exp"))
    out["sum_exp"] = {"rows": m, "kernel_ms": se_ms, "gbs": 8 * m / (se_ms * 1e-3) / 1e9}
    out["kernel_launches"] = _lib.launch_count()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
