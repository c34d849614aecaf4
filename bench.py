#!/usr/bin/env python
"""bench.py -- proposed live-points/sec of FlowProposal.populate (BASELINE.json).

Workload (config C2): 16-D correlated Gaussian, RealNVP with 4 coupling layers and
a [64, 64] MLP conditioner (weights: the reference-trained golden fixture),
z-score reparameterisation, constant-volume latent radius (0.95), uniform box
prior [-10, 10]^16, ``poolsize = drawsize = 1e6`` per GPU.  One "step" is one
``populate(worst_point, n_samples=pool)`` call = as many 1e6-row turns as the
reference's own loop makes (2 with the default ``max_samples``).

  value : proposed rows / s, device pipeline only (draw + accept kernels; no D2H)
  e2e   : n_proposed / population_time through ``B200FlowProposal.populate`` with
          host structured arrays in and out (the D2H copy of the accepted records
          and the host bookkeeping are inside the timed region)

``--impl reference`` times the UNMODIFIED reference's FlowProposal.populate
(baseline/_ref on the restated glasflow.nflows shim) on the host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

D = 16
POOL = 1_000_000
FLOPS_PER_ROW = 47.7e3  # SURVEY.md 8(d), C2 RealNVP/MLP
BYTES_PER_ROW = 68.0  # sample_and_log_prob, in-kernel RNG, z not returned
SEED = 20251017
TRAFFIC_BYTES = 33.7e6  # dram read + write per launch of the dominant kernel, ncu --set full (profiles/r2_ncu_populate_tcgen05_final.txt)


def load_c2():
    g = np.load(os.path.join(REPO, "tests", "golden", "c2_realnvp_mlp.npz"))
    cfg = json.loads(str(g["flow_config"]))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    return g, cfg, sd


def live_points(n=2000):
    rng = np.random.default_rng(SEED)
    idx = np.arange(D)
    cov = 0.5 ** np.abs(idx[:, None] - idx[None, :])
    return rng.multivariate_normal(np.zeros(D), cov, size=n), cov


class GaussianModel:
    """16-D correlated Gaussian likelihood, uniform prior on [-10, 10]^16."""

    def __init__(self):
        self.names = [f"x{i}" for i in range(D)]
        self.bounds = {n: [-10.0, 10.0] for n in self.names}
        _, cov = live_points(2)
        self._icov = np.linalg.inv(cov)
        self._norm = -0.5 * (D * np.log(2 * np.pi) + np.linalg.slogdet(cov)[1])

    def _arr(self, x):
        return np.stack([x[n] for n in self.names], axis=-1)

    def log_prior(self, x):
        a = self._arr(x)
        lp = np.full(a.shape[0] if a.ndim > 1 else 1, -D * np.log(20.0))
        return np.where(np.all((a >= -10) & (a <= 10), axis=-1), lp, -np.inf)

    def log_likelihood(self, x):
        a = self._arr(x)
        return self._norm - 0.5 * np.einsum("...i,ij,...j->...", a, self._icov, a)

    def log_likelihood_torch(self, x):
        """The same likelihood on (n, D) float64 device rows: the pool's logL is
        evaluated on the accepted records while they are still in HBM -- outside population_time
        in both arms, as flowproposal.py:518-523 times it."""
        import torch

        if getattr(self, "_icov_t", None) is None or self._icov_t.device != x.device:
            self._icov_t = torch.from_numpy(self._icov).to(x.device)
        return self._norm - 0.5 * ((x @ self._icov_t) * x).sum(dim=1)


def rosenbrock_live_points(n=2000, d=32):
    """The curved live-point set the C3 fixture was trained on (tests/golden/make_golden.py:
    rosenbrock_chain with default_rng(SEED + 32)): the ridge x_{i+1} ~ x_i^2 around (1, ..., 1)."""
    rng = np.random.default_rng(SEED + 32)
    x = np.empty((n, d))
    x[:, 0] = rng.normal(1.0, 0.3, n)
    for i in range(1, d):
        x[:, i] = 0.25 * x[:, i - 1] ** 2 + 0.75 + rng.normal(0, 0.1, n)
    return np.clip(x, -4.9, 4.9)


class RosenbrockModel:
    """32-D Rosenbrock likelihood, uniform prior on [-5, 5]^32 (/root/reference/examples/rosenbrock.py:20-45)."""

    def __init__(self, d=32):
        self.names = [f"x{i}" for i in range(d)]
        self.bounds = {n: [-5.0, 5.0] for n in self.names}
        self.d = d

    def _arr(self, x):
        return np.stack([x[n] for n in self.names], axis=-1)

    def log_prior(self, x):
        a = self._arr(x)
        lp = np.full(a.shape[0] if a.ndim > 1 else 1, -self.d * np.log(10.0))
        return np.where(np.all((a >= -5) & (a <= 5), axis=-1), lp, -np.inf)

    def log_likelihood(self, x):
        a = self._arr(x)
        return -np.sum(100.0 * (a[..., 1:] - a[..., :-1] ** 2.0) ** 2.0 + (1.0 - a[..., :-1]) ** 2.0, axis=-1)


def problem(fixture):
    """(model, raw live points) of a fixture's configuration."""
    if fixture.startswith("c3"):
        return RosenbrockModel(32), rosenbrock_live_points()
    return GaussianModel(), live_points()[0]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [v.strip() for v in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": mx,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------ ours
def load_fixture(name):
    g = np.load(os.path.join(REPO, "tests", "golden", f"{name}.npz"))
    cfg = json.loads(str(g["flow_config"]))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    return cfg, sd


def build_proposal(fixture, model, live_s, local_rank, pool):
    """B200FlowProposal on the reference-trained golden weights of ``fixture``."""
    import torch

    from nessai_b200.proposal import B200FlowProposal

    # (a dict: that flow_config, freshly initialised -- the reference's init -- instead of fixture weights)
    cfg, sd = (dict(fixture), None) if isinstance(fixture, dict) else load_fixture(fixture)
    torch.manual_seed(SEED)
    prop = B200FlowProposal(
        model, rng=np.random.default_rng(SEED), flow_config=cfg,
        training_config=dict(device_tag=f"cuda:{local_rank}"),
        output=tempfile.mkdtemp(), poolsize=pool, drawsize=pool, device_prior="auto",
    )
    prop.initialise()
    prop.check_state(live_s)  # z-score statistics
    if sd is not None:
        prop.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    prop.flow.model.eval()
    return prop


def quartiles(v):
    q = np.percentile(np.asarray(v, dtype=np.float64), [25, 50, 75])
    return float(q[1]), float(q[2] - q[0])


def measure(prop, worst, pool, steps, warmup, repeats, world, dev, kernel_reps=20, e2e_repeats=None):
    """One configuration through the three lenses of the bench line.  Every timed region is exactly
    ``steps`` steps between two barrier + synchronize brackets; the region is repeated ``repeats``
    times (>= 1 s of timed work in total) and the MEDIAN region is reported, with the inter-quartile
    range beside it.  Returns a dict; collective-safe (every rank calls it with the same arguments)."""
    import torch
    import torch.distributed as dist

    from nessai_b200 import _lib

    eng = prop._get_engine()
    max_samples = pool

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        # populate()'s loop, device side only: the same turn loop as the e2e call
        # (PopulateEngine.run) with the accepted records left in HBM
        return eng.run(pool, pool, max_samples=max_samples, to_host=False)[1]

    eng._ensure(1, pool, False)
    for _ in range(warmup):
        device_step()
    # ---- device pipeline (value)
    region_ms, region_rows = [], []
    launches = 0
    for r in range(repeats):
        barrier()
        _lib.reset_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        n_prop = 0
        for _ in range(steps):
            n_prop += device_step()
        ev1.record()
        barrier()
        launches = _lib.launch_count()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        region_ms.append(float(ms.item()))
        region_rows.append(n_prop)
    rates = [n / (ms * 1e-3) for n, ms in zip(region_rows, region_ms)]
    # ---- dominant kernel alone: the fused draw kernel, CUDA events on its stream
    n_local = eng._shard(pool)[0]
    kt = []
    for _ in range(kernel_reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.draw_turn(pool)
        b.record()
        torch.cuda.synchronize()
        kt.append(a.elapsed_time(b))
    # ---- end to end through the plugin-facing call (host arrays in/out)
    for _ in range(min(warmup, 2)):
        prop.populate(worst, n_samples=pool, max_samples=max_samples)
    # (a populate's WALL time is dominated by what the sampler does between two populates -- the
    # host likelihood of the whole pool, outside population_time -- so its regions are fewer)
    e2e_rates, d2h, wall = [], 0, 0.0
    for r in range(e2e_repeats or repeats):
        barrier()
        prop.population_time *= 0
        n_prop = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            if world > 1:
                # align the ranks: between two populates every rank evaluates the pool's likelihood on
                # its host, and the first collective of a populate would otherwise charge the skew of
                # that phase to population_time (the reference's own timer, flowproposal.py:400,518)
                dist.barrier()
            prop.populate(worst, n_samples=pool, max_samples=max_samples)
            n_prop += prop.n_proposed
            d2h = prop.samples.nbytes
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
        pop_s = torch.tensor([prop.population_time.total_seconds()], device=dev)
        if world > 1:
            dist.all_reduce(pop_s, op=dist.ReduceOp.MAX)
        e2e_rates.append(n_prop / float(pop_s.item()))
    v_med, v_iqr = quartiles(rates)
    ms_med, _ = quartiles(region_ms)
    e_med, e_iqr = quartiles(e2e_rates)
    k_med, k_iqr = quartiles(kt)
    return dict(value=v_med, value_iqr=v_iqr, ms_per_step=ms_med / steps, e2e=e_med, e2e_iqr=e_iqr,
                kernel_ms=k_med, kernel_ms_iqr=k_iqr, n_local=n_local, launches=int(launches),
                turns_per_step=int(round(region_rows[-1] / steps / pool)), d2h_bytes=int(d2h), repeats=repeats,
                e2e_regions=len(e2e_rates), timed_s=dict(device=sum(region_ms) * 1e-3, e2e_wall=wall),
                regions_value=[float(f"{v:.4g}") for v in rates], regions_e2e=[float(f"{v:.4g}") for v in e2e_rates],
                population_acceptance=prop.population_acceptance)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from nessai_b200.livepoint import numpy_array_to_live_points

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    live, _ = live_points()
    model = GaussianModel()
    live_s = numpy_array_to_live_points(live, model.names)
    live_s["logL"] = model.log_likelihood(live_s)
    worst = live_s[np.argmin(live_s["logL"])]
    pool = args.pool * world  # weak scaling: 1e6 rows per GPU per turn
    prop = build_proposal("c2_realnvp_mlp", model, live_s, local_rank, pool)
    # >= 1000 timed steps (>= 1 s of timed work) per lens, in regions of exactly `steps` steps
    repeats = int(min(200, max(5, -(-args.min_steps // args.steps))))
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e2e_repeats = int(min(repeats, max(5, -(-(400 if world == 1 else 100) // args.steps))))
    m = measure(prop, worst, pool, args.steps, args.warmup, repeats, world, dev, e2e_repeats=e2e_repeats)
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        peaks, which = measured_peaks()
        k_rows_s = m["n_local"] / (m["kernel_ms"] * 1e-3)
        tf = k_rows_s * FLOPS_PER_ROW / 1e12
        # the dominant kernel is event-timed on its own (launch, synchronise): the burst figure is
        # the denominator; the sustained one is quoted beside it
        peak_tf = peaks["bf16_tflops"]
        out = {
            "metric": "proposed live-points/sec (FlowProposal.populate, 16-D, pool 1e6)",
            "value": m["value"],
            "unit": "rows/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": m["ms_per_step"],
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic (Philox latent draws; reference-trained golden weights)",
            "config": {
                "workload": "C2: 16-D RealNVP 4x[64,64] MLP, poolsize=drawsize=1e6 per GPU, "
                            "zscore, constant-volume radius 0.95, uniform box prior",
                "rows_per_turn_per_gpu": args.pool,
                "turns_per_step": m["turns_per_step"],
                "l2": "each turn writes 80 MB of fresh outputs; L2 (126 MB) is flushed between steps by the accept kernels and the next turn; inputs are generated in-kernel",
                "tc_kernel": bool(os.environ.get("NB200_DISABLE_TC", "0") != "1"),
                "timing": f"value / kernel_ms are MEDIANS over {m['repeats']} (e2e: {m['e2e_regions']}) timed regions of exactly "
                          f"{args.steps} steps each (barrier + synchronize on both sides of every region, CUDA events, "
                          "max over ranks); *_iqr = inter-quartile range over the regions",
                "e2e_note": "n_proposed / population_time of B200FlowProposal.populate (max over ranks); N>1: ranks "
                            "aligned by a barrier before each populate, pool assembled in node-local shared pinned "
                            "host memory, d2h_bytes_per_step = bytes of the whole pool (each rank copies its own share)",
            },
            "spread": {"value_iqr": m["value_iqr"], "e2e_iqr": m["e2e_iqr"], "kernel_ms_iqr": m["kernel_ms_iqr"],
                       "regions": m["repeats"], "e2e_regions": m["e2e_regions"], "timed_s": m["timed_s"],
                       "regions_value": m["regions_value"], "regions_e2e": m["regions_e2e"]},
            "clocks": clk,
            "e2e": {
                "value": m["e2e"],
                "unit": "rows/s",
                "h2d_bytes_per_step": 4 * D * 8,
                "d2h_bytes_per_step": m["d2h_bytes"],
                "population_acceptance": m["population_acceptance"],
            },
            "gpu_launches": m["launches"],
            "roofline": {
                "kernel": "populate_draw (fused draw + inverse flow + rescale + weights)",
                "bound": "tensor",
                "achieved": tf,
                "peak": peak_tf,
                "unit": "TFLOP/s",
                "frac": tf / peak_tf,
                "peak_source": f"{which} bf16 burst (kernel timed alone); sustained: "
                               f"{peaks.get('bf16_tflops_sustained')} -> frac {tf / peaks.get('bf16_tflops_sustained', peak_tf):.4f}",
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, profiles/):
                # below the 80 MB the kernel writes because the tail of the output is still in the
                # 126 MB L2 at kernel end
                "traffic": TRAFFIC_BYTES,
                "kernel_ms": m["kernel_ms"],
                "rows_per_launch": m["n_local"],
                "hbm": {
                    "achieved": k_rows_s * BYTES_PER_ROW / 1e9,
                    "peak": peaks["hbm_gbs"],
                    "unit": "GB/s",
                    "frac": k_rows_s * BYTES_PER_ROW / 1e9 / peaks["hbm_gbs"],
                    "written_bytes_per_row": 80,
                },
            },
        }
    if world > 1 and not args.no_extras:
        extras = multi_gpu_extras(prop, model, live_s, worst, args, pool, repeats, world, rank, local_rank, dev)
        if rank == 0:
            out.update(extras)
    if world == 1 and not args.no_extras:
        # nessai's DEFAULT conditioner (ResidualNet, 2 blocks of 64) through the same call
        prop_r = build_proposal("c2_realnvp_resnet", model, live_s, local_rank, pool)
        mr = measure(prop_r, worst, pool, args.steps, args.warmup, max(5, repeats // 4), world, dev, kernel_reps=10)
        kr = mr["n_local"] / (mr["kernel_ms"] * 1e-3)
        resnet = {
            "workload": "C2': same as the headline with nessai's default ResidualNet conditioner (2 blocks x 64), "
                        "through B200FlowProposal.populate",
            "value": mr["value"], "value_iqr": mr["value_iqr"], "ms_per_step": mr["ms_per_step"],
            "e2e": mr["e2e"], "e2e_iqr": mr["e2e_iqr"], "kernel_ms": mr["kernel_ms"],
            "roofline": {"bound": "tensor", "achieved": kr * 146e3 / 1e12, "peak": peaks["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": kr * 146e3 / 1e12 / peaks["bf16_tflops"], "flops_per_row": 146e3,
                         "peak_source": f"{which} bf16 burst"},
        }
        if not args.no_cpu_baseline:
            resnet["cpu_baseline"] = cpu_baseline(threads=1, pool=min(args.cpu_pool, 200_000), fixture="c2_realnvp_resnet")
        # the reference's DEFAULT flow for a 16-parameter model: ResidualNet conditioner of width
        # 2 * n_inputs = 32 (flowmodel/utils.py:39-42); freshly initialised weights (no fixture of that shape)
        del prop_r
        prop_d = build_proposal(dict(n_inputs=D, n_blocks=4, n_layers=2, ftype="realnvp"), model, live_s, local_rank, pool)
        md = measure(prop_d, worst, pool, args.steps, args.warmup, max(5, repeats // 4), world, dev, kernel_reps=10)
        default_flow = {
            "workload": "C2'': the headline call with the reference's default flow_config for 16 parameters "
                        "(ResidualNet, n_neurons = 2 * n_inputs = 32; freshly initialised), through B200FlowProposal.populate",
            "value": md["value"], "value_iqr": md["value_iqr"], "ms_per_step": md["ms_per_step"],
            "e2e": md["e2e"], "e2e_iqr": md["e2e_iqr"], "kernel_ms": md["kernel_ms"],
        }
        del prop_d
        out["coupling_forward"] = coupling_roofline(dev, peaks, which)
        out["variants"] = {"c2_resnet_default_conditioner": resnet,
                           "c2_default_width": default_flow,
                           "train": train_variant("ours"),
                           "c3": c3_block(args, local_rank, dev, repeats, peaks, which),
                           "c5": c5_variant("ours"),
                           "nonaffine_tail_and_accumulate": isolated_variants(args.pool)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(threads=1, pool=args.cpu_pool)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_extras(prop, model, live_s, worst, args, pool, repeats, world, rank, local_rank, dev):
    """N > 1 only (collective: every rank calls it).  Beside the weak-scaling headline:
    ``strong``          one pool of ``args.pool`` rows per turn partitioned over the N GPUs;
    ``nccl_allgather``  the headline workload with the accepted records all-gathered over NCCL / NVLink
                        (north_star's data plane) instead of the node-local shared host pool;
    ``parity_vs_n1``    the pool the N ranks return against the pool ONE rank returns for the same seed
                        (the claim that the pool does not depend on the number of GPUs), byte for byte."""
    import torch
    import torch.distributed as dist

    from nessai_b200.proposal import PopulateEngine

    few = max(5, min(repeats // 4, -(-100 // args.steps)))
    out = {}
    prop_s = build_proposal("c2_realnvp_mlp", model, live_s, local_rank, args.pool)
    ms = measure(prop_s, worst, args.pool, args.steps, args.warmup, few, world, dev, kernel_reps=5)
    out["strong"] = {"workload": f"ONE pool of {args.pool} rows per turn partitioned over {world} GPUs", "scaling": "strong",
                     "value": ms["value"], "value_iqr": ms["value_iqr"], "ms_per_step": ms["ms_per_step"],
                     "e2e": ms["e2e"], "e2e_iqr": ms["e2e_iqr"], "rows_per_gpu_per_turn": ms["n_local"],
                     "kernel_ms": ms["kernel_ms"]}
    del prop_s
    eng = prop._get_engine()
    # the accepted records cross NVLink: one all_gather_into_tensor per populate (gather_records)
    eng._pool_ok, eng._pool = False, None
    mg = measure(prop, worst, pool, args.steps, args.warmup, few, world, dev, kernel_reps=5)
    out["nccl_allgather"] = {"workload": "headline workload, accepted records all-gathered over NCCL (no shared host pool)",
                             "value": mg["value"], "ms_per_step": mg["ms_per_step"], "e2e": mg["e2e"], "e2e_iqr": mg["e2e_iqr"]}
    eng._pool_ok = None  # back to the shared host pool
    # same seed, N ranks vs one rank
    seed = 424242
    eng.set_seed(seed)
    dist.barrier()
    rows_n, p_n, a_n = eng.run(pool, pool, max_samples=pool)
    rows_n = np.array(rows_n, copy=True)
    verdict = None
    if rank == 0:
        e1 = PopulateEngine(prop.flow, eng.names, eng.row_dtype)
        e1.rank, e1.world, e1.group = 0, 1, None
        e1.configure(*eng._cfg_host, eng.log_prior_const, eng.r_max, eng.sqrt_t)
        e1.set_seed(seed)
        rows_1, p_1, a_1 = e1.run(pool, pool, max_samples=pool)
        verdict = {"seed": seed, "rows_n": int(len(rows_n)), "rows_1": int(len(rows_1)),
                   "n_proposed": [int(p_n), int(p_1)], "n_accepted": [int(a_n), int(a_1)],
                   "pool_bytes_equal": bool(len(rows_n) == len(rows_1) and rows_n.tobytes() == rows_1.tobytes())}
        del e1
    torch.cuda.synchronize()
    dist.barrier()
    out["parity_vs_n1"] = verdict
    out["c5"] = c5_variant("ours", world)  # collective: log_prob_all shards its rows over the ranks
    return out


def c5_levels(ModelClass, levels, n_train, n_draw, training_config, dev_sync=None):
    """Config C5 (BASELINE.json configs[4]): the flow work of the importance nested sampler, level by
    level as /root/reference/src/nessai/proposal/importance.py:250-440 drives it: add a flow
    (importance.py:80-99 of the flow model), train it on the level's weighted samples, draw the next
    level's samples from it, and evaluate EVERY stored flow on ALL samples drawn so far
    (``log_prob_all``: the meta-proposal density).  The same loop runs either arm: ``ModelClass`` is
    ``B200ImportanceFlowModel`` or the reference's ``ImportanceFlowModel``.  Returns per-phase seconds."""
    import torch

    sync = dev_sync or (lambda: None)
    torch.manual_seed(SEED)
    rng = np.random.default_rng(SEED)
    fm = ModelClass(flow_config=dict(n_inputs=D, n_neurons=64, n_blocks=4, n_layers=2),
                    training_config=dict(training_config), output=tempfile.mkdtemp(), rng=rng)
    fm.initialise()
    idx = np.arange(D)
    cov = 0.5 ** np.abs(idx[:, None] - idx[None, :])
    x_all = rng.multivariate_normal(np.zeros(D), 4.0 * cov, size=n_train)  # level 0: a broad prior-like set
    t = dict(train=0.0, draw=0.0, log_prob_all=0.0)
    rows_evaluated = 0
    for level in range(levels):
        fm.add_new_flow(reset=True)
        cur = x_all[-n_train:]
        w = rng.uniform(0.5, 1.5, size=len(cur))
        xs = (cur - cur.mean(0)) / cur.std(0)
        sync()
        t0 = time.perf_counter()
        fm.train(xs, weights=w, max_epochs=50, patience=50, plot=False,
                 output=os.path.join(fm.output, f"level_{level}"))
        sync()
        t1 = time.perf_counter()
        new = fm.sample_ith(level, N=n_draw)
        sync()
        t2 = time.perf_counter()
        x_all = np.concatenate([x_all, np.asarray(new, dtype=np.float64)])
        lp = fm.log_prob_all(x_all)
        sync()
        t3 = time.perf_counter()
        assert lp.shape == (len(x_all), level + 1)
        rows_evaluated += lp.size
        if level:  # the first level carries the one-off costs (allocations, module load)
            t["train"] += t1 - t0
            t["draw"] += t2 - t1
            t["log_prob_all"] += t3 - t2
    n = max(levels - 1, 1)
    return {"levels_timed": n, "train_s_per_level": t["train"] / n, "draw_s_per_level": t["draw"] / n,
            "log_prob_all_s_per_level": t["log_prob_all"] / n, "level_s": sum(t.values()) / n,
            "n_train": n_train, "n_draw": n_draw, "flow_evaluations_total": int(rows_evaluated),
            "finite_fraction_last": float(np.isfinite(lp).mean())}


def c5_variant(impl, world=1):
    """``variants.c5``: seconds per level of the importance sampler's flow work (see c5_levels)."""
    try:
        if impl == "reference":
            import oracle.refenv as refenv

            refenv.activate()
            from nessai.flowmodel.importance import ImportanceFlowModel

            out = c5_levels(ImportanceFlowModel, 3, 2000, 20_000, dict())
            out["note"] = "reference ImportanceFlowModel on the host cores; bounded: 3 levels, 2e4 draws per level"
            return out
        import torch

        from nessai_b200.importance import B200ImportanceFlowModel

        out = c5_levels(B200ImportanceFlowModel, 9, 2000, 200_000,
                        dict(device_tag=f"cuda:{torch.cuda.current_device()}"), dev_sync=torch.cuda.synchronize)
        out["note"] = (f"B200ImportanceFlowModel, 9 levels, 2e5 draws per level, 16-D RealNVP (default ResidualNet conditioner); "
                       f"log_prob_all rows sharded over {world} rank(s), training replicated")
        return out
    except Exception as e:  # noqa: BLE001 - a variant must never break the bench line
        return {"error": f"{type(e).__name__}: {e}"}


def train_variant(impl):
    """FlowModel.train on the C2 live points (scripts/train_bench.py) in a subprocess."""
    try:
        res = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "train_bench.py"), impl],
                             capture_output=True, text=True, timeout=600, cwd=REPO)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": f"exit {res.returncode}: {(res.stderr or res.stdout)[-400:]}"}
    except Exception as e:  # noqa: BLE001 - a variant must never break the bench line
        return {"error": f"{type(e).__name__}: {e}"}


def coupling_roofline(dev, peaks, which, n=8_000_000):
    """The element-wise coupling stage alone (conditioner output supplied), SURVEY.md 8(d):
    196 B/row at D = 16 (x in, shift|scale in, y out, log|det| out).  8e6 rows = 1.57 GB per
    launch, far beyond the 126 MB L2; CUDA events on the launching stream."""
    import torch

    from nessai_b200.coupling import coupling_transform

    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, D, device=dev, generator=g)
    p = torch.randn(n, D, device=dev, generator=g)
    tf = list(range(1, D, 2))
    for _ in range(3):
        coupling_transform(x, p, tf)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        coupling_transform(x, p, tf)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    gbs = n * 196.0 / (ms * 1e-3) / 1e9
    return {
        "kernel": "coupling_vec_kernel (AffineCouplingTransform forward, conditioner output supplied)",
        "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": gbs / peaks["hbm_gbs"], "peak_source": f"{which} copy bandwidth",
        "bytes_per_row": 196, "rows_per_launch": n, "kernel_ms": ms,
        "note": "includes the two torch.empty output allocations of the Python wrapper",
    }


def isolated_variants(pool):
    """The non-affine populate tail and accumulate_weights (scripts/tail_accumulate_variants.py),
    written after the round's GPU budget was spent: measured in a SUBPROCESS with its own CUDA
    context, after everything above, so a fault there cannot touch the headline numbers."""
    try:
        res = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "tail_accumulate_variants.py"), str(pool)],
                             capture_output=True, text=True, timeout=240, cwd=REPO)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": f"exit {res.returncode}: {(res.stderr or res.stdout)[-400:]}"}
    except Exception as e:  # noqa: BLE001 - a variant must never break the bench line
        return {"error": f"{type(e).__name__}: {e}"}


def c3_block(args, local_rank, dev, repeats, peaks, which):
    """Config C3 (BASELINE.json configs[2]): 32-D Rosenbrock, neural-spline flow with 6 coupling
    layers (ResidualNet 64, 8 bins) TRAINED BY THE REFERENCE on the live points
    (tests/golden/c3_nsf_trained.npz), ``populate(n_samples=2e6)`` through B200FlowProposal --
    the same three lenses as the headline, plus the reference's CPU populate on a bounded pool."""
    from nessai_b200.livepoint import numpy_array_to_live_points

    pool = args.c3_pool
    model, live = problem("c3_nsf_trained")
    live_s = numpy_array_to_live_points(live, model.names)
    live_s["logL"] = model.log_likelihood(live_s)
    worst = live_s[np.argmin(live_s["logL"])]
    prop = build_proposal("c3_nsf_trained", model, live_s, local_rank, pool)
    m = measure(prop, worst, pool, max(2, args.steps // 4), 3, max(3, repeats // 10), 1, dev, kernel_reps=5)
    k_rows_s = m["n_local"] / (m["kernel_ms"] * 1e-3)
    tf = k_rows_s * 0.50e6 / 1e12
    out = {
        "workload": f"C3: 32-D Rosenbrock, reference-trained NSF (6 coupling layers, ResidualNet 64, 8 bins), "
                    f"poolsize=drawsize={pool}, zscore, constant-volume radius 0.95, uniform box prior",
        "value": m["value"], "value_iqr": m["value_iqr"], "unit": "rows/s", "ms_per_step": m["ms_per_step"],
        "turns_per_step": m["turns_per_step"], "gpu_launches": m["launches"],
        "e2e": {"value": m["e2e"], "iqr": m["e2e_iqr"], "unit": "rows/s", "h2d_bytes_per_step": 4 * 32 * 8,
                "d2h_bytes_per_step": m["d2h_bytes"], "population_acceptance": m["population_acceptance"]},
        "roofline": {"kernel": "flow_tc_nsf_kernel x 6 (one launch per coupling layer: tcgen05 conditioner + RQ-spline epilogue)",
                     "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": tf / peaks["bf16_tflops"], "peak_source": f"{which} bf16 burst", "flops_per_row": 0.50e6,
                     "kernel_ms": m["kernel_ms"], "rows_per_launch": m["n_local"], "traffic": None,
                     "hbm": {"achieved": k_rows_s * 260.0 / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": k_rows_s * 260.0 / 1e9 / peaks["hbm_gbs"], "bytes_per_row": 260}},
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(threads=1, pool=50_000, fixture="c3_nsf_trained")
    return out


# ------------------------------------------------------------------------- reference
def reference_populate(threads, pool, steps=1, warmup=0, fixture="c2_realnvp_mlp"):
    """Time the UNMODIFIED reference's FlowProposal.populate on the host cores."""
    import oracle.refenv as refenv

    refenv.activate()
    import torch
    from nessai.livepoint import numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal

    torch.set_num_threads(threads)
    cfg, sd = load_fixture(fixture)
    ours, live = problem(fixture)

    class RefModel(Model):
        """The same problem as a nessai Model (names, bounds, uniform prior, likelihood)."""

        def __init__(self):
            self.names = list(ours.names)
            self.bounds = {n: list(b) for n, b in ours.bounds.items()}

        def log_prior(self, x):
            lp = np.log(self.in_bounds(x), dtype="float")
            for n in self.names:
                lp -= np.log(self.bounds[n][1] - self.bounds[n][0])
            return lp

        def log_likelihood(self, x):
            return ours.log_likelihood(x)

    model = RefModel()
    rng = np.random.default_rng(SEED)
    model.set_rng(rng)
    torch.manual_seed(SEED)
    prop = FlowProposal(
        model, rng=rng, flow_config=dict(cfg),
        output=tempfile.mkdtemp(), poolsize=pool, drawsize=pool, plot=False,
        fallback_reparameterisation="zscore",
    )
    prop.initialise()
    live_s = numpy_array_to_live_points(live, model.names)
    live_s["logL"] = model.log_likelihood(live_s)
    prop.check_state(live_s)
    prop.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    prop.flow.model.eval()
    worst = live_s[np.argmin(live_s["logL"])]
    # count proposed rows (populate keeps n_proposed in a local variable)
    counter = {"n": 0}
    draw = prop.sample_latent_distribution

    def counting_draw(n):
        z = draw(n)
        counter["n"] += len(z)
        return z

    prop.sample_latent_distribution = counting_draw
    for _ in range(warmup):
        prop.populate(worst, n_samples=pool, plot=False, max_samples=pool)
    prop.population_time *= 0
    counter["n"] = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        prop.populate(worst, n_samples=pool, plot=False, max_samples=pool)
    n_prop = counter["n"]
    wall = time.perf_counter() - t0
    pop_s = prop.population_time.total_seconds()
    return n_prop, pop_s, wall, prop.population_acceptance


def cpu_baseline(threads, pool, fixture="c2_realnvp_mlp"):
    import oracle.refenv as refenv

    if not refenv.reference_available():
        return {"value": None, "unit": "rows/s", "cores": threads, "kind": "reference",
                "sample": "baseline/_ref not present"}
    n_prop, pop_s, wall, acc = reference_populate(threads, pool, steps=1, warmup=1, fixture=fixture)
    return {
        "value": n_prop / pop_s,
        "unit": "rows/s",
        "cores": threads,
        "kind": "reference",
        "sample": f"one populate(n_samples={pool}, drawsize={pool}) = {n_prop} proposed rows in "
                  f"{pop_s:.1f} s after one warm-up populate; unmodified nessai FlowProposal (baseline/_ref) on the "
                  f"restated glasflow.nflows shim (oracle/shims), {fixture} weights / live points as the GPU arm",
        "population_acceptance": acc,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle.refenv as refenv

    if not refenv.reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref (pip --target install of /root/reference) not present"}))
        return
    threads = os.cpu_count() or 1
    pool = args.cpu_pool
    n_prop, pop_s, wall, acc = reference_populate(threads, pool, steps=args.steps, warmup=min(args.warmup, 1))
    v = n_prop / pop_s
    out = {
        "impl": "reference",
        "metric": "proposed live-points/sec (FlowProposal.populate, 16-D, pool 1e6)",
        "value": v,
        "unit": "rows/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * pop_s / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": "C2: 16-D RealNVP 4x[64,64] MLP, zscore, constant-volume radius 0.95, uniform box prior, "
                        f"poolsize=drawsize={pool} per populate()",
        },
        "cpu_baseline": {
            "value": v, "unit": "rows/s", "cores": threads, "kind": "reference",
            "sample": f"{args.steps} x populate(n_samples={pool}); nessai FlowProposal on the "
                      "restated glasflow.nflows shim, torch threads = all host cores",
        },
        "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "population_acceptance": acc,
    }
    if not args.no_extras:
        n3, s3, _, a3 = reference_populate(threads, 100_000, steps=1, warmup=1, fixture="c3_nsf_trained")
        out["variants"] = {"train": train_variant("reference"), "c5": c5_variant("reference"),
                           "c3": {"value": n3 / s3, "unit": "rows/s", "cores": threads, "population_acceptance": a3,
                                  "sample": f"one populate(n_samples=100000, drawsize=100000) = {n3} proposed rows in {s3:.1f} s "
                                            "after one warm-up; reference-trained 32-D NSF (c3_nsf_trained)"}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool", type=int, default=POOL)
    ap.add_argument("--cpu-pool", type=int, default=1_000_000,
                    help="poolsize = drawsize of the reference arm and of the 1-thread cpu_baseline (the GPU arm's 1e6)")
    ap.add_argument("--c3-pool", type=int, default=2_000_000)
    ap.add_argument("--min-steps", type=int, default=1000,
                    help="timed steps per lens in total (regions of --steps steps are repeated up to this)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary measurements (coupling roofline, ResidualNet variant): "
                         "the timed populate step only, e.g. for an ncu launch list")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
