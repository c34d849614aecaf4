"""Turn gpurun_out/ ncu artefacts into the tracked text summaries under profiles/."""
import collections, csv, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("PROFILE_OUT") or os.path.join(REPO, "profiles")  # PROFILE_OUT: summarise on the GPU box
G = os.path.join(REPO, "gpurun_out")

def launches(path, dst, title):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("kernel,launches,mean_us,total_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k[:100]}\",{len(v)},{sum(v)/len(v)/1e3:.1f},{sum(v)/1e3:.1f},{sum(v)/tot:.4f}\n")

def full(rep, dst, title, top=30):
    a = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "ncu_source.py"), rep, str(top)], capture_output=True, text=True).stdout
    with open(dst, "w") as f:
        f.write(f"# {title}\n# ncu --set full --clock-control none --import-source on (one launch)\n\n## raw metrics\n{a}\n## top stalled SASS instructions (warp stall samples)\n{b}")

os.makedirs(OUT, exist_ok=True)
jobs = [
    ("launches_r1_generic.csv", "r1_launches_generic_fp32.csv", launches, "bench.py step, generic fp32 kernel (before the tcgen05 kernel)"),
    ("launches_r1_tc.csv", "r1_launches_tcgen05.csv", launches, "bench.py step, tcgen05 populate kernel"),
    ("prof_r1_generic.ncu-rep", "r1_ncu_populate_generic_fp32.txt", full, "populate_draw_kernel<relu> (generic fp32 interpreter), 1e6 rows"),
    ("prof_r1_tc_pop_v1.ncu-rep", "r1_ncu_populate_tcgen05_v1_smemA.txt", full, "flow_tc_populate_kernel v1 (A operand in shared memory, 2 groups), 1e6 rows"),
    ("prof_r1_tc_pop_v2.ncu-rep", "r1_ncu_populate_tcgen05_v2_tmemA.txt", full, "flow_tc_populate_kernel v2 (A operand in TMEM, 4 groups), 1e6 rows"),
    ("prof_r1_tc_pop_v3.ncu-rep", "r1_ncu_populate_tcgen05_v3.txt", full, "flow_tc_populate_kernel v3 (fp32 x' output, smem constants, suspended waits), 1e6 rows"),
    ("prof_r1_tc_pop_v6.ncu-rep", "r1_ncu_populate_tcgen05_v6_converged_issuer.txt", full, "flow_tc_populate_kernel v6 (converged issuer warps, affine folded into GEMM1, single bias MMA), 1e6 rows"),
    ("prof_r1_tc_pop_v7.ncu-rep", "r1_ncu_populate_tcgen05_v7_fp16_split.txt", full, "flow_tc_populate_kernel v7 (fp16 hi/lo split operands, FFMA2 affine, bias MMA), 1e6 rows"),
    ("prof_r1_tc_nsf_v1.ncu-rep", "r1_ncu_nsf_tcgen05_v1.txt", full, "flow_tc_nsf_kernel<0> (C3: 32-D spline flow, one layer of 6, 2e6 rows)"),
    ("prof_r1_tc_nsf_v2.ncu-rep", "r1_ncu_nsf_tcgen05_v2_double_buffered_chunks.txt", full, "flow_tc_nsf_kernel<0> v2 (final-layer chunks double-buffered across D2 / D), one layer of 6, 2e6 rows"),
    ("prof_r1_accept_fused.ncu-rep", "r1_ncu_accept_fused.txt", full, "accept_fused_kernel (rejection step + in-order compaction, single pass), 1e6 rows"),
    ("prof_r1_tc_res_v1.ncu-rep", "r1_ncu_populate_tcgen05_resnet_v1.txt", full, "flow_tc_res_kernel<1> (ResidualNet conditioner, last of 2 layer passes), 1e6 rows"),
    ("prof_r1_coupling.ncu-rep", "r1_ncu_coupling_transform.txt", full, "coupling_vec_kernel<4> (affine coupling transform alone, 8e6 rows x 196 B)"),
    ("prof_r1_tc_v2.ncu-rep", "r1_ncu_apply_tcgen05_v2.txt", full, "flow_tc_apply_kernel v2 (FlowModel.inverse, z supplied), 1e6 rows"),
]
jobs += [  # round 2
    ("launches_r2.csv", "r2_launches_bench_step.csv", launches, "bench.py steps (C2 MLP, populate through PopulateEngine.run), round 2"),
    ("launches_r2_train.csv", "r2_train_launches.csv", launches, "FlowModel.train on C2 (persistent kernel: one launch per chunk of epochs), round 2"),
    ("prof_r2_tc_pop.ncu-rep", "r2_ncu_populate_tcgen05_affmma.txt", full, "flow_tc_populate_kernel, round 2 (affine on the tensor core: GEMM1 N = 80), 1e6 rows"),
    ("prof_r2_tc_res.ncu-rep", "r2_ncu_populate_tcgen05_resnet.txt", full, "flow_tc_res_kernel<1> round 2 (fp16-split build, affine on the tensor core), ResidualNet conditioner, 1e6 rows"),
    ("prof_r2_tc_nsf.ncu-rep", "r2_ncu_nsf_tcgen05.txt", full, "flow_tc_nsf_kernel<1> (C3: reference-trained 32-D spline flow, one layer of 6), 2e6 rows"),
    ("prof_r2_tail.ncu-rep", "r2_ncu_reparam_tail.txt", full, "reparam_tail_kernel (logit on 8 of 16 parameters; coalesced tile IO), 1e6 rows"),
    ("prof_r2_sumexp.ncu-rep", "r2_ncu_sum_exp.txt", full, "sum_exp_kernel, 8e6 rows"),
    ("prof_r2_accept64.ncu-rep", "r2_ncu_accept_x64.txt", full, "accept_fused_kernel, float64-row flavour, 1e6 rows"),
    ("prof_r2_train.ncu-rep", "r2_ncu_train_persistent.txt", full, "tr_train_kernel (persistent cooperative training kernel, 8 epochs x 2 steps, C2 MLP)"),
]
for src, dst, fn, title in jobs:
    p = os.path.join(G, src)
    if os.path.exists(p):
        fn(p, os.path.join(OUT, dst), title)
        print("wrote", dst)
for j in ("bench_r2_n1.json", "bench_r2_ref.json", "gputest_r2.log", "tr_fine.txt",
          "bench_r1_n1.json", "bench_r1_ref.json", "bench_r1_n2.json", "bench_r1_n4.json", "bench_r1_n8.json",
          "r1_tc_phase_timeline.txt"):
    p = os.path.join(G, j)
    if os.path.exists(p):
        open(os.path.join(OUT, j), "w").write(open(p).read())
