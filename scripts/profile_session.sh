set -x
mkdir -p gpurun_out/prof
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q > gpurun_out/gputest_r2.log 2>&1; tail -3 gpurun_out/gputest_r2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 300 gpurun_out/bench_r2_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --min-steps 4 --no-extras --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r2_train.csv python scripts/train_bench.py ours > /dev/null 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:flow_tc_populate_kernel -s 8 -c 1 -o gpurun_out/prof_r2_tc_pop -f python scripts/ubench/ab_populate.py > /dev/null 2>&1
$N -k regex:flow_tc_res_kernel -s 17 -c 1 -o gpurun_out/prof_r2_tc_res -f python scripts/ubench/ab_populate.py > /dev/null 2>&1
$N -k regex:reparam_tail_kernel -s 4 -c 1 -o gpurun_out/prof_r2_tail -f python scripts/tail_accumulate_variants.py > /dev/null 2>&1
$N -k regex:sum_exp_kernel -s 1 -c 1 -o gpurun_out/prof_r2_sumexp -f python scripts/tail_accumulate_variants.py > /dev/null 2>&1
$N -k regex:tr_train_kernel -s 2 -c 1 -o gpurun_out/prof_r2_train -f python scripts/train_prof_short.py > /dev/null 2>&1
$N -k regex:flow_tc_nsf_kernel -s 8 -c 1 -o gpurun_out/prof_r2_tc_nsf -f python scripts/nsf_prof.py > /dev/null 2>&1
# summarise on the box and drop the raw reports (gpurun merges at most 64 MiB back)
PROFILE_OUT=gpurun_out/prof python scripts/make_profile_summary.py
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out gpurun_out/prof | tail -30
