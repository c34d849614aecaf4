# ncu --set full of the kernels added late in round 2: the narrow ResidualNet kernel on the reference's default
# flow (D = 16, width 32) and the affine-coupling mode of the tile kernel (D = 20, width 40); summarised on the box.
set -x
mkdir -p gpurun_out/prof
N="ncu --set full --clock-control none --import-source on"
$N -k regex:flow_tc_res_kernel -s 5 -c 1 -o gpurun_out/prof_late_res_narrow -f python scripts/ubench/default_width.py 16 > /dev/null 2>&1
$N -k regex:flow_tc_nsf_kernel -s 9 -c 1 -o gpurun_out/prof_late_ac -f python scripts/ubench/default_width.py 20 > /dev/null 2>&1
for k in res_narrow ac; do
  { echo "# $k"; python scripts/ncu_summary.py gpurun_out/prof_late_$k.ncu-rep; echo; echo "## top stalled SASS instructions (warp stall samples)"; python scripts/ncu_source.py gpurun_out/prof_late_$k.ncu-rep 25; } > gpurun_out/prof/late_$k.txt 2>&1
done
rm -f gpurun_out/*.ncu-rep
head -30 gpurun_out/prof/late_res_narrow.txt
