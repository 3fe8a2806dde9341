#!/bin/bash
# One gpurun call that produces the round-2 evidence under gpurun_out/ (copied to profiles/ afterwards):
# bench lines of the workloads, the sweep, the ncu launch list and ncu --set full captures of the current kernels.
set -u
O=gpurun_out
B="python bench.py --steps 20 --warmup 5"
$B > $O/r2_bench_rir256.json 2> $O/r2_bench_rir256.err
$B --sweep-all --skip-e2e --no-module --skip-cpu > $O/r2_bench_rir256_sweep_all.json 2>> $O/r2_bench_rir256.err
for w in speech32 echoed64 loc16; do $B --workload $w --no-sweep --no-module --skip-e2e > $O/r2_bench_$w.json 2>> $O/r2_bench_rir256.err; done
$B --no-onehot --no-sweep --no-module --skip-e2e --skip-cpu > $O/r2_bench_rir256_noonehot.json 2>> $O/r2_bench_rir256.err
S="python bench.py --steps 2 --warmup 3 --no-sweep --no-module --skip-cpu --skip-e2e"
$S > $O/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_launches_rir256.csv $S > $O/r2_ncu_launches.log 2>&1
$S > $O/r2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'vq_screen|backward_kernel|prep_codebook' -s 9 -c 3 -o $O/r2_full_rir256 $S > $O/r2_ncu_full.log 2>&1
for w in sweep_k1024_d64 sweep_k4096_d128; do
  $S --workload $w > $O/r2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'vq_screen|backward_' -s 6 -c 2 -o $O/r2_full_$w $S --workload $w > $O/r2_ncu_full_$w.log 2>&1
done
ls -la $O | tail -20
