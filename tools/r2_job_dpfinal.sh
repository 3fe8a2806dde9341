#!/bin/bash
# gpurun --gpus N job: the data-parallel bench line as the driver runs it (20 steps, 5 warm-up) and over 200 steps;
# N = 8 also: the side-stream schedule, NCCL, and configs[3] points under weak and strong scaling
set -u
N=${1:-2}
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
: > $O/r2_dp_n$N.err
show() { python - "$1" "$2" <<PY
import json, sys
try:
    txt = open(sys.argv[2]).read()
    d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    cc = d.get("collective_check") or {}
    print(sys.argv[1], "|", d["n_gpus"], "gpus", round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us", {k: (v["avg_us"], v["launches"]) for k, v in d["kernels"].items()},
          "check", cc.get("max_abs_diff"), cc.get("step_dE_max_abs_diff"), cc.get("checksums_identical_on_all_ranks"), cc.get("status"), "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 1))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
$T bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_rir256_n${N}.json 2>> $O/r2_dp_n$N.err; show "default s20 = the driver's command" $O/r2_bench_rir256_n${N}.json
$T bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e > $O/r2_bench_rir256_n${N}_run2.json 2>> $O/r2_dp_n$N.err; show "default s20 again" $O/r2_bench_rir256_n${N}_run2.json
$T bench.py --gpus $N --steps 200 --warmup 20 --skip-e2e > $O/r2_bench_rir256_n${N}_s200.json 2>> $O/r2_dp_n$N.err; show "default s200" $O/r2_bench_rir256_n${N}_s200.json
if [ "$N" = "8" ]; then
  $T bench.py --gpus $N --steps 200 --warmup 20 --skip-e2e --dp-mode side > $O/r2_bench_rir256_n${N}_side_s200.json 2>> $O/r2_dp_n$N.err; show "side s200" $O/r2_bench_rir256_n${N}_side_s200.json
  $T bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e --workload sweep_k4096_d128 > $O/r2_bench_sweep_k4096_d128_n${N}_weak.json 2>> $O/r2_dp_n$N.err; show "sweep 4096x128 weak" $O/r2_bench_sweep_k4096_d128_n${N}_weak.json
  $T bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e --workload sweep_k4096_d128 --strong > $O/r2_bench_sweep_k4096_d128_n${N}_strong.json 2>> $O/r2_dp_n$N.err; show "sweep 4096x128 strong" $O/r2_bench_sweep_k4096_d128_n${N}_strong.json
  $T bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e --workload sweep_k1024_d64 --strong > $O/r2_bench_sweep_k1024_d64_n${N}_strong.json 2>> $O/r2_dp_n$N.err; show "sweep 1024x64 strong" $O/r2_bench_sweep_k1024_d64_n${N}_strong.json
fi
grep -v "frame #\|^$\|OMP_NUM\|\*\*\*\*" $O/r2_dp_n$N.err | head -20
