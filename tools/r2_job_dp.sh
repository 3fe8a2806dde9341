#!/bin/bash
# gpurun --gpus N job: the data-parallel bench in its scheduling modes (+ a 20-step run, the driver's likely setting)
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
$T bench.py --gpus $N --steps 200 --warmup 20 > $O/r2_bench_rir256_n$N.json 2>> $O/r2_dp_n$N.err; show "tail graph s200 (default)" $O/r2_bench_rir256_n$N.json
$T bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e > $O/r2_bench_rir256_n${N}_s20.json 2>> $O/r2_dp_n$N.err; show "tail graph s20" $O/r2_bench_rir256_n${N}_s20.json
$T bench.py --gpus $N --steps 200 --warmup 20 --skip-e2e --no-graph > $O/r2_dp_x.json 2>> $O/r2_dp_n$N.err; show "tail eager" $O/r2_dp_x.json
$T bench.py --gpus $N --steps 200 --warmup 20 --skip-e2e --dp-mode inline > $O/r2_bench_rir256_n${N}_inline.json 2>> $O/r2_dp_n$N.err; show "inline graph" $O/r2_bench_rir256_n${N}_inline.json
if [ "${2:-}" = "all" ]; then
  $T bench.py --gpus $N --steps 200 --warmup 20 --skip-e2e --dp-mode side > $O/r2_dp_x.json 2>> $O/r2_dp_n$N.err; show "side graph" $O/r2_dp_x.json
  $T bench.py --gpus $N --steps 200 --warmup 20 --skip-e2e --dp-mode deferred > $O/r2_dp_x.json 2>> $O/r2_dp_n$N.err; show "deferred graph" $O/r2_dp_x.json
  $T bench.py --gpus $N --steps 200 --warmup 20 --skip-e2e --nccl > $O/r2_bench_rir256_n${N}_nccl.json 2>> $O/r2_dp_n$N.err; show "nccl" $O/r2_bench_rir256_n${N}_nccl.json
  $T tools/mgpu_allreduce.py > $O/r2_mgpu_allreduce_n$N.log 2>&1; echo "allreduce rc=$?"; tail -n 4 $O/r2_mgpu_allreduce_n$N.log
  $T tools/mgpu_module_dp.py > $O/r2_mgpu_module_dp_n$N.log 2>&1; echo "module dp rc=$?"; tail -n 3 $O/r2_mgpu_module_dp_n$N.log
fi
python bench.py --steps 200 --warmup 20 --no-sweep --no-module --skip-cpu --skip-e2e > $O/r2_dp_x.json 2>> $O/r2_dp_n$N.err; show "1 gpu of this box" $O/r2_dp_x.json
grep -v "frame #\|^$\|OMP_NUM\|\*\*\*\*" $O/r2_dp_n$N.err | head -20
