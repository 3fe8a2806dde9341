#!/bin/bash
# gpurun --gpus N job: the data-parallel bench (overlapped exchange, in-line exchange, NCCL) and the hand-run multi-GPU checks
set -u
N=${1:-2}
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --steps 200 --warmup 20 > $O/r2_bench_rir256_n$N.json 2> $O/r2_bench_rir256_n$N.err; echo "rc=$?"; tail -3 $O/r2_bench_rir256_n$N.err
$T bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_rir256_n${N}_s20.json 2>> $O/r2_bench_rir256_n$N.err; echo "rc=$?"
$T bench.py --gpus $N --steps 200 --warmup 20 --no-overlap > $O/r2_bench_rir256_n${N}_inline.json 2>> $O/r2_bench_rir256_n$N.err; echo "rc=$?"
python bench.py --steps 200 --warmup 20 --no-sweep --no-module --skip-cpu --skip-e2e > $O/r2_bench_rir256_n${N}_box1gpu.json 2>> $O/r2_bench_rir256_n$N.err
$T tools/mgpu_allreduce.py > $O/r2_mgpu_allreduce_n$N.log 2>&1; echo "allreduce rc=$?"; tail -4 $O/r2_mgpu_allreduce_n$N.log
$T tools/mgpu_module_dp.py > $O/r2_mgpu_module_dp_n$N.log 2>&1; echo "module dp rc=$?"; tail -3 $O/r2_mgpu_module_dp_n$N.log
python - <<PY
import json
for f in ["", "_s20", "_inline", "_box1gpu"]:
    try:
        d = json.load(open("$O/r2_bench_rir256_n$N%s.json" % f))
        print(f or "overlap", d["n_gpus"], round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us", d.get("collective_check"), d["kernels"])
    except Exception as e:
        print(f, "ERR", e)
PY
