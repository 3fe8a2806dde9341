#!/bin/bash
# gpurun --gpus N job: ONE run of the driver's data-parallel command (kept short: N x box time is charged)
set -u
N=${1:-4}
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e > $O/r2_bench_rir256_n${N}.json 2> $O/r2_dp_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r2_bench_rir256_n${N}.json").read().splitlines() if l.startswith("{")][-1])
    cc = d.get("collective_check") or {}
    print(d["n_gpus"], "gpus", round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us", {k: v["avg_us"] for k, v in d["kernels"].items()}, cc.get("max_abs_diff"), cc.get("checksums_identical_on_all_ranks"), cc.get("status"))
except Exception as e:
    print("ERR", e)
PY
grep -v "frame #\|^$\|OMP_NUM\|\*\*\*\*" $O/r2_dp_n$N.err | head -8
