#!/bin/bash
# gpurun --gpus N job: the default data-parallel bench line (+ a 20-step run, the driver's likely setting)
set -u
N=${1:-2}
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --steps 200 --warmup 20 > $O/r2_bench_rir256_n$N.json 2> $O/r2_bench_rir256_n$N.err; echo "rc=$?"
$T bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e > $O/r2_bench_rir256_n${N}_s20.json 2>> $O/r2_bench_rir256_n$N.err; echo "rc=$?"
python - <<PY
import json
for f in ["", "_s20"]:
    try:
        txt = open("$O/r2_bench_rir256_n$N%s.json" % f).read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(f or "s200", d["n_gpus"], round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us", {k: v["avg_us"] for k, v in d["kernels"].items()}, d["collective_check"], "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 1))
    except Exception as e:
        print(f, "ERR", e)
PY
grep -v "frame #\|^$\|OMP_NUM\|\*\*\*\*" $O/r2_bench_rir256_n$N.err | head -20
