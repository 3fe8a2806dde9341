#!/bin/bash
set -u
N=${1:-2}
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
A="--steps 200 --warmup 20 --skip-e2e"
$T bench.py --gpus $N $A --no-graph > $O/r2_n${N}_eager_overlap.json 2> $O/r2_n${N}_b.err; echo "rc=$?"
$T bench.py --gpus $N $A --no-graph --no-overlap > $O/r2_n${N}_eager_inline.json 2>> $O/r2_n${N}_b.err; echo "rc=$?"
$T bench.py --gpus $N $A > $O/r2_n${N}_graph_overlap.json 2>> $O/r2_n${N}_b.err; echo "rc=$?"
python bench.py $A --no-sweep --no-module --skip-cpu --graph > $O/r2_n${N}_1gpu_graph.json 2>> $O/r2_n${N}_b.err
python bench.py $A --no-sweep --no-module --skip-cpu > $O/r2_n${N}_1gpu_eager.json 2>> $O/r2_n${N}_b.err
python - <<PY
import json
for f in ["eager_overlap", "eager_inline", "graph_overlap", "1gpu_graph", "1gpu_eager"]:
    try:
        txt = open("$O/r2_n${N}_%s.json" % f).read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(f, d["n_gpus"], round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us", {k: v["avg_us"] for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
