#!/bin/bash
# 1-GPU job: the sweep with the row epilogue fused / split off, and the GPU tests with the split forced
set -u
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --sweep-all --skip-e2e --no-module --skip-cpu"
B200VQ_SPLIT_ROWS=0 $B > $O/r2_sweep_fused.json 2> $O/r2_sweep.err
B200VQ_SPLIT_ROWS=1 $B > $O/r2_sweep_split.json 2>> $O/r2_sweep.err
$B > $O/r2_sweep_auto.json 2>> $O/r2_sweep.err
B200VQ_SPLIT_ROWS=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_backward.py -m gpu -x -q > $O/r2_pytest_split.log 2>&1; echo "pytest(split) rc=$?"; tail -3 $O/r2_pytest_split.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest_gpu.log
python - <<PY
import json
r = {}
for f in ["fused", "split", "auto"]:
    try:
        d = json.loads([l for l in open("$O/r2_sweep_%s.json" % f).read().splitlines() if l.startswith("{")][-1])
        r[f] = {(p["K"], p["D"]): p for p in d["sweep"]["points"]}
    except Exception as e:
        print(f, "ERR", e)
for key in sorted(r.get("fused", {})):
    print(key, " ".join("%s fwd %.0f bwd %.0f frac %.3f |" % (f, r[f][key]["fwd_us"], r[f][key]["bwd_us"], r[f][key]["frac"]) for f in r if key in r[f]), r.get("auto", {}).get(key, {}).get("forward_kernels"))
PY
tail -5 $O/r2_sweep.err

A="--steps 200 --warmup 20 --no-sweep --no-module --skip-cpu --skip-e2e"
for m in "--graph" "--emulate-dp" "--emulate-dp --no-overlap" "--emulate-dp --no-graph" "--emulate-dp --no-graph --no-overlap"; do
  python bench.py $A $m > $O/r2_emu.json 2>> $O/r2_sweep.err
  python - "$m" <<PY
import json, sys
d = json.loads([l for l in open("$O/r2_emu.json").read().splitlines() if l.startswith("{")][-1])
print(sys.argv[1], round(d["ms_per_step"] * 1e3, 2), "us", {k: v["avg_us"] for k, v in d["kernels"].items()})
PY
done
