#!/bin/bash
# the last 1-GPU job of round 2: GPU tests, smoke(), the default bench line with the driver's arguments, the full sweep,
# the other workloads, the reference arm on the box's host cores
set -u
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/r2_pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/r2_smoke_final.log
B="python bench.py --steps 20 --warmup 5"
timeout 200 $B > $O/r2_bench_rir256_final.json 2> $O/r2_final.err; echo "bench rc=$?"
timeout 120 $B --sweep-all --skip-e2e --no-module --skip-cpu > $O/r2_bench_rir256_sweep_all_final.json 2>> $O/r2_final.err
for w in speech32 echoed64 loc16; do timeout 60 $B --workload $w --no-sweep --no-module --skip-e2e --skip-cpu > $O/r2_bench_${w}_final.json 2>> $O/r2_final.err; done
timeout 60 $B --no-onehot --no-sweep --no-module --skip-e2e --skip-cpu > $O/r2_bench_rir256_noonehot_final.json 2>> $O/r2_final.err
timeout 100 $B --impl reference > $O/r2_bench_rir256_reference_arm.json 2>> $O/r2_final.err
python - <<PY
import json
def rd(f):
    return json.loads([l for l in open("$O/" + f).read().splitlines() if l.startswith("{")][-1])
try:
    d = rd("r2_bench_rir256_final.json")
    print("final", round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us frac", d["roofline"]["frac"], "step", d["roofline"]["step"], "traffic", d["roofline"]["traffic"])
    print("  e2e", round(d["e2e"]["value"] / 1e6, 1), "lean", round(d["e2e_lean"]["value"] / 1e6, 1), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "module", d["module"], "tf32", d["peaks"]["tf32"])
    print("  kernels", d["kernels"])
    for p in d["sweep"]["points"]: print("  corner", p)
except Exception as e:
    print("final ERR", e)
try:
    d = rd("r2_bench_rir256_sweep_all_final.json")
    for p in d["sweep"]["points"]: print("  sweep", p["K"], p["D"], p["fwd_us"], p["bwd_us"], p["frac_fwd"], p["bwd_frac_hbm"], p["frac"], p["backward_path"])
except Exception as e:
    print("sweep ERR", e)
for f in ["speech32", "echoed64", "loc16", "rir256_noonehot"]:
    try:
        d = rd("r2_bench_%s_final.json" % f)
        print(f, round(d["value"] / 1e6, 1), "M", round(d["ms_per_step"] * 1e3, 2), "us frac", d["roofline"]["frac"], d["roofline"]["step"]["frac_of_hbm"], {k: v["avg_us"] for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
try:
    d = rd("r2_bench_rir256_reference_arm.json")
    print("reference arm", d["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], d["config"]["rows_per_gpu"])
except Exception as e:
    print("reference ERR", e)
PY
tail -n 5 $O/r2_final.err
