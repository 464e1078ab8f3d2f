#!/bin/bash
# one gpurun call: GPU tests + the default bench line (+ optional extras); outputs under gpurun_out/$1
OUT=gpurun_out/${1:-r2}
mkdir -p $OUT
python -m pytest tests -m gpu -q --durations=12 > $OUT/pytest_gpu.log 2>&1
tail -25 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err || tail -20 $OUT/bench_c2.err
python - <<PY
import json
d=json.load(open("$OUT/bench_c2.json"))
r=d["roofline"]
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"eager",round(d["eager_api"]["value"]),"e2e",round(d["e2e"]["value"]), "u8", round(d["e2e"]["uint8_io"]["value"]))
print("roofline",r["bound"],round(r["frac"],4),r["kernel"],r["pipe_utilisation_pct"])
print({k:round(v,3) for k,v in r["kernels_ms_per_step"].items()})
print("secondary",{k:(round(v["value"]) if isinstance(v,dict) and "value" in v else v) for k,v in d["secondary"].items() if k!="workload"})
print("cpu",d["cpu_baseline"]["value"],d["cpu_baseline"]["kind"],"link",d["e2e"]["link_probe"]["h2d_gbs_per_rank"],d["e2e"]["link_probe"]["d2h_gbs_per_rank"], d["e2e"]["link_probe"]["e2e_over_link_bound"])
PY
