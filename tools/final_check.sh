#!/bin/bash
# What the driver runs at round end, on one box: GPU tests, smoke(), the default bench line and the reference arm.
O=gpurun_out/final; mkdir -p $O
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-400 $O/bench_reference.json
python bench.py > $O/bench_default.json 2> $O/bench_default.err; python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"u8",round(d["e2e"]["uint8_io"]["value"]),"frac",round(d["roofline"]["frac"],4),"launches",d["gpu_launches"],"clocks",d["clocks"],"steps",d["steps"],d["warmup"])
PY
