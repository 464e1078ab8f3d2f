# A/B of library builds / configurations on the same box: each argument is an env assignment list "A=1,B=2"
for E in "$@"; do
env $(echo $E | tr ',' ' ') timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary $BENCH_ARGS > /tmp/o.json 2> /tmp/o.err; python -c "
import json;d=json.load(open('/tmp/o.json'));print('$E',round(d['value']),round(d['ms_per_step'],2),{k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items() if v>0.1})"
done
