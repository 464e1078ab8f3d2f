# per-kernel milliseconds of one bench configuration: bash tools/bench_kernels.sh <bench.py args...>
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" > /tmp/o.json 2> /tmp/o.err || tail -5 /tmp/o.err
python -c "
import json;d=json.load(open('/tmp/o.json'));print('$*',round(d['value']),round(d['ms_per_step'],2),{k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items() if v>0.1})"
