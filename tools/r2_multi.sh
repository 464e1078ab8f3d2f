#!/bin/bash
# multi-GPU check (gpurun --gpus N): two-rank bitwise test over NCCL, then the bench line at N ranks
N=${1:-2}; OUT=gpurun_out/multi$N; mkdir -p $OUT
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_round2.py -m gpu -q -k "two_rank" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err || tail -20 $OUT/bench_n$N.err
python - <<PY
import json
d=json.load(open("$OUT/bench_n$N.json"))
print("N",d["n_gpus"],d["config"]["name"],"value",round(d["value"]),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"]),"u8",round(d["e2e"]["uint8_io"]["value"]))
print("link",d["e2e"]["link_probe"]["h2d_gbs_per_rank"],d["e2e"]["link_probe"]["d2h_gbs_per_rank"],round(d["e2e"]["link_probe"]["e2e_over_link_bound"],3))
print("secondary",{k:(round(v["value"]) if isinstance(v,dict) and "value" in v else v) for k,v in d["secondary"].items() if k!="workload"})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | cut -c1-300
