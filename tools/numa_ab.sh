#!/bin/bash
# gpurun --gpus N: e2e of N ranks with and without the NUMA binding of bench.py (BENCH_NUMA), C2 shape per rank to keep it short
N=${1:-4}; OUT=gpurun_out/numa$N; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" > $OUT/lscpu.txt
for NUMA in 1 0 1 0; do
  BENCH_NUMA=$NUMA python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$NUMA bench.py --gpus $N --config C2 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/o.json 2> $OUT/o.err || tail -20 $OUT/o.err
  python - <<PY | tee -a $OUT/ab.log
import json
d=json.load(open("$OUT/o.json"))
lp=d["e2e"]["link_probe"]
print("NUMA=$NUMA N",d["n_gpus"],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"u8",round(d["e2e"]["uint8_io"]["value"]),"h2d",lp["h2d_gbs_per_rank"],"d2h",lp["d2h_gbs_per_rank"],"nodes",lp.get("numa_node_per_rank"),"over_bound",round(lp["e2e_over_link_bound"],3))
PY
done
cat $OUT/lscpu.txt
