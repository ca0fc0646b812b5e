#!/bin/bash
# ONE chain over a SNP-sharded store at C4 (n=50,000 x p=1,000,000), N GPUs of one box:
#   tools/run_sharded_scaling.sh N      (under gpurun --gpus N)
# writes gpurun_out/shard_c4_n<N>.json (n_rao = 500, the testdata.ini setting) and _rao10.json (scan-heavy variant)
N=${1:-2}
PORT=$((29600 + N))
run() {
  if [ "$N" = "1" ]; then timeout 500 python bench.py "$@"
  else timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py "$@"; fi
}
run --gpus $N --sharded --workload C4 --steps 6 --warmup 3 > gpurun_out/shard_c4_n$N.json 2> gpurun_out/shard_c4_n$N.err
PORT=$((PORT + 20))
run --gpus $N --sharded --workload C4 --n-rao 10 --steps 50 --warmup 10 > gpurun_out/shard_c4_n${N}_rao10.json 2> gpurun_out/shard_c4_n${N}_rao10.err
python - <<PY
import json
for f in ("shard_c4_n$N", "shard_c4_n${N}_rao10"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, round(d["value"]), "it/s", round(d["ms_per_step"], 3), "ms/step; scan", round(d["roofline"]["avg_launch_ms"], 3), "ms", round(d["roofline"]["frac"], 3), d["breakdown"])
    except Exception as e:
        print(f, "ERR", e)
PY
