#!/bin/sh
# Multi-GPU evidence of one round, on a box with 8 GPUs (gpurun --gpus 8): writes gpurun_out/scaling_<tag>.jsonl
# (bench.py lines at N = 2, 4, 8 and the north-star shape 1 024 chains = 8 x 128) and the 2-GPU sharded == single test.
TAG=${1:-r02}
OUT=gpurun_out/scaling_$TAG.jsonl
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi_test_$TAG.log 2>&1; tail -2 gpurun_out/multi_test_$TAG.log
for n in 2 4; do
  $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-nrpt 2>gpurun_out/scaling_${TAG}_n$n.err | tail -1 >> $OUT
done
$TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/scaling_${TAG}_n8.err | tail -1 >> $OUT
$TR --nproc-per-node 8 --master-port 29509 bench.py --gpus 8 --steps 10 --warmup 3 --chains-per-gpu 128 --no-cpu --no-nrpt 2>gpurun_out/scaling_${TAG}_n8x128.err | tail -1 >> $OUT
$TR --nproc-per-node 8 --master-port 29510 tools/bench_sharded.py C4 2>gpurun_out/scaling_${TAG}_c4.err | tail -1 >> $OUT
python - <<PY
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: print("unparsed:", l[:200]); continue
    if "metric" in d:
        print("N=%d  %s  value %.0f  ms/step %.3f  ms/launch %.3f  e2e %s  round_trips/s %s" % (d["n_gpus"], d["config"]["workload"][:60], d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["e2e"]["value"] if d.get("e2e") else None, d.get("round_trips_per_s")))
    else: print(json.dumps(d)[:300])
PY
