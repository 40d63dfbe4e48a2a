#!/bin/bash
# usage: scripts/scale.sh N workload steps   (run under gpurun --gpus N)
N=$1; WL=${2:-cfg2}; STEPS=${3:-20}
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/scale_${WL}_n$N.json 2> gpurun_out/scale_${WL}_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --steps $STEPS --warmup 3 $EXTRA > gpurun_out/scale_${WL}_n$N.json 2> gpurun_out/scale_${WL}_n$N.err
fi
python - <<PY
import json
for line in open("gpurun_out/scale_${WL}_n$N.json"):
    if line.startswith("{"):
        d=json.loads(line); print("$WL N=$N value %.1f ms/step %.4f e2e %.4f verified %s stages %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("composite_bit_identical_to_1gpu"), d["roofline"]["stage_ms"]))
PY
tail -3 gpurun_out/scale_${WL}_n$N.err
