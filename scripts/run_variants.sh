#!/bin/bash
# run the bench for every built variant (under gpurun)
for d in variants/*/; do
  n=$(basename $d)
  VKP_LIB=$PWD/variants/$n/libvkp.so python bench.py --workload ${WORKLOAD:-cfg2} --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$n.json")); print("%-10s ms/frame %.4f  stages %s  e2e %.3f" % ("$n", d["ms_per_step"], {k[2:]:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$n failed", e, open("gpurun_out/var_$n.err").read()[-500:])
PY
done
