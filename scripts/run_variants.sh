#!/bin/bash
# run the bench for every built variant (under gpurun): VKP_LIB picks the library
for d in variants/*/; do
  n=$(basename $d)
  VKP_LIB=$PWD/variants/$n/libvkp.so python bench.py --workload ${WORKLOAD:-cfg2} --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --no-cfg3 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$n.json")); print("%-12s ms/frame %.4f (min %.4f)  stages %s  e2e %.3f" % ("$n", d["ms_per_step"], d["frame_ms"]["min"], {k[2:]:round(v,4) for k,v in d["roofline"]["stage_ms_all"].items()}, d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$n failed", e, open("gpurun_out/var_$n.err").read()[-500:])
PY
done
