#!/bin/bash
# A/B runs of bench.py under different environment knobs or library variants (run under gpurun, 1 GPU); one line per run.
#   scripts/ab.sh TAG [--chain N] [--workload cfg2|cfg3|cfg4|cfg5] [--steps K] NAME "ENV=VAL ENV2=VAL2" [NAME "..." ...]
# e.g. scripts/variants.sh gb5 "-DVKP_LB_GBUFFER=5" && scripts/ab.sh r9 base "A=1" gb5 "VKP_LIB=$PWD/variants/gb5/libvkp.so" nosky "VKP_SKY_CODES=0"
# Results: gpurun_out/TAG_NAME.json (the bench line), copied to profiles/ when DESIGN.md quotes them (section 4b).
TAG=$1; shift
CHAIN=""; WL=cfg2; STEPS=40
while [ "${1#--}" != "$1" ]; do
  case $1 in --chain) CHAIN="--chain $2"; shift 2;; --workload) WL=$2; shift 2;; --steps) STEPS=$2; shift 2;; *) echo "unknown option $1"; exit 2;; esac
done
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  NAME=$1; ENVS=$2; shift 2
  env $ENVS timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 5 --no-cpu-baseline --no-cfg3 $CHAIN > gpurun_out/${TAG}_$NAME.json 2> gpurun_out/${TAG}_$NAME.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_$NAME.json") if l.startswith("{")][-1])
    print("%-16s ms %.4f min %.4f  stages %s  chain stream %.3f  e2e %.4f" % ("$NAME", d["ms_per_step"], d["frame_ms"]["min"],
          {k[2:]: round(v, 3) for k, v in d["roofline"]["stage_ms_all"].items()}, d["roofline"].get("chain_stream_ms", 0), d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$NAME failed:", e, open("gpurun_out/${TAG}_$NAME.err").read()[-400:])
PY
done
