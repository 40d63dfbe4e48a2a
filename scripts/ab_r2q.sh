TAG=r3m; . scripts/ab_r2n.sh
runc() { c=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-cfg3 --chain $c > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("%-14s ms %.4f min %.4f stages %s chain %.3f e2e %.4f" % ("$name", d["ms_per_step"], d["frame_ms"]["min"], {k[2:]:round(v,3) for k,v in d["roofline"]["stage_ms_all"].items()}, d["roofline"].get("chain_stream_ms",0), d["e2e"]["ms_per_step"]))
PY
}
runc 7 c7 A=1
runc 8 c8 A=1
