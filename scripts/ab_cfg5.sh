run5() { wl=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3j_${wl}_$name.json 2> gpurun_out/r3j_${wl}_$name.err
  python - <<PY
import json
for l in open("gpurun_out/r3j_${wl}_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("%-6s %-10s ms %.3f stages %s" % ("$wl", "$name", d["ms_per_step"], {k[2:]:round(v,3) for k,v in d["roofline"]["stage_ms_all"].items()}))
PY
}
run5 cfg5 keepw A=1
run5 cfg5 nokeepw VKP_LIB=$PWD/variants/nokeepw/libvkp.so
run5 cfg3 keepw A=1
run5 cfg3 nokeepw VKP_LIB=$PWD/variants/nokeepw/libvkp.so
