run5() { wl=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3q_${wl}_$name.json 2> gpurun_out/r3q_${wl}_$name.err
  python - <<PY
import json
for l in open("gpurun_out/r3q_${wl}_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("%-6s %-10s ms %.3f stages %s" % ("$wl", "$name", d["ms_per_step"], {k[2:]:round(v,3) for k,v in d["roofline"]["stage_ms_all"].items()}))
PY
}
run5 cfg5 rf12 A=1
run5 cfg5 rf8 VKP_LIB=$PWD/variants/rf8/libvkp.so
run5 cfg5 rf20 VKP_LIB=$PWD/variants/rf20/libvkp.so
run5 cfg5 rf12s12 VKP_LIB=$PWD/variants/rf12s12/libvkp.so
run5 cfg5 rf16lb5 VKP_LIB=$PWD/variants/rf16lb5/libvkp.so
