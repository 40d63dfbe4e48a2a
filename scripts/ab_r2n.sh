run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-cfg3 --chain 5 > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err
  python - <<PY
import json
for l in open("gpurun_out/r2o_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("%-14s ms %.4f min %.4f stages %s e2e %.4f" % ("$name", d["ms_per_step"], d["frame_ms"]["min"], {k[2:]:round(v,3) for k,v in d["roofline"]["stage_ms_all"].items()}, d["e2e"]["ms_per_step"]))
PY
}
run base A=1
run ctas1 VKP_CHAIN_CTAS=1
run ctas2 VKP_CHAIN_CTAS=2
run lb4 VKP_LIB=$PWD/variants/chain_lb4/libvkp.so
run lb5 VKP_LIB=$PWD/variants/chain_lb5/libvkp.so
run lb6 VKP_LIB=$PWD/variants/chain_lb6/libvkp.so
run lb5_ctas2 VKP_LIB=$PWD/variants/chain_lb5/libvkp.so VKP_CHAIN_CTAS=2
run lb5_ctas3 VKP_LIB=$PWD/variants/chain_lb5/libvkp.so VKP_CHAIN_CTAS=3
run lb6_ctas3 VKP_LIB=$PWD/variants/chain_lb6/libvkp.so VKP_CHAIN_CTAS=3
