TAG=r3w; . scripts/ab_r2n.sh
run scan_gb A=1
run scan_sep VKP_SCAN_IN_GBUFFER=0
