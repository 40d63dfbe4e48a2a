TAG=r3t; . scripts/ab_r2n.sh
run vl4 A=1
run novl VKP_CHAIN_VERY_LONG=0
