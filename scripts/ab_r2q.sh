TAG=r3z; . scripts/ab_r2n.sh
run sky A=1
run nosky VKP_SKY_CODES=0
run sky2 A=1
