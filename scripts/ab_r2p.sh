TAG=r2p; . scripts/ab_r2n.sh
run base A=1
run noseed VKP_SEED_PLANE=0
run gb3 VKP_LIB=$PWD/variants/gb3/libvkp.so
run gb5 VKP_LIB=$PWD/variants/gb5/libvkp.so
run gb6 VKP_LIB=$PWD/variants/gb6/libvkp.so
run res4 VKP_LIB=$PWD/variants/res4/libvkp.so
run res6 VKP_LIB=$PWD/variants/res6/libvkp.so
