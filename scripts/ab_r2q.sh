TAG=r4f; . scripts/ab_r2n.sh
run lb8 A=1
run lb7 VKP_LIB=$PWD/variants/sh7/libvkp.so
run lb6 VKP_LIB=$PWD/variants/sh6/libvkp.so
run lb10 VKP_LIB=$PWD/variants/sh10/libvkp.so
