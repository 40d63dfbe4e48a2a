for v in "" rt3 rt5; do
  if [ -n "$v" ]; then export VKP_LIB=$PWD/variants/$v/libvkp.so; else unset VKP_LIB; fi
  echo "variant [$v]"; scripts/scale.sh 1 cfg3 3 | grep cfg
done
