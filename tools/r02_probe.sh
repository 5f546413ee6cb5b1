for lib in dust_b200/lib/libdust_b200.so dust_b200/lib/var/binary.so dust_b200/lib/var/wide4.so; do
  echo $lib; DUST_B200_LIB=$PWD/$lib python tools/gpu_probe.py wavefront 2 2>&1 | grep "dust dbg" | tail -3
done
