set -x
for shape in "128 128 1 1 80" "64 64 3 1 80" "128 128 1 1 20" "128 64 3 1 80"; do
  tag=$(echo $shape | tr ' ' '_')
  python tools/conv_probe.py $shape 64 5 2 > gpurun_out/probe_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 3 -c 1 -f -o gpurun_out/s6_conv_$tag python tools/conv_probe.py $shape 64 5 2 > gpurun_out/ncu_$tag.log 2>&1
  cat gpurun_out/probe_$tag.log | tail -1
done
