# norm_bwd probe on a GroupNorm-shaped and a BatchNorm-shaped tensor + ncu --set full of the two kernels
python tools/probe_norm.py 128 80 80 64
python tools/probe_norm.py 1 10240 80 128
ncu --set full --clock-control none --import-source on -k regex:norm_bwd -s 6 -c 2 -f -o gpurun_out/s8_norm_bwd python tools/probe_norm.py 128 80 80 64 > gpurun_out/s8_ncu_norm.log 2>&1
tail -2 gpurun_out/s8_ncu_norm.log
