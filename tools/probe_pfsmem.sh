for pf in 1 0; do
export YAD_CONV2_PFSMEM=$pf
echo "PFSMEM=$pf"
ACT=none ADD=1 python tools/conv_probe.py 128 128 1 1 80 64 10 0
ACT=none ADD=1 MUL=1 python tools/conv_probe.py 128 128 1 1 80 64 10 0
ACT=none ADD=1 GATE=1 python tools/conv_probe.py 128 128 1 1 80 64 10 0
ACT=sigmoid ADD=1 MUL=1 python tools/conv_probe.py 128 64 1 1 80 64 10 0
ADD=1 python tools/conv_probe.py 64 64 3 1 80 64 10 0
ADD=1 python tools/conv_probe.py 64 64 1 1 80 64 10 0
done
