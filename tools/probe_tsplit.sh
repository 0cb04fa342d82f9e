for ts in 1 3; do
export YAD_CONV2_TSPLIT=$ts
echo "TSPLIT=$ts"
ACT=silu python tools/conv_probe.py 128 128 1 1 80 64 10 0
ACT=none ADD=1 python tools/conv_probe.py 128 128 1 1 80 64 10 0
ACT=none ADD=1 GATE=1 python tools/conv_probe.py 128 128 1 1 80 64 10 0
python tools/conv_probe.py 192 128 1 1 80 64 10 0
python tools/conv_probe.py 96 128 1 1 80 64 10 0
python tools/conv_probe.py 128 128 1 1 40 64 10 0
python tools/conv_probe.py 64 80 1 1 80 64 10 0
done
unset YAD_CONV2_TSPLIT
python tools/conv_probe.py 64 64 1 1 80 64 10 0
python tools/conv_probe.py 32 32 1 1 160 64 10 0
ADD=1 python tools/conv_probe.py 8 16 3 1 160 64 10 0
