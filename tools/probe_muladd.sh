# A/B of the mul / add / gate epilogue of conv2_kernel on the lateral 1x1 shape (128 -> 128 @80^2, batch 64)
S="128 128 1 1 80 64 10 2"
export ACT=none
python tools/conv_probe.py $S
ADD=1 python tools/conv_probe.py $S
ADD=1 MUL=1 python tools/conv_probe.py $S
ADD=1 GATE=1 python tools/conv_probe.py $S
GATE=1 python tools/conv_probe.py $S
python tools/conv_probe.py 64 80 1 1 80 64 10 2
YAD_CONV2_NPAD=0 python tools/conv_probe.py 64 80 1 1 80 64 10 2
ACT=sigmoid ADD=1 MUL=1 python tools/conv_probe.py 128 64 1 1 80 64 10 2
