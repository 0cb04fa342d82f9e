# epilogue-bound shapes of conv2_kernel (tile-split 1x1) and conv3_kernel (small-channel 3x3) at batch 64
python tools/conv_probe.py 64 64 1 1 80 64 10 0
python tools/conv_probe.py 32 32 1 1 160 64 10 0
python tools/conv_probe.py 64 32 1 1 80 64 10 0
ADD=1 python tools/conv_probe.py 8 16 3 1 160 64 10 0
python tools/conv_probe.py 16 8 3 1 160 64 10 0
ADD=1 python tools/conv_probe.py 16 32 3 1 80 64 10 0
ADD=1 python tools/conv_probe.py 64 64 3 1 80 64 10 0
ACT=sigmoid ADD=1 MUL=1 python tools/conv_probe.py 128 64 1 1 80 64 10 0
ACT=none ADD=1 GATE=1 python tools/conv_probe.py 128 128 1 1 80 64 10 0
