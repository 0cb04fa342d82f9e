# A/B of the TMA-store epilogue of conv_tma_kernel (YAD_CONV_TMA_STORE=0: staged transpose + st.global; 1: bulk tensor store)
for d in 0 1; do
  echo "== YAD_CONV_TMA_STORE=$d"
  for shape in "128 128 1 1 80" "128 128 1 1 20" "64 64 1 1 80" "256 256 1 1 20" "48 64 1 1 160" "32 32 1 1 160" "128 64 1 1 80" "512 256 1 1 20" "64 64 3 1 80"; do
    YAD_CONV_TMA_STORE=$d python tools/conv_probe.py $shape 64 20 2 2>&1 | tail -1
  done
done
