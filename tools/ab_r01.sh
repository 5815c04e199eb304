# alternating A/B of the SN kernel builds (one process each; N = 1e7 as in bench.py)
for rep in 1 2 3; do
  for lib in variants/v8.so cosmopmc_b200/libpmc_b200.so variants/v9_rep1.so; do
    nc=0; [ $lib = variants/v9_rep1.so ] && nc=1
    echo -n "rep$rep NO_CNODE=$nc "; PMCB200_SN_NO_CNODE=$nc PMCB200_LIB=$PWD/$lib timeout 100 python tools/time_sn.py --n 10000000 2>&1 | tail -1
    nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
  done
done
