# size x speed of the match finder's table sizes (variant builds -DMTS_HL_BITS / -DMTS_HS_BITS), AP and LFP
for v in 12_11:2 12_12:2 default:2 13_13:2 14_13:1; do
  lib=${v%%:*}; ctas=${v##*:}
  if [ $lib = default ]; then L=mtscomp_b200/_build/libmtscomp_b200.so; else L=mtscomp_b200/_build/lib_$lib.so; fi
  echo "== tables $lib, CTAs per SM $ctas"
  MTSCOMP_B200_LIB=$L timeout 300 python tools/quick_bench.py 64 8 lz_ctas_per_sm=$ctas 2>&1 | grep "compress \|ratio" | tail -2 | cut -c1-200
  MTSCOMP_B200_LIB=$L MTSCOMP_B200_PARAMS=lz_ctas_per_sm=$ctas timeout 300 python tools/lfp_probe.py 2>&1 | grep "compress" | cut -c1-200
done
