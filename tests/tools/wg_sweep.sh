# sweep of the weight-gradient pixel-split heuristics (S2R_WG_* knobs): one block of conv_bench wgrad lines per setting
for cfg in "2.0 8" "1.0 8" "0.5 8" "1.0 16" "0.5 32" "0.25 32"; do
  set -- $cfg
  echo "== ctas_per_sm=$1 min_chunks=$2"
  S2R_WG_CTAS_PER_SM=$1 S2R_WG_MIN_CHUNKS=$2 python tests/tools/conv_bench.py 2>&1 | grep wgrad
done
