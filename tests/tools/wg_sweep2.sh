for k in 3.0 6.0 8.0 12.0; do echo "== ctas_per_sm=$k"; S2R_WG_CTAS_PER_SM=$k python tests/tools/conv_bench.py 2>&1 | grep wgrad | grep -E "aspp 3x3|dec 3x3|D 4x4s2 (64|128|256)"; done
