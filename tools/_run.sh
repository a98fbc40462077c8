for v in "B2G_LAYER_DBG=0" "B2G_LAYER_DBG=64"; do echo "== $v"; env $v python tools/layer_bench.py --workload C4s8 --reps 5 2>&1 | grep "^k_layer_tf32 " | tail -1 | cut -c1-60; done
