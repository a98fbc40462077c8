python -m pytest tests/test_gpu_layer.py -x -q -m gpu 2>&1 | tail -3
python tools/layer_bench.py --workload C4s8 --reps 5 2>&1 | grep "^k_layer_tf32 \|^k_layer_tf32+\|^k_adjT\|^layer_fused" | tail -4
B2G_LAYER_DBG=8 python tools/layer_bench.py --workload C4s8 --reps 1 2>&1 | grep "CTA0 cycles" | tail -1
