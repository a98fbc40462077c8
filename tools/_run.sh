python -m pytest tests/test_gpu_layer.py -x -q -m gpu 2>&1 | tail -3
python tools/layer_bench.py --workload C4s8 --reps 5 2>&1 | grep "^k_layer_tf32 \|^k_adjT\|^layer_fused" | tail -4
python tools/layer_bench.py --workload C2 --reps 5 --only-layer 2>&1 | grep "^layer_fused" | tail -1
