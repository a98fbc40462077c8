python -m pytest tests/test_gpu_ingest.py tests/test_gpu_e2e_parity.py::test_bulk_imputation_hidden_256_matches_oracle tests/test_gpu_parity.py::test_eval_metrics_match_reference -q -m gpu -s 2>&1 | grep -v Warning | grep "^\[\|passed\|failed\|Error\|error\|assert" | head -30
python bench.py --workload C5 --steps 5 > gpurun_out/r2_bench_c5_1gpu.json 2> gpurun_out/r2_bench_c5_1gpu.err; echo "bench C5 rc=$?"; tail -c 1500 gpurun_out/r2_bench_c5_1gpu.err | grep -v Warning | tail -5
cut -c1-900 gpurun_out/r2_bench_c5_1gpu.json
