python tools/layer_bench.py --workload C4s8 --reps 1 --only-layer > gpurun_out/lb_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_layer_traffic_c4s8.csv python tools/layer_bench.py --workload C4s8 --reps 1 --only-layer > gpurun_out/ncu_traffic.log 2>&1
python tools/layer_bench.py --workload C4s8 --reps 1 > gpurun_out/lb_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_layer_tf32|k_adjT_tf32" -c 6 -o gpurun_out/r2_ncu_layer_kernels python tools/layer_bench.py --workload C4s8 --reps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_traffic.log gpurun_out/ncu_full.log; wc -l gpurun_out/r2_layer_traffic_c4s8.csv
