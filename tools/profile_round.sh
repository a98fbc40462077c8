#!/bin/bash
# Round-end measurements on ONE B200 (run under gpurun): plain bench line, ncu launch list of the same step, per-launch DRAM
# traffic of one HeteroConv layer, one `ncu --set full` capture of the layer kernels and of the exact-fp32 gather / BatchNorm
# reduction kernels.  Numbers printed under ncu are never bench values.  Outputs go to gpurun_out/ (copy what is to be judged
# into profiles/).
set -u
TAG=${1:-r2_final}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_c4s8_1gpu.json 2> gpurun_out/${TAG}_bench_c4s8_1gpu.err || exit 1
python tools/layer_bench.py --workload C4s8 --reps 7 --out gpurun_out/${TAG}_layer_bench_c4s8.json > gpurun_out/${TAG}_layer_bench.log 2>&1 || exit 1
python tools/gather_bench.py --out gpurun_out/${TAG}_gather_bench_c4s8.json > gpurun_out/${TAG}_gather_bench.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/${TAG}_launches_c4s8.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_layer_traffic_c4s8.csv python tools/layer_bench.py --workload C4s8 --reps 1 --only-layer > gpurun_out/${TAG}_ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_layer_tf32|k_adjT_tf32" -c 8 -f -o gpurun_out/${TAG}_ncu_layer_kernels \
    python tools/layer_bench.py --workload C4s8 --reps 1 --only-layer > gpurun_out/${TAG}_ncu_full_layer.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gather_reduce|k_col_reduce|k_bn_apply" -c 12 -f \
    -o gpurun_out/${TAG}_ncu_gather_bn_kernels python tools/gather_bench.py --reps 1 > gpurun_out/${TAG}_ncu_full_gather.log 2>&1
# condense the reports on the box: gpurun copies gpurun_out/ back only while it stays under 64 MiB, and a full-set report of a few
# dozen launches is 20-100 MB (a 101 MB report of the step's top kernels was lost that way this round)
for rep in gpurun_out/${TAG}_ncu_layer_kernels gpurun_out/${TAG}_ncu_gather_bn_kernels; do
  if [ -f ${rep}.ncu-rep ]; then
    python tools/ncu_summary.py ${rep}.ncu-rep > ${rep}_summary.csv
    if [ $(du -sm gpurun_out | cut -f1) -gt 56 ]; then rm -f ${rep}.ncu-rep; fi
  fi
done
ls -la gpurun_out | grep ${TAG}
