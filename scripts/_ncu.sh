mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r3_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/r3_attn python scripts/prof_kernels.py attn > gpurun_out/r3_ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sinkhorn_fused -s 1 -c 1 -f -o gpurun_out/r3_sinkhorn python scripts/prof_kernels.py sinkhorn > gpurun_out/r3_ncu_sk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rowcol_ -s 2 -c 2 -f -o gpurun_out/r3_rowcol python scripts/prof_kernels.py assign > gpurun_out/r3_ncu_rowcol.log 2>&1
ls -la gpurun_out/r3_*.ncu-rep gpurun_out/r3_bench_launches.csv
