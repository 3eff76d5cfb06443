ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -f -o gpurun_out/r3_attn python scripts/prof_kernels.py attn > gpurun_out/r3_ncu_attn.log 2>&1; tail -3 gpurun_out/r3_ncu_attn.log
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_configs.py tests/test_gpu_defaults.py -x -q -m gpu 2>&1 | tail -5
for c in cfg5 cfg1; do timeout 300 python bench.py --config $c --no-cpu-baseline 2>/dev/null | cut -c1-260; done
