mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r3_pytest4.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r3_smoke.log
timeout 500 python bench.py 2>gpurun_out/r3_bench_cfg2.err > gpurun_out/r3_bench_cfg2.json; cut -c1-300 gpurun_out/r3_bench_cfg2.json
timeout 400 python bench.py --config cfg5 --no-cpu-baseline 2>gpurun_out/r3_bench_cfg5.err > gpurun_out/r3_bench_cfg5.json; cut -c1-300 gpurun_out/r3_bench_cfg5.json
timeout 300 python bench.py --config cfg1 --no-cpu-baseline 2>gpurun_out/r3_bench_cfg1.err > gpurun_out/r3_bench_cfg1.json; cut -c1-300 gpurun_out/r3_bench_cfg1.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 2>gpurun_out/r3_bench_ref.err > gpurun_out/r3_bench_ref.json; cut -c1-400 gpurun_out/r3_bench_ref.json
