# Final-build evidence of the round (no sanitizer pass: the kernels are those of profiles/r2_sanitizer_*.log).
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; tail -3 gpurun_out/r2v_tests.log
python bench.py > gpurun_out/r2v_bench_1gpu.json 2> gpurun_out/r2v_bench_1gpu.err; cut -c1-300 gpurun_out/r2v_bench_1gpu.json
python bench.py --impl reference > gpurun_out/r2v_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2v_bench_reference_arm.json
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 46 -c 8 -o gpurun_out/r2v_step python scripts/ncu_step.py 1000000 9 > gpurun_out/r2v_ncu_full.log 2>&1; tail -2 gpurun_out/r2v_ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_ncu_launch.log 2>&1; tail -1 gpurun_out/r2v_ncu_launch.log | cut -c1-200
python scripts/run_configs.py > gpurun_out/r2v_configs.md 2>&1; cat gpurun_out/r2v_configs.md
python scripts/mini_probe.py > gpurun_out/r2v_mini_probe.log 2>&1; tail -5 gpurun_out/r2v_mini_probe.log
