set -x
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_step.py 4000 > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r2_sanitizer_racecheck.log
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 46 -c 8 -o gpurun_out/r2_step python scripts/ncu_step.py 1000000 9 > gpurun_out/r2_ncu_full.log 2>&1; tail -2 gpurun_out/r2_ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1; tail -1 gpurun_out/r2_ncu_launch.log | cut -c1-200
python scripts/mini_probe.py > gpurun_out/r2_mini_probe.log 2>&1; cat gpurun_out/r2_mini_probe.log
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; cut -c1-400 gpurun_out/r2_bench_1gpu.json
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-400 gpurun_out/r2_bench_reference_arm.json
python scripts/run_configs.py > gpurun_out/r2_configs.md 2>&1; cat gpurun_out/r2_configs.md
