set -x
python scripts/policy_chunks_check.py 1000000 3 > gpurun_out/r2u_pipe.log 2>&1; cat gpurun_out/r2u_pipe.log | tail -8
python scripts/policy_chunks_check.py 300001 2 > gpurun_out/r2u_pipe_k2.log 2>&1; tail -6 gpurun_out/r2u_pipe_k2.log
python scripts/c3_probe.py > gpurun_out/r2u_c3.log 2>&1; tail -3 gpurun_out/r2u_c3.log
FGNN_AB_EXTRA="FGNN_PDL=1" python scripts/ab_variants.py 1000000 200 > gpurun_out/r2u_ab.log 2>&1; tail -4 gpurun_out/r2u_ab.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; tail -3 gpurun_out/r2u_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; cut -c1-1500 gpurun_out/r2u_bench.json
