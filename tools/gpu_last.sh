mkdir -p gpurun_out
HS_PROFILE=1 timeout 600 python tools/gemm_vs_cublas.py > gpurun_out/r02_gemm_vs_cublas.txt 2> gpurun_out/r2_gvc.err
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t_full.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_t_full.log
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?" >> gpurun_out/r2_bench_final.err
