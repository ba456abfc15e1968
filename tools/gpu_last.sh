mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2_quick_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_quick_tests.log
timeout 300 python tools/profile_run.py 2048 poisson 3 > gpurun_out/r2_quick_prof.log 2>&1
timeout 300 python tools/profile_run.py 1024 helmholtz 2 >> gpurun_out/r2_quick_prof.log 2>&1
HS_PROFILE=1 timeout 600 python tools/gemm_vs_cublas.py > gpurun_out/r02_gemm_vs_cublas.txt 2> gpurun_out/r2_gvc.err
