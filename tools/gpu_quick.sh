# quick GPU check after a kernel change: solve/factor parity tests, then phase timings of the 2048^2 workloads
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2_quick_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_quick_tests.log
: > gpurun_out/r2_quick_prof.log
timeout 300 python tools/profile_run.py 2048 poisson 3 >> gpurun_out/r2_quick_prof.log 2>&1
HS_PROFILE=1 timeout 300 python tools/profile_run.py 2048 poisson 2 >> gpurun_out/r2_quick_prof.log 2>&1
timeout 300 python tools/profile_run.py 2048 helmholtz 3 >> gpurun_out/r2_quick_prof.log 2>&1
HS_PROFILE=1 timeout 300 python tools/profile_run.py 2048 helmholtz 2 >> gpurun_out/r2_quick_prof.log 2>&1
