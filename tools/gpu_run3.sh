mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_hss.py -q -m gpu > gpurun_out/r2_t_hss.log 2>&1; echo "hss rc=$?" >> gpurun_out/r2_t_hss.log
for g in 256 512 1024 2048; do
  timeout 900 python tools/hss_run.py $g poisson 128 1e-5 32 > gpurun_out/r2_hssrun_$g.log 2>&1; echo "rc=$?" >> gpurun_out/r2_hssrun_$g.log
  if ! grep -q "^ok" gpurun_out/r2_hssrun_$g.log; then
    if [ $g -le 512 ]; then timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python tools/hss_run.py $g poisson 128 1e-5 32 > gpurun_out/r2_san_$g.log 2>&1; fi
    break
  fi
done
timeout 600 python tools/hss_run.py 512 helmholtz 128 1e-4 32 > gpurun_out/r2_hssrun_h512.log 2>&1; echo "rc=$?" >> gpurun_out/r2_hssrun_h512.log
