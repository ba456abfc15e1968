mkdir -p gpurun_out
timeout 600 python tools/hss_debug.py poisson 33 2 > gpurun_out/r2_dbg_p.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dbg_p.log
timeout 600 python tools/hss_debug.py helmholtz 33 -2 > gpurun_out/r2_dbg_h.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dbg_h.log
if ! grep -q "^ok" gpurun_out/r2_dbg_h.log; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/hss_debug.py helmholtz 33 -2 > gpurun_out/r2_san_h.log 2>&1
fi
if ! grep -q "^ok" gpurun_out/r2_dbg_p.log; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/hss_debug.py poisson 33 2 > gpurun_out/r2_san_p.log 2>&1
fi
timeout 1800 python -m pytest tests/test_gpu_hss.py -q -m gpu > gpurun_out/r2_t_hss.log 2>&1; echo "hss rc=$?" >> gpurun_out/r2_t_hss.log
timeout 900 python -m pytest tests/test_ordering.py tests/test_gpu_compress.py -q -m gpu > gpurun_out/r2_t_ord.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t_ord.log
