mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_hss.py -x -q -m gpu -k "not children" > gpurun_out/r2_t_hss.log 2>&1; echo "hss rc=$?" >> gpurun_out/r2_t_hss.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_hss.py > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?" >> gpurun_out/r2_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err; echo "ref rc=$?" >> gpurun_out/r2_ref1.err
tail -3 gpurun_out/r2_t_hss.log gpurun_out/r2_t_all.log gpurun_out/r2_bench1.err gpurun_out/r2_ref1.err
