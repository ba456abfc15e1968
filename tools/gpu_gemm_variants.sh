# usage: bash tools/gpu_gemm_variants.sh <variant> ...   (libhsolve_<variant>.so next to the product library)
mkdir -p gpurun_out
LOG=gpurun_out/r2_gemm_variants_c64.log
: > $LOG
for v in "$@"; do
  L=$PWD/hierarchicalsolvers.jl_b200/libhsolve_$v.so
  echo "=== $v" >> $LOG
  LIBHSOLVE_CUDA=$L timeout 300 python tools/profile_run.py 2048 helmholtz 3 >> $LOG 2>&1
  LIBHSOLVE_CUDA=$L HS_PROFILE=1 timeout 300 python tools/profile_run.py 2048 helmholtz 2 >> $LOG 2>&1
  LIBHSOLVE_CUDA=$L timeout 300 python tools/profile_run.py 64^3 helmholtz 3 >> $LOG 2>&1
done
