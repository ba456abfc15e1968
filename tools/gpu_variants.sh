# usage: bash tools/gpu_variants.sh <grid> <kind> <variant> ...   — phase timings with libhsolve_<variant>.so
mkdir -p gpurun_out
LOG=gpurun_out/r2_variants.log
: > $LOG
G=$1; K=$2; shift 2
for v in "$@"; do
  L=$PWD/hierarchicalsolvers.jl_b200/libhsolve_$v.so
  echo "=== $v" >> $LOG
  LIBHSOLVE_CUDA=$L timeout 300 python tools/profile_run.py $G $K 3 >> $LOG 2>&1
  LIBHSOLVE_CUDA=$L HS_PROFILE=1 timeout 300 python tools/profile_run.py $G $K 2 >> $LOG 2>&1
done
