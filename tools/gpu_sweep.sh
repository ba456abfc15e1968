# environment-switch sweep of the factorization at the 2048^2 workloads (last warm factorization of three)
mkdir -p gpurun_out
LOG=gpurun_out/r2_sweep.log
: > $LOG
run() { echo "== $*" >> $LOG; env "$@" timeout 300 python tools/profile_run.py 2048 poisson 3 2>&1 | grep "^grid" >> $LOG; }
run HS_OUTER_BLOCK=256
run HS_OUTER_BLOCK=128
run HS_OUTER_BLOCK=192
run HS_OUTER_BLOCK=320
run HS_OUTER_BLOCK=384
run HS_OUTER_BLOCK=512
run HS_LOOKAHEAD=16
run HS_LOOKAHEAD=64
run HS_LOOKAHEAD=128
run HS_LOOKAHEAD=512
