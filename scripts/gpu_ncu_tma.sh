set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tma_rows_nn -s 3 -c 1 -o gpurun_out/ncu_tma_nn_final -f python scripts/microbench.py --only "linear_fwd:32+0->32@L0" --iters 2 --warmup 3 > gpurun_out/ncu_nn.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tma_rows_tn -s 6 -c 2 -o gpurun_out/ncu_tma_tn_wide -f python scripts/microbench.py --only "linear_bwd_weight:128+32->32@L1" --iters 2 --warmup 3 > gpurun_out/ncu_tn.log 2>&1
ls -la gpurun_out/*.ncu-rep
