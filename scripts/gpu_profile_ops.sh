#!/bin/bash
TAG=${1:-r01}
KERN=${2:-'gram_fast|kmeans_assign|ade_fde_fast|reconstruct_fast|reconstruct_bwd_fast|eig_jacobi|svd_small|project_fast|kmeans_seed_step'}
mkdir -p gpurun_out
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:"$KERN" \
  -s ${SKIP:-20} -c ${COUNT:-40} -f -o gpurun_out/prof_ops_$TAG python scripts/exp/run_ops_once.py > gpurun_out/ncu_ops_$TAG.log 2>&1
echo "ncu ops exit $?"; tail -3 gpurun_out/ncu_ops_$TAG.log
