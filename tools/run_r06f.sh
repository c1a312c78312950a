#!/bin/bash
# same-box A/B of the stored-stencil tile kernel variants (libraries prebuilt under voxelfem_b200/csrc/build_ab)
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r06f}; O=gpurun_out; mkdir -p $O
cp voxelfem_b200/libvoxelfem_b200.so /tmp/lib_keep.so
export VF_ST_EARLY_TMA=0
for lib in new SERIAL_TAIL NO_SHUFFLE IDX64 SERIAL_TAIL_DVF_ST_NO_SHUFFLE_DVF_ST_IDX64; do
  cp voxelfem_b200/csrc/build_ab/lib_$lib.so voxelfem_b200/libvoxelfem_b200.so
  for pt in 0 1; do
    echo "== $lib postab=$pt" | tee -a $O/${T}_ab.log
    VF_ST_POSTAB=$pt timeout 200 python tools/time_ops.py 2>&1 | grep -E "level 1 (smooth|residual|apply)|level 2 (smooth|residual)|vcycle from level [12]|FMG" | tee -a $O/${T}_ab.log
  done
done
cp /tmp/lib_keep.so voxelfem_b200/libvoxelfem_b200.so
