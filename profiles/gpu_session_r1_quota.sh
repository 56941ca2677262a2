#!/bin/bash
# quota-kernel geometry change: parity first, then the allocation storm (both libraries, same box)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_differential.py -x -q -m gpu 2>&1 | tail -n 6
UUID=$(nvidia-smi --query-gpu=uuid --format=csv,noheader | head -1)
B=tests/_build; NEW=vgpu_manager_b200/libvgpu-control.so; REF=oracle/_ref/libvgpu-control.so
run_alloc() { # lib tag vmem
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/lock $SB/vmem
  env CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_MEM_LIMIT_0=4g VMEMORY_NODE_ENABLED=$3 LOGGER_LEVEL=0 \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $1" timeout 120 $B/allocstorm --n 3000 > gpurun_out/alloc_$2.json 2> gpurun_out/alloc_$2.err
}
for i in 1 2; do
  run_alloc $NEW new$i false; run_alloc $REF ref$i false
done
run_alloc $NEW new_vmem true; run_alloc $REF ref_vmem true
for t in new1 ref1 new2 ref2 new_vmem ref_vmem; do echo alloc_$t; cat gpurun_out/alloc_$t.json; done
