mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT 2>/dev/null || true
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 200 > gpurun_out/pytest_parity.log 2>&1
UUID=$(nvidia-smi --query-gpu=uuid --format=csv,noheader | head -1)
B=tests/_build
run_storm() { # lib tag extra-env...
  LIB=$1; TAG=$2; shift 2
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/lock $SB/vmem
  env "$@" CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_CORE_LIMIT_0=25 CUDA_MEM_LIMIT_0=4g \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $LIB" timeout 90 $B/storm --steps 3 --warmup 1 --per-step 50000 --max-seconds 40 > gpurun_out/storm_$TAG.json 2> gpurun_out/storm_$TAG.err
  echo "rc=$?" >> gpurun_out/storm_$TAG.err
}
timeout 60 $B/storm --steps 3 --warmup 1 --per-step 50000 > gpurun_out/storm_bare.json 2> gpurun_out/storm_bare.err
run_storm vgpu_manager_b200/libvgpu-control.so new LOGGER_LEVEL=3
run_storm oracle/_ref/libvgpu-control.so ref LOGGER_LEVEL=2
timeout 600 python -m pytest tests/test_gpu_differential.py -q --timeout 250 > gpurun_out/pytest_diff.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 500 python bench.py --steps 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_parity.log; tail -3 gpurun_out/pytest_diff.log; cat gpurun_out/storm_new.json; tail -2 gpurun_out/bench.log
