mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
UUID=$(nvidia-smi --query-gpu=uuid --format=csv,noheader | head -1)
B=tests/_build
run_storm() { LIB=$1; TAG=$2; STEPS=$3; PER=$4; MAXS=$5; shift 5
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/lock $SB/vmem
  env "$@" CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_MEM_LIMIT_0=4g \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $LIB" timeout 120 $B/storm --steps $STEPS --warmup 1 --per-step $PER --max-seconds $MAXS > gpurun_out/storm_$TAG.json 2> gpurun_out/storm_$TAG.err; }
NEW=vgpu_manager_b200/libvgpu-control.so; REF=oracle/_ref/libvgpu-control.so
for W in 4 6 8; do
 run_storm $NEW new_25_block$W 200 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=25 VGPU_B200_UTIL_WINDOW_PERIODS=$W
 run_storm $NEW new_10_block$W 200 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=10 VGPU_B200_UTIL_WINDOW_PERIODS=$W
done
run_storm $REF ref_25b 200 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=25
run_storm $REF ref_10b 200 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=10
timeout 300 python -m pytest tests/test_gpu_reference_suite.py tests/test_gpu_differential.py -q --timeout 250 > gpurun_out/pytest_gpu2.log 2>&1
for t in new_25_block4 new_25_block6 new_25_block8 ref_25b new_10_block4 new_10_block6 new_10_block8 ref_10b; do echo -n "$t "; python -c "import json;d=json.load(open('gpurun_out/storm_$t.json'));print(round(d['launches_per_s']), d['p50_ns'], d['p99_ns'], d['max_ns'])"; done; tail -3 gpurun_out/pytest_gpu2.log
