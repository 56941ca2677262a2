# GPU session driver (run under gpurun): parity, storm diagnostics, bench, sweep, ncu
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
UUID=$(nvidia-smi --query-gpu=uuid --format=csv,noheader | head -1)
B=tests/_build
run_storm() { # lib tag steps perstep extra-env...
  LIB=$1; TAG=$2; STEPS=$3; PER=$4; shift 4
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/lock $SB/vmem
  env "$@" CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_CORE_LIMIT_0=25 CUDA_MEM_LIMIT_0=4g \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $LIB" timeout 100 $B/storm --steps $STEPS --warmup 1 --per-step $PER --max-seconds 40 > gpurun_out/storm_$TAG.json 2> gpurun_out/storm_$TAG.err
  echo "rc=$?" >> gpurun_out/storm_$TAG.err
}
( nvidia-smi --query-gpu=utilization.gpu,clocks.sm --format=csv,noheader -lms 250 > gpurun_out/util_trace.csv & echo $! > /tmp/smi.pid )
run_storm vgpu_manager_b200/libvgpu-control.so new 5 200000 LOGGER_LEVEL=3
echo "=== ref" >> gpurun_out/util_trace.csv
run_storm oracle/_ref/libvgpu-control.so ref 5 200000 LOGGER_LEVEL=2
kill $(cat /tmp/smi.pid)
timeout 300 python -m pytest tests/test_gpu_differential.py -q --timeout 120 -k storm > gpurun_out/pytest_storm.log 2>&1
timeout 600 python bench.py --steps 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 python bench.py --steps 3 --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 400 python profiles/sweep_spill.py > gpurun_out/sweep_spill.json 2> gpurun_out/sweep_spill.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vgpu_ -c 60 --csv --log-file gpurun_out/launches_r1.csv python profiles/run_kernels.py > gpurun_out/ncu_launches.log 2>&1
ITERS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:vgpu_spill -s 1 -c 1 -o gpurun_out/prof_spill_r1 -f python profiles/run_kernels.py > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/storm_new.json; tail -3 gpurun_out/storm_new.err; cat gpurun_out/storm_ref.json; tail -1 gpurun_out/bench.log; tail -1 gpurun_out/bench_ref.log; tail -3 gpurun_out/pytest_storm.log
