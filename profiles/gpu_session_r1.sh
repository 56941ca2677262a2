# GPU session driver (run under gpurun): storm diagnostics vs reference, tests, bench, ncu
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
UUID=$(nvidia-smi --query-gpu=uuid --format=csv,noheader | head -1)
B=tests/_build
run_storm() { # lib tag steps perstep maxsec extra-env...
  LIB=$1; TAG=$2; STEPS=$3; PER=$4; MAXS=$5; shift 5
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/lock $SB/vmem
  echo "=== $TAG" >> gpurun_out/util_trace.csv
  env "$@" CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_MEM_LIMIT_0=4g \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $LIB" timeout 120 $B/storm --steps $STEPS --warmup 1 --per-step $PER --max-seconds $MAXS > gpurun_out/storm_$TAG.json 2> gpurun_out/storm_$TAG.err
  echo "rc=$?" >> gpurun_out/storm_$TAG.err
}
rm -f gpurun_out/util_trace.csv
( nvidia-smi --query-gpu=utilization.gpu --format=csv,noheader -lms 200 >> gpurun_out/util_trace.csv & echo $! > /tmp/smi.pid )
NEW=vgpu_manager_b200/libvgpu-control.so; REF=oracle/_ref/libvgpu-control.so
run_storm $NEW new_nolimit 3 200000 30 LOGGER_LEVEL=1
run_storm $REF ref_nolimit 3 200000 30 LOGGER_LEVEL=1
run_storm $NEW new_25 40 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=25
run_storm $REF ref_25 40 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=25
run_storm $NEW new_50 40 200000 10 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=50
run_storm $REF ref_50 40 200000 10 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=50
kill $(cat /tmp/smi.pid)
timeout 700 python -m pytest tests -m gpu -q --timeout 250 > gpurun_out/pytest_gpu.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vgpu_ -s 9 -c 40 --csv --log-file gpurun_out/launches_r1.csv python profiles/run_kernels.py > gpurun_out/ncu_launches.log 2>&1
ITERS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:vgpu_spill -s 1 -c 1 -o gpurun_out/prof_spill_r1 -f python profiles/run_kernels.py > gpurun_out/ncu_full.log 2>&1
for t in new_nolimit ref_nolimit new_25 ref_25 new_50 ref_50; do echo $t; cat gpurun_out/storm_$t.json | cut -c1-700; done; tail -1 gpurun_out/bench.log | cut -c1-1500; tail -1 gpurun_out/bench_ref.log | cut -c1-600; tail -3 gpurun_out/pytest_gpu.log
