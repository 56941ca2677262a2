# GPU session driver (run under gpurun): limiter behaviour vs reference, tests, bench (both arms)
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
run_storm $NEW new_25 100 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=25
run_storm $NEW new_25_avg1 100 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=25 VGPU_B200_UTIL_MODE=average VGPU_B200_UTIL_WINDOW_PERIODS=1
run_storm $REF ref_25 100 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=25
run_storm $NEW new_10 100 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=10
run_storm $REF ref_10 100 200000 20 LOGGER_LEVEL=1 CUDA_CORE_LIMIT_0=10
kill $(cat /tmp/smi.pid)
run_alloc() { # lib tag vmem
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/lock $SB/vmem
  env CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_MEM_LIMIT_0=4g VMEMORY_NODE_ENABLED=$3 LOGGER_LEVEL=0 \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $1" timeout 120 $B/allocstorm --n 3000 > gpurun_out/alloc_$2.json 2> gpurun_out/alloc_$2.err
}
timeout 60 $B/allocstorm --n 3000 > gpurun_out/alloc_bare.json 2>/dev/null
run_alloc $NEW new false; run_alloc $REF ref false; run_alloc $NEW new_vmem true; run_alloc $REF ref_vmem true
timeout 700 python -m pytest tests -m gpu -q --timeout 250 > gpurun_out/pytest_gpu.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --steps 5 --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --steps 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
# launch list of bench.py's device kernels (--roofline-only: the storm tenant runs with the interposer preloaded
# and cannot also run under ncu's).  Never a bench value: numbers under ncu are discarded.
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vgpu_ -c 3000 --csv \
  --log-file gpurun_out/ncu_bench_launches.csv python bench.py --roofline-only > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_bench.log
for t in new_25 new_25_avg1 ref_25 new_10 ref_10; do echo $t; cat gpurun_out/storm_$t.json | cut -c1-640; done; tail -1 gpurun_out/bench.log | cut -c1-700; tail -1 gpurun_out/bench_ref.log | cut -c1-300; tail -3 gpurun_out/pytest_gpu.log; for t in bare new ref new_vmem ref_vmem; do echo alloc_$t; cat gpurun_out/alloc_$t.json; done
