#!/bin/bash
# Round 1, extra measurements: config-3 GEMM tenants, 8-thread storms, external SM-watcher mode
# end to end (vgpu-smwatcher producing sm_util.config from the real NVML for both libraries).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python __graft_entry__.py > gpurun_out/extra_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_config3_gemm.py -x -q -m gpu > gpurun_out/extra_config3.log 2>&1
echo "config3 rc=$?" > gpurun_out/extra_status.txt
UUID=$(nvidia-smi --query-gpu=uuid --format=csv,noheader | head -1)
B=tests/_build
NEW=vgpu_manager_b200/libvgpu-control.so; REF=oracle/_ref/libvgpu-control.so
run_storm() { # lib tag threads maxsec extra-env...
  LIB=$1; TAG=$2; THR=$3; MAXS=$4; shift 4
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/etc/vgpu-manager/watcher $SB/lock $SB/vmem
  if [ -n "$WATCH" ]; then
    vgpu_manager_b200/vgpu-smwatcher --file $SB/etc/vgpu-manager/watcher/sm_util.config --verbose 2> gpurun_out/extra_watcher_$TAG.err &
    WPID=$!
    sleep 0.5
  fi
  env "$@" CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_MEM_LIMIT_0=4g LOGGER_LEVEL=1 \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $LIB" timeout 120 $B/storm --steps 100 --warmup 1 --per-step 200000 --threads $THR --max-seconds $MAXS \
    > gpurun_out/extra_storm_$TAG.json 2> gpurun_out/extra_storm_$TAG.err
  echo "rc=$?" >> gpurun_out/extra_storm_$TAG.err
  if [ -n "$WATCH" ]; then
    python - $SB/etc/vgpu-manager/watcher/sm_util.config > gpurun_out/extra_watcher_$TAG.txt <<'PY'
import struct, sys
raw = open(sys.argv[1], "rb").read()
ns, = struct.unpack_from("<I", raw, 32768); last, = struct.unpack_from("<Q", raw, 32776)
nc, = struct.unpack_from("<I", raw, 57360); ng, = struct.unpack_from("<I", raw, 81944)
print("file bytes", len(raw), "samples", ns, "compute", nc, "graphics", ng, "last_seen_us", last)
for i in range(min(ns, 4)):
    print(" sample", struct.unpack_from("<IxxxxQIIII", raw, 32 * i))
for i in range(min(nc, 4)):
    print(" compute", struct.unpack_from("<IxxxxQII", raw, 32784 + 24 * i))
PY
    kill $WPID; wait $WPID 2>/dev/null
  fi
}
WATCH=
run_storm $NEW new_8thr_nocap 8 8
run_storm $REF ref_8thr_nocap 8 8
run_storm $NEW new_8thr_25 8 12 CUDA_CORE_LIMIT_0=25
run_storm $REF ref_8thr_25 8 12 CUDA_CORE_LIMIT_0=25
WATCH=1
run_storm $NEW new_extwatch_25 1 12 CUDA_CORE_LIMIT_0=25 EXTERNAL_SM_WATCHER_ENABLED=true
run_storm $REF ref_extwatch_25 1 12 CUDA_CORE_LIMIT_0=25 EXTERNAL_SM_WATCHER_ENABLED=true
for t in new_8thr_nocap ref_8thr_nocap new_8thr_25 ref_8thr_25 new_extwatch_25 ref_extwatch_25; do echo $t; cut -c1-420 gpurun_out/extra_storm_$t.json; tail -n 2 gpurun_out/extra_storm_$t.err; done
cat gpurun_out/extra_watcher_*.txt
tail -n 5 gpurun_out/extra_config3.log; cat gpurun_out/extra_status.txt
