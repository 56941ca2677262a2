# GPU session 3 of round 2: full per-step logs of both libraries for the two shapes that disagree
# (50 % storm: who sees which first reading; 4 x 25 % busy tenants: share trajectories), side by side.
mkdir -p gpurun_out/logs_r2
python __graft_entry__.py > gpurun_out/build.log 2>&1
UUID=$(nvidia-smi --query-gpu=uuid --format=csv,noheader | head -1)
B=tests/_build
tenant() { # lib tag cap seconds busy
  SB=$(mktemp -d); mkdir -p $SB/etc/vgpu-manager/config $SB/lock $SB/vmem
  EXTRA=""; [ "$5" = "busy" ] && EXTRA="--spin-iters 20000 --grid 592 --block 256"
  PER=200000; [ "$5" = "busy" ] && PER=200
  env CUDA_VISIBLE_DEVICES=0 MANAGER_COMPATIBILITY_MODE=0 MANAGER_VISIBLE_DEVICES=$UUID CUDA_MEM_LIMIT_0=4g CUDA_CORE_LIMIT_0=$3 LOGGER_LEVEL=4 \
    VGPU_REDIRECT="/etc/vgpu-manager=$SB/etc/vgpu-manager:/tmp/.vgpu_lock=$SB/lock:/tmp/.vmem_node=$SB/vmem" \
    LD_PRELOAD="$B/libredirect.so $1" timeout 120 $B/storm --steps 1000000 --warmup 0 --per-step $PER --max-seconds $4 $EXTRA \
    > gpurun_out/logs_r2/$2.json 2> gpurun_out/logs_r2/$2.err
}
NEW=vgpu_manager_b200/libvgpu-control.so; REF=oracle/_ref/libvgpu-control.so
tenant $REF storm50_ref 50 10 empty
tenant $NEW storm50_b200 50 10 empty
tenant $REF storm50_ref_b 50 10 empty
tenant $NEW storm50_b200_b 50 10 empty
for L in ref b200; do
  LIB=$REF; [ $L = b200 ] && LIB=$NEW
  for i in 0 1 2 3; do tenant $LIB fair4_${L}_$i 25 8 busy & done
  wait
done
for L in b200; do for i in 0 1 2 3; do tenant $NEW fair4b_${L}_$i 25 8 busy & done; wait; done
for f in gpurun_out/logs_r2/*.json; do echo $f; cut -c1-200 $f; done
