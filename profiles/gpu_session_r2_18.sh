# GPU session 18 of round 2: compute-sanitizer over the control kernels added this round (refill, armed quota, slab table)
# and the slab-mode tenant (spill copy + clear launched from inside the hooks)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for tool in memcheck racecheck synccheck; do
  VSLAB_OPS=60 timeout 400 compute-sanitizer --tool $tool python profiles/run_control_kernels.py > gpurun_out/sanitizer_${tool}_control_r2.log 2>&1
  tail -3 gpurun_out/sanitizer_${tool}_control_r2.log
done
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke_r2.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke_r2.log
