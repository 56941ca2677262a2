#!/bin/bash
# compute-sanitizer over one pass of every library kernel (the smoke() path: quota, slab via bring-up
# warm-ups, spill copy, clear, sampler tail + controller)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool rc=$? $(tail -n 1 gpurun_out/sanitizer_$tool.out | cut -c1-120)"
  tail -n 3 gpurun_out/sanitizer_$tool.log
done
