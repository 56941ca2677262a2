mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vgpu_ -c 3000 --csv --log-file gpurun_out/ncu_bench_launches.csv python bench.py --roofline-only > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_bench.log
tail -n 3 gpurun_out/ncu_bench.log | cut -c1-400; wc -l gpurun_out/ncu_bench_launches.csv
