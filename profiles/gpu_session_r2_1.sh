# GPU session 1 of round 2 (run under gpurun): VMM feasibility probe, parity of the new kernels, the
# reference's tolerance band + the B200 library through the same shapes, bench both arms.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nproc > gpurun_out/nproc.txt
gcc -O2 -I/usr/local/cuda/include -o /tmp/vmm_probe profiles/probes/vmm_probe.c -ldl && /tmp/vmm_probe 64 > gpurun_out/vmm_probe_r2.json 2> gpurun_out/vmm_probe_r2.err
/tmp/vmm_probe 2 >> gpurun_out/vmm_probe_r2.json 2>> gpurun_out/vmm_probe_r2.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_differential.py -m gpu -q -x --timeout 250 > gpurun_out/pytest_gpu_s1.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1500 python tests/band.py --impl reference --runs 5 --out gpurun_out/tolerance_band.json > gpurun_out/band_ref.log 2> gpurun_out/band_ref.err
timeout 900 python tests/band.py --impl b200 --runs 2 --out gpurun_out/band_b200_s1.json > gpurun_out/band_b200.log 2> gpurun_out/band_b200.err
timeout 300 python bench.py --steps 10 --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --steps 10 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"vgpu_refill|vgpu_quota" -c 12 -o gpurun_out/ncu_control_r2 python profiles/run_control_kernels.py > gpurun_out/ncu_control.log 2>&1
tail -3 gpurun_out/pytest_gpu_s1.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/vmm_probe_r2.json; tail -1 gpurun_out/band_ref.log; tail -1 gpurun_out/band_b200.log; tail -1 gpurun_out/bench_ref.log | cut -c1-600; tail -1 gpurun_out/bench.log | cut -c1-1500
