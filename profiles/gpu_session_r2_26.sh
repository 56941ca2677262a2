mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_smwatcher.py -m gpu -q -k on_device > gpurun_out/s26_smwatcher.txt 2>&1; tail -4 gpurun_out/s26_smwatcher.txt
timeout 40 python - > gpurun_out/s26_reset_uva.txt 2>&1 <<'PY'
import sys, subprocess
sys.path.insert(0, "tests")
import helpers as H
uuid = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True).stdout.splitlines()[0].strip()
env = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": uuid, "CUDA_VISIBLE_DEVICES": "0", "LOGGER_LEVEL": "3",
       "CUDA_MEM_LIMIT_0": "8g", "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true"}
script = "init 0\nalloc 1610612736\nalloc 1073741824\nalloc 4096\nledger 0\nreset\nledger 0\nalloc 4096\nalloc 8192\nfree 0\nfree 1\nledger 0\nnvmlinfo\n"
for lib in (H.REF_SO, H.NEW_SO):
    out, err, rc = H.run_scenario(lib, script, env, stub=False, check=False, timeout=30)
    print(lib.split("/")[-2], "rc", rc); print(out); print(err[-1200:])
PY
tail -30 gpurun_out/s26_reset_uva.txt
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k random_reference > gpurun_out/s26_parity_random.txt 2>&1; tail -4 gpurun_out/s26_parity_random.txt
