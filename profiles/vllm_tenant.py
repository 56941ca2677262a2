"""Experiment (not part of the test suite): vLLM 0.22 serving a tiny random-weight Llama as the tenant -
CUDA graphs, its own allocator profile run, FlashAttention/FlashInfer kernels - with no library and
under the B200 library (16 GiB / 50 % cap).  Greedy decoding of fixed token prompts must give the
same tokens.  Usage: python profiles/vllm_tenant.py  ->  gpurun_out/vllm_tenant_r1.json"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

TENANT = r'''
import json, os, sys, time
t0 = time.time()
from transformers import LlamaConfig
d = sys.argv[1]
LlamaConfig(vocab_size=4096, hidden_size=512, intermediate_size=1408, num_hidden_layers=4, num_attention_heads=8,
            num_key_value_heads=8, max_position_embeddings=512, torch_dtype="bfloat16").save_pretrained(d)
from vllm import LLM, SamplingParams
llm = LLM(model=d, load_format="dummy", skip_tokenizer_init=True, max_model_len=256, gpu_memory_utilization=0.25,
          seed=0, compilation_config={"mode": 0, "cudagraph_capture_sizes": [1, 2, 4, 8]})
prompts = [{"prompt_token_ids": [(7 * i + 3 * j) % 4096 for j in range(16)]} for i in range(8)]
sp = SamplingParams(temperature=0.0, max_tokens=24, detokenize=False)
outs = llm.generate(prompts, sp)
toks = [list(o.outputs[0].token_ids) for o in outs]
import torch
print("RESULT " + json.dumps({"tokens": toks, "total": torch.cuda.mem_get_info()[1], "seconds": round(time.time() - t0, 1)}))
'''


def run(lib):
    import helpers as H
    from test_gpu_framework import gpu0_uuid
    sb = H.Sandbox()
    knobs = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "CUDA_VISIBLE_DEVICES": "0",
             "CUDA_MEM_LIMIT_0": "16g", "CUDA_CORE_LIMIT_0": "50", "LOGGER_LEVEL": "2"}
    env = H.preload_env(lib, sb, knobs, stub=False) if lib else dict(os.environ, CUDA_VISIBLE_DEVICES="0")
    env.update({"VLLM_ENABLE_V1_MULTIPROCESSING": "0", "VLLM_LOGGING_LEVEL": "WARNING", "TOKENIZERS_PARALLELISM": "false"})
    d = tempfile.mkdtemp(prefix="tinyllama_")
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-c", TENANT, d], env=env, capture_output=True, text=True, timeout=420)
    except subprocess.TimeoutExpired as e:
        sb.cleanup()
        return {"error": "timeout", "stderr": (e.stderr or b"")[-1500:].decode("utf-8", "replace") if isinstance(e.stderr, bytes) else str(e.stderr)[-1500:]}
    sb.cleanup()
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    if r.returncode != 0 or not line:
        return {"error": "rc=%d" % r.returncode, "stderr": r.stderr[-2500:], "wall_s": round(time.time() - t0, 1)}
    out = json.loads(line[-1][7:])
    out["wall_s"] = round(time.time() - t0, 1)
    out["vgpu_log"] = [l for l in r.stderr.splitlines() if "vGPU" in l][-8:]
    return out


if __name__ == "__main__":
    import helpers as H
    H.build_all()
    report = {"bare": run(None)}
    if "tokens" in report["bare"]:
        report["b200"] = run(H.NEW_SO)
        report["same_tokens"] = report["b200"].get("tokens") == report["bare"]["tokens"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "vllm_tenant_r1.json"), "w") as f:
        json.dump(report, f, indent=1)
    brief = {k: ({kk: vv for kk, vv in v.items() if kk != "tokens"} if isinstance(v, dict) else v) for k, v in report.items()}
    print(json.dumps(brief, indent=1)[:4000])
