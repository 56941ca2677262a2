"""GPU (B200): a small transformer training loop (transformers' Llama blocks, random init, AdamW, bf16
autocast) as the tenant - attention / norm / optimizer kernels, cuBLAS workspaces, the caching
allocator growing and trimming - under both libraries and without one: same losses, same capped
memory view, and the run completes under a 30 % core cap."""
import json
import os
import subprocess
import sys

import pytest

import helpers as H
from test_gpu_framework import gpu0_uuid

pytestmark = pytest.mark.gpu

TENANT = r'''
import json, torch
from transformers import LlamaConfig, LlamaForCausalLM
torch.manual_seed(0)
cfg = LlamaConfig(vocab_size=2048, hidden_size=512, intermediate_size=1408, num_hidden_layers=4, num_attention_heads=8,
                  num_key_value_heads=8, max_position_embeddings=256)
model = LlamaForCausalLM(cfg).cuda()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
g = torch.Generator(device="cuda").manual_seed(1)
losses = []
for step in range(12):
    ids = torch.randint(0, 2048, (8, 128), device="cuda", generator=g)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model(input_ids=ids, labels=ids).loss
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    losses.append(round(float(loss), 4))
torch.cuda.empty_cache()
free, total = torch.cuda.mem_get_info()
print(json.dumps({"losses": losses, "total": total}))
'''


def run(lib):
    sb = H.Sandbox()
    knobs = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "CUDA_VISIBLE_DEVICES": "0",
             "CUDA_MEM_LIMIT_0": "10g", "CUDA_CORE_LIMIT_0": "30", "LOGGER_LEVEL": "1"}
    env = H.preload_env(lib, sb, knobs, stub=False) if lib else dict(os.environ, CUDA_VISIBLE_DEVICES="0")
    r = subprocess.run([sys.executable, "-c", TENANT], env=env, capture_output=True, text=True, timeout=400)
    sb.cleanup()
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def test_transformer_training_loop(built):
    bare = run(None)
    ours = run(H.NEW_SO)
    assert ours["losses"] == bare["losses"] and ours["total"] == 10 << 30
    assert len(set(ours["losses"])) > 1  # the optimizer did move the weights
    if os.path.exists(H.REF_SO):
        assert run(H.REF_SO) == ours
