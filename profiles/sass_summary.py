"""SASS evidence for the device image (no GPU needed): cuobjdump -sass of kernels.fatbin, split per
kernel; full listings of the two bandwidth kernels, mnemonic counts for all.
    python profiles/sass_summary.py   -> profiles/sass_r2/{summary.json, <kernel>.sass for the bandwidth, refill, quota, sampler and slab-placement kernels}
"""
import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FATBIN = os.path.join(ROOT, "vgpu_manager_b200", "csrc", "kernels.fatbin")
OUT = os.path.join(ROOT, "profiles", "sass_r2")
FULL = ("vgpu_spill_copy_kernel", "vgpu_clear_kernel", "vgpu_refill_kernel", "vgpu_quota_kernel", "vgpu_sampler_kernel", "vgpu_vslab_kernel")
WATCH = ("UBLKCP", "SYNCS", "STG.E.128", "STG.E.EF.128", "LDG.E.128", "LDG.E.CONSTANT.128", "LDG.E.128.CONSTANT", "UTMALDG", "NANOSLEEP",
         "ATOMG", "RED", "BAR.SYNC", "SHFL", "MEMBAR", "CS2R", "S2UR", "LDS", "STS")


def main():
    text = subprocess.run(["cuobjdump", "-sass", FATBIN], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", text)))
    parts = re.split(r"\n\s*Function : ", text)
    os.makedirs(OUT, exist_ok=True)
    summary = {"fatbin": os.path.relpath(FATBIN, ROOT), "arch": arch, "kernels": {}}
    for part in parts[1:]:
        name, body = part.split("\n", 1)
        name = name.strip()
        ops = collections.Counter()
        n = 0
        for line in body.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                n += 1
                ops[m.group(1)] += 1
        watched = {}
        for w in WATCH:
            c = sum(v for k, v in ops.items() if k == w or k.startswith(w + ".") or (w in ("STG.E.128", "LDG.E.128") and w.split(".")[-1] in k.split(".") and k.startswith(w.split(".")[0])))
            if c:
                watched[w] = c
        summary["kernels"][name] = {"instructions": n, "watched": watched, "top": ops.most_common(8)}
        if name in FULL:
            with open(os.path.join(OUT, name + ".sass"), "w") as f:
                f.write("Function : " + name + "\n" + body)
    with open(os.path.join(OUT, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    for k, v in summary["kernels"].items():
        print(k, v["instructions"], v["watched"])


if __name__ == "__main__":
    main()
