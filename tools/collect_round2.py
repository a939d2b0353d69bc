#!/usr/bin/env python
"""Gather the round's bench JSON lines (written by the GPU runs into gpurun_out/) and the SASS mnemonic counts of the
built library into profiles/ (r2_bench_lines.jsonl, r2_sass_mnemonics.txt)."""
import json
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
RUNS = [("b_1_final.json", "python bench.py  (1 GPU, defaults: BASELINE config 2 shape)"),
        ("b_1_ref.json", "python bench.py --impl reference  (the reference's PyTorch CPU path on the box's host cores)"),
        ("b_2b.json", "torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 50 --warmup 5"),
        ("b_4.json", "torchrun --nproc-per-node 4 bench.py --gpus 4 --steps 50 --warmup 5"),
        ("b_8c.json", "torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 50 --warmup 5"),
        ("b64k_1.json", "python bench.py --batch 65536 --steps 100 --warmup 5"),
        ("b64k_8.json", "torchrun --nproc-per-node 8 bench.py --gpus 8 --batch 65536 --steps 50 --warmup 5"),
        ("c1_1.json", "python bench.py --config 1 --steps 20"),
        ("c3_1.json", "python bench.py --config 3"),
        ("c3_8b.json", "torchrun --nproc-per-node 8 bench.py --gpus 8 --config 3"),
        ("c4_1.json", "python bench.py --config 4"),
        ("c4_2.json", "torchrun --nproc-per-node 2 bench.py --gpus 2 --config 4"),
        ("c4_4.json", "torchrun --nproc-per-node 4 bench.py --gpus 4 --config 4"),
        ("c4_8.json", "torchrun --nproc-per-node 8 bench.py --gpus 8 --config 4"),
        ("c5_1.json", "STDADK_CONFIGS_PER_GPU=4 python bench.py --config 5 --steps 50"),
        ("c5_8.json", "STDADK_CONFIGS_PER_GPU=2 torchrun --nproc-per-node 8 bench.py --gpus 8 --config 5 --steps 50  (before the workers were forked / shared the host cores; with the reference leg)"),
        ("c5_8b.json", "STDADK_SKIP_REFERENCE_LEG=1 STDADK_CONFIGS_PER_GPU=2 torchrun --nproc-per-node 8 bench.py --gpus 8 --config 5 --steps 50")]


def main():
    n = 0
    with open(os.path.join(PROF, "r2_bench_lines.jsonl"), "w") as f:
        for name, cmd in RUNS:
            p = os.path.join(OUT, name)
            if not os.path.exists(p):
                print("missing", name)
                continue
            lines = [ln for ln in open(p).read().splitlines() if ln.startswith("{")]
            if not lines:
                print("no JSON line in", name)
                continue
            f.write(json.dumps({"command": cmd, "line": json.loads(lines[-1])}) + "\n")
            n += 1
    print(n, "bench lines")
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "st_dadk_b200", "libstdadk.so")], capture_output=True,
                          text=True).stdout
    c = Counter()
    for m in re.finditer(r"\b(UTCHMMA|UTCBAR|LDTM|STTM|UBLKCP|SYNCS|FFMA2|FADD2|FMUL2|HMMA|UTCCP)\b", sass):
        c[m.group(1)] += 1
    with open(os.path.join(PROF, "r2_sass_mnemonics.txt"), "w") as f:
        for k in sorted(c):
            f.write(f"{c[k]:7d} {k}\n")
    print(dict(c))


if __name__ == "__main__":
    sys.exit(main())
