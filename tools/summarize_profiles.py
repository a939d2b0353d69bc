#!/usr/bin/env python
"""Turn the .ncu-rep files of tools/profile_round1.sh (in gpurun_out/) into the committed summaries under profiles/:
r1_ncu_summary.md (per-kernel table + stall reasons), traffic.json (DRAM bytes per launch, read by bench.py for
roofline.traffic) and the launch list of the bench command."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "L1 pipe: LSU shared %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "L1 pipe: tensor operands %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("smsp__inst_executed.sum", "warp instr")]
MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def section(f, title, rep, names):
    if not os.path.exists(rep):
        print("missing report, section skipped:", rep)
        return {}
    r = raw(rep)
    h, u = r[0], r[1]
    f.write(f"## {title}\n\n| kernel | " + " | ".join(s for _, s in WANT) + " |\n|" + "---|" * (len(WANT) + 1) + "\n")
    traffic, stalls = {}, []
    seen = {}
    for row in r[2:]:
        name = row[h.index("Kernel Name")].replace("void ", "").replace("stdadk::", "")
        cells = []
        for w, _ in WANT:
            if w not in h:
                cells.append("-")
                continue
            v, un = row[h.index(w)], u[h.index(w)]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {un}".strip())
        f.write("| `" + name[:46] + "` | " + " | ".join(cells) + " |\n")
        rd = float(row[h.index("dram__bytes_read.sum")]) * MUL[u[h.index("dram__bytes_read.sum")]]
        wr = float(row[h.index("dram__bytes_write.sum")]) * MUL[u[h.index("dram__bytes_write.sum")]]
        k = seen.get(name, 0)
        seen[name] = k + 1
        key = names(name, k) if names else None
        if key:
            traffic[key] = rd + wr
        st = {h[i]: float(row[i]) for i in range(len(h)) if "pcsamp_warps_issue_stalled" in h[i]
              and "not_issued" not in h[i] and row[i] not in ("", "nan")}
        tot = sum(st.values()) or 1
        top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
        stalls.append("* `" + name[:46] + "`: " + ", ".join(
            f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for k, v in top))
    f.write("\nWarp-state samples (share per kernel):\n\n" + "\n".join(stalls) + "\n\n")
    return traffic


def train_names(name, k):
    if name.startswith("layer_fwd_kernel<1"):
        return "layer_fwd[0]"
    if name.startswith("layer_fwd_kernel<0"):
        return f"layer_fwd[{1 + k % 2}]"
    if name.startswith("layer_bwd_kernel<1"):
        return "layer_bwd[0]"
    if name.startswith("layer_bwd_kernel<0") and name.rstrip(">(BwdK)").endswith("1"):
        return "layer_bwd[2]"
    if name.startswith("layer_bwd_kernel<0"):
        return f"layer_bwd[{1 - k % 2}]"      # per step: block 2 (head), then 1, then 0 (x saved: image variant)
    if name.startswith("adamw"):
        return "adamw_ema_step"
    if name.startswith("sqnorm"):
        return "grad_sqnorm"
    return None


def main_r1():
    src = os.path.join(OUT, "launches_r1_final.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, "r1_launches_bench_nograph.csv"))
    traffic = {"source": "ncu --set full --clock-control none (tools/profile_round1.sh): dram__bytes_read.sum + "
                         "dram__bytes_write.sum per launch; training kernels from `bench.py --steps 5 --warmup 3 "
                         "--no-graph`, prediction kernels from `tools/prof_predict.py` (1M grid points); summary in "
                         "profiles/r1_ncu_summary.md"}
    with open(os.path.join(PROF, "r1_ncu_summary.md"), "w") as f:
        f.write("# Round-1 ncu summaries (B200, `ncu --set full --clock-control none --import-source on`)\n\n"
                "Produced by `tools/profile_round1.sh` (each command first run plain, exit 0) and "
                "`tools/summarize_profiles.py`; the 10-25 MB reports stay in `gpurun_out/`.  Launch list of the bench "
                "command: `profiles/r1_launches_bench_nograph.csv` (cold-cache, serialised per-launch times: compare "
                "shares, not absolutes).  Kernel replay inside CUDA-graph stream capture fails with `LaunchFailed` "
                "(`profiles/r1_launches_bench_graph_capture_fails.csv`), hence `--no-graph` for profiling; the benchmark "
                "itself replays graphs.  Template arguments: `layer_fwd_kernel<BASIS, CG, NS, CL>`, "
                "`layer_bwd_kernel<BASIS, CG, NS, LN, HEAD>`, `wgrad_kernel<BASIS, CG>`, `predict_fused_kernel<TRAIN>`.\n\n")
        t = section(f, "training step, batch 4096 = 32 tiles: one wave, latency-bound",
                    os.path.join(OUT, "prof_train_r1_final.ncu-rep"), train_names)
        traffic["train"] = t
        p = section(f, "dense prediction, 1M grid points: whole-network kernel (stdadk_predict)",
                    os.path.join(OUT, "prof_fused_r1_final.ncu-rep"), lambda n, k: "predict_fused")
        q = section(f, "dense prediction, 1M grid points: one layer_fwd per block (the path training uses)",
                    os.path.join(OUT, "prof_layered_r1_final.ncu-rep"), lambda n, k: f"layer_fwd[{k if n.startswith('layer_fwd_kernel<1') else 1 + k}]")
        traffic["predict"] = {**p, **q}
        f.write(NOTES)
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, "r1_ncu_summary.md")).read())


def tensor_pct(rep):
    r = raw(rep)
    h = r[0]
    i = h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    return float(r[2][i])


def main_r2():
    """profiles/r2_*: the round-2 captures of tools/profile_round2.sh."""
    src = os.path.join(OUT, "launches_r2.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, "r2_launches_bench_nograph.csv"))
    traffic = {"source": "ncu --set full --clock-control none (tools/profile_round2.sh): dram__bytes_read.sum + "
                         "dram__bytes_write.sum per launch; training kernels from `bench.py --steps 5 --warmup 3 "
                         "--no-graph`, block 1 alone from `tools/prof_block1.py` (1M explicit random points), field kernel "
                         "from `tools/prof_field.py` (10M-point grid), whole-network kernel from `tools/prof_predict.py` "
                         "(1M grid points); summary in profiles/r2_ncu_summary.md"}
    with open(os.path.join(PROF, "r2_ncu_summary.md"), "w") as f:
        f.write("# Round-2 ncu summaries (B200, `ncu --set full --clock-control none --import-source on`)\n\n"
                "Produced by `tools/profile_round2.sh` (each command first run plain, exit 0) and "
                "`tools/summarize_profiles.py r2`; the reports stay in `gpurun_out/`.  Launch list of the bench command: "
                "`profiles/r2_launches_bench_nograph.csv` (cold-cache, serialised per-launch times: compare shares, not "
                "absolutes).  Template arguments: `layer_fwd_kernel<BASIS, CG, NS, CL>`, `layer_bwd_kernel<BASIS, CG, NS, "
                "LN, HEAD>`, `wgrad_kernel<BASIS, CG>`, `adamw_ema_kernel<FUSED>`.\n\n")
        traffic["train"] = section(f, "training step, batch 4096 = 32 tiles: one wave, latency-bound",
                                   os.path.join(OUT, "prof_train_r2.ncu-rep"), train_names)
        b1 = section(f, "fused basis + Linear1 + LayerNorm/ReLU forward alone, 1M explicit random points (bench.py `roofline`)",
                     os.path.join(OUT, "prof_block1_r2.ncu-rep"), lambda n, k: "layer_fwd[0]")
        fd = section(f, "space-time field kernel, 10M-point grid (bench.py `roofline_gemm`, BASELINE config 3)",
                     os.path.join(OUT, "prof_field_r2.ncu-rep"), lambda n, k: "predict_field")
        fu = section(f, "whole-network kernel for explicit points, 1M grid points (stdadk_predict)",
                     os.path.join(OUT, "prof_fused_r2.ncu-rep"), lambda n, k: "predict_fused")
        sp = section(f, "support-walking block 1 at BASELINE config 4's size (K_s = 99,812, W1 102 MB, batch 65,536): forward gather and wgrad scatter",
                     os.path.join(OUT, "prof_sparse_r2.ncu-rep"),
                     lambda n, k: "sparse_fwd" if "fwd" in n else "sparse_wgrad")
        traffic["predict"] = {**b1, **fd, **fu}
        traffic["config4"] = sp
        traffic["tensor_pipe_pct"] = {k: tensor_pct(os.path.join(OUT, r)) for k, r in
                                      (("predict_field_kernel", "prof_field_r2.ncu-rep"),
                                       ("predict_fused_kernel", "prof_fused_r2.ncu-rep")) if os.path.exists(os.path.join(OUT, r))}
        f.write(NOTES_R2)
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, "r2_ncu_summary.md")).read())


def main():
    return main_r2() if sys.argv[1:] == ["r2"] else main_r1()


NOTES_R2 = """## Reading

* Training kernels run ONE wave of 32 CTAs on 148 SMs: the time is the serial latency of one tile; per-kernel times
  inside the replayed graph come from CUPTI (`bench.py: kernel_us_per_step`).
* Block 1 alone: 388 warp instructions per row (round 1: 639) at ~46 % issue utilisation; DRAM write = the h1 image
  (1.02 GB), read = weights and points: traffic / algorithmic = 1.0.  No pipe above 40 %: the kernel is bound by per-warp
  latency with 4-5 warps per scheduler (DESIGN.md section 4 has the phase timers).
* Field kernel: DRAM traffic is y_hat plus weights; tensor pipe ~23 %: the worker warps' epilogues (LayerNorm over
  TMEM loads, TF32 rounding, tcgen05.st of the next operand) take ~2.6x the MMA time per (tile, time step).
* Support walk at config 4's size: the forward gathers ~59 knot rows of 1 KB per point (3.9 GB per 65,536-point launch)
  out of the 102 MB W1, ~80 % of it from L2 (DRAM read 0.81 GB; L2 43 %, DRAM 26 % of peak; 65 % of the warp samples wait
  on those loads); the wgrad scatters the same rows back with 16-byte vector atomics (L2 63 %, DRAM write 1.15 GB,
  `mio_throttle` 25 %: the atomics queue).
* SASS evidence of the instruction mix: `profiles/r2_sass_mnemonics.txt`, checked by
  `tests/test_abi.py::test_sass_is_blackwell_native`.
"""

NOTES = """## Reading

* Training kernels run ONE wave of 32 CTAs on 148 SMs: tensor pipe and DRAM are a few per cent busy, the time is the
  serial latency of one tile (operand generation, MMA chain, two-pass epilogue) plus launch and prologue.
* `predict_fused_kernel<0>`: DRAM traffic is the 4 MB of y_hat plus weights -- the 2 GB/launch of activation images
  of the layered path is gone.  What limits it is the L1/shared-memory data pipe: tensor-core operand fetches
  (A 16 KB + B 32 KB per 32-wide K slab) and the weight-slab fills share it with the worker warps' LDS/STS
  (parameters and knots are warp-broadcast loads, 4 bytes per wavefront); the pipe columns above add up to the
  kernel's real ceiling (DESIGN.md section 4).
* Layered prediction kernels: block 2 moves exactly the algorithmic bytes (1.02 GB in + 0.98 GB out).
* SASS evidence of the instruction mix (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = TMA bulk copies, SYNCS =
  mbarrier, FFMA2/FADD2 = packed FP32; no HMMA): `profiles/r1_sass_mnemonics.txt`, checked by
  `tests/test_abi.py::test_sass_is_blackwell_native`.
"""

if __name__ == "__main__":
    sys.exit(main())
