#!/usr/bin/env python
"""Turn one `ncu --set full --import-source on` capture of a bench step into profiles/r02_counters.json: per kernel the
executed instruction mix (source page), FP32 / FP64 operations per parameter sample, DRAM bytes per sample and the pipe /
issue utilisation (raw page).  bench.py reads the file to report the EXECUTED-flop roofline fraction and the DRAM traffic of
the build it is timing; the file carries a hash of the kernel sources and bench.py ignores it when the hash differs.

Usage (after `gpurun -- 'ncu --set full --clock-control none --import-source on -k regex:"^k_|^kw_" -c 24 -o gpurun_out/r02_step
       python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary --samples 4096'`):
    python tools/ncu_counters.py gpurun_out/r02_step.ncu-rep --samples 4096 --trials 200 --T 1200 --dims 2 3 1 2 2
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# ncu function name -> the kernel kinds of lqgk_profile_read / bench.py ("k_" + kind)
KIND = {"k_pack": "k_pack", "k_repack_obs": "k_pack", "k_lqr_fwd": "k_lqr_fwd", "k_kf_fwd": "k_kf_fwd", "k_cov_fwd": "k_cov_fwd",
        "kw_cov_fwd": "k_cov_fwd", "k_trial_fwd": "k_trial_fwd", "k_store_ll": "k_misc", "k_load_w": "k_misc",
        "k_trial_rev": "k_trial_rev", "k_cov_seq_rev": "k_cov_rev", "kw_cov_seq_rev": "k_cov_rev", "k_cov_contrib": "k_cov_contrib",
        "kw_cov_contrib": "k_cov_contrib", "k_kf_rev": "k_kf_rev", "k_lqr_rev": "k_lqr_rev", "k_unpack": "k_unpack",
        "k_fwd_pipe": "k_fwd_pipe", "k_rev_pipe": "k_rev_pipe"}
FP32 = {"FFMA2": 4, "FFMA": 2, "FADD": 1, "FMUL": 1, "FADD2": 2, "FMUL2": 2}      # flops per lane
FP64 = {"DFMA": 2, "DADD": 1, "DMUL": 1}


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--samples", type=int, required=True, help="parameter samples per launch in the capture")
    ap.add_argument("--trials", type=int, required=True, help="trials per sample as the kernels see them (after factorisation)")
    ap.add_argument("--T", type=int, required=True)
    ap.add_argument("--dims", type=int, nargs=5, required=True)
    ap.add_argument("-o", default=os.path.join(ROOT, "profiles", "r02_counters.json"))
    a = ap.parse_args()
    import bench
    raw = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "raw", "--csv"]))))
    hdr = raw[0]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        v = r[col[k]].replace(",", "") if k in col else ""
        try:
            return float(v)
        except ValueError:
            return 0.0

    def unit_scale(k):
        u = raw[1][col[k]] if k in col else ""
        return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)

    per = defaultdict(lambda: defaultdict(float))
    fn_of = defaultdict(set)
    for r in raw[2:]:
        name = r[col["Kernel Name"]]
        fn = re.sub(r"^void\s+", "", name).split("<")[0].split("(")[0].replace("lqgk::", "")
        kind = KIND.get(fn)
        if kind is None:
            continue
        fn_of[kind].add(fn)
        p = per[kind]
        p["launches"] += 1
        p["ms"] += num(r, "gpu__time_duration.sum") * unit_scale("gpu__time_duration.sum")
        p["dram_bytes"] += num(r, "dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum") + \
            num(r, "dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum")
        p["inst"] += num(r, "smsp__inst_executed.sum")
        w = num(r, "gpu__time_duration.sum")
        p["_w"] += w
        p["fma_pipe"] += w * num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
        p["fp64_pipe"] += w * num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
        p["issue"] += w * num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        p["regs"] = max(p["regs"], num(r, "launch__registers_per_thread"))
    kernels = {}
    for kind, p in per.items():
        ops = Counter()
        for fn in fn_of[kind]:
            src = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "source", "--csv", "--kernel-name", "regex:^" + fn + "$"]))))
            if len(src) < 3:
                continue
            shdr = src[1]
            sc = {h: i for i, h in enumerate(shdr)}
            # a report holding several results prints the listing of one function once per result: instruction counts of the
            # launches of one function are summed over its results, every SASS address counted once per result
            seen = Counter()
            nres = max(1, sum(1 for r in raw[2:] if re.sub(r"^void\s+", "", r[col["Kernel Name"]]).split("<")[0].replace("lqgk::", "") == fn))
            for r in src[2:]:
                if len(r) < len(shdr) or not r[sc["Instructions Executed"]].isdigit():
                    continue
                addr = r[sc["Address"]]
                seen[addr] += 1
                if seen[addr] > nres:
                    continue
                s = re.sub(r"^@!?U?P\w+\s+", "", r[sc["Source"]].strip())
                ops[s.split()[0].split(".")[0]] += int(r[sc["Instructions Executed"]])
        tot = sum(ops.values()) or p["inst"]
        fp32_inst = sum(ops[k] for k in FP32)
        fp32_flops = sum(ops[k] * v * 32 for k, v in FP32.items())
        fp64_flops = sum(ops[k] * v * 32 for k, v in FP64.items())
        kernels[kind] = {
            "functions": sorted(fn_of[kind]), "launches_per_step": p["launches"], "ms_under_ncu": p["ms"], "registers": p["regs"],
            "inst_per_sample_step": tot / (a.samples * a.T), "fp32_inst_share": fp32_inst / tot if tot else 0.0,
            "fp32_flops_per_sample": fp32_flops / a.samples, "fp64_flops_per_sample": fp64_flops / a.samples,
            "dram_bytes_per_sample": p["dram_bytes"] / a.samples,
            "fma_pipe_cycles_active_pct": p["fma_pipe"] / p["_w"] if p["_w"] else 0.0,
            "fp64_pipe_cycles_active_pct": p["fp64_pipe"] / p["_w"] if p["_w"] else 0.0,
            "issue_active_pct": p["issue"] / p["_w"] if p["_w"] else 0.0,
            "opcodes_per_sample_step": {k: round(v / (a.samples * a.T), 2) for k, v in ops.most_common(16)},
        }
    out = {"source_sha": bench.source_sha(), "report": os.path.basename(a.rep), "samples_per_launch": a.samples, "trials": a.trials,
           "T": a.T, "kernel_dims": a.dims, "flop_convention": "per lane: FFMA2 4, FFMA 2, FADD/FMUL 1, FADD2/FMUL2 2 (x32 lanes, masked lanes "
           "included); FP64: DFMA 2, DADD/DMUL 1", "kernels": kernels}
    json.dump(out, open(a.o, "w"), indent=1)
    print(f"wrote {a.o} (source_sha {out['source_sha']})")
    for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms_under_ncu"]):
        print(f"  {k:14s} {v['ms_under_ncu']:8.3f} ms  inst/sample-step {v['inst_per_sample_step']:7.1f}  fp32 share {v['fp32_inst_share']:.2f}  "
              f"fma pipe {v['fma_pipe_cycles_active_pct']:.1f}%  fp64 pipe {v['fp64_pipe_cycles_active_pct']:.1f}%  issue {v['issue_active_pct']:.1f}%  "
              f"DRAM/sample {v['dram_bytes_per_sample'] / 1e3:.1f} KB")


if __name__ == "__main__":
    main()
