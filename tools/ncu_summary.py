"""Summarises an `ncu --set full` report of tools/ncu_kernels.py (one launch per shipped kernel and shape, L2 flushed):
    python tools/ncu_summary.py gpurun_out/r2a_kernels.ncu-rep profiles/r2_ncu_kernels.txt [profiles/ncu_traffic.json]
Per launch: duration, DRAM bytes read + written (bench.py's `roofline.traffic` source), achieved DRAM GB/s, tensor-pipe / XU /
issue utilisation, registers, occupancy.  The launch order is the order of tools/ncu_kernels.py, which labels the shapes."""
import csv
import json
import subprocess
import sys

LABELS = [  # (kernel substring, label, traffic key or None) in launch order of tools/ncu_kernels.py attn,attn3840,gn,geglu,ln,publish
    ("fmha_fwd", "attention 1024^2 level 1  b2 Lq4096 Lkv4096 h10 d64", "fmha_fwd_kernel b2 lq4096 lkv4096 h10 d64"),
    ("fmha_fwd", "attention 1024^2 level 2  b2 Lq1024 Lkv1024 h20 d64", "fmha_fwd_kernel b2 lq1024 lkv1024 h20 d64"),
    ("fmha_fwd", "attention 3840^2 n=4 level 2  b1 Lq3600 Lkv14400 h20 d64", "fmha_fwd_kernel b1 lq3600 lkv14400 h20 d64"),
    ("gn_stats", "GroupNorm stats  b2 C320 128x128", None), ("gn_apply", "GroupNorm apply+SiLU  b2 C320 128x128", None),
    ("gn_stats", "GroupNorm stats  b2 C640 64x64", None), ("gn_apply", "GroupNorm apply+SiLU  b2 C640 64x64", None),
    ("gn_stats", "GroupNorm stats  b2 C1280 32x32", None), ("gn_apply", "GroupNorm apply+SiLU  b2 C1280 32x32", None),
    ("gn_fused", "GroupNorm fused (stats+apply+SiLU)  b2 C320 128x128", None),
    ("gn_fused", "GroupNorm fused (stats+apply+SiLU)  b2 C640 64x64", None),
    ("gn_fused", "GroupNorm fused (stats+apply+SiLU)  b2 C1280 32x32", None),
    ("geglu", "GEGLU gate  rows 8192 cols 2560", None), ("geglu", "GEGLU gate  rows 2048 cols 5120", None),
    ("add_layernorm", "add+LayerNorm  rows 8192 C640", None), ("add_layernorm", "add+LayerNorm  rows 2048 C1280", None),
    ("bias_residual", "conv bias + residual  b2 C320 128x128", None), ("bias_residual", "conv bias + residual  b2 C1280 32x32", None),
    ("publish", "K|V publication 10.5 MB to one peer (loopback)", None),
    ("linear_kernel", "GEMM + GEGLU  M2048 K1280 D5120 (level-2 FF1)", None),
    ("linear_kernel", "GEMM bias+residual  M2048 N1280 K5120 (level-2 FF2)", None),
    ("linear_kernel", "GEMM plain  M2048 N3840 K1280 (level-2 q|k|v)", None),
]
M = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)


def main(rep, out_txt, out_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(v) for k, v in M.items()}
    kn, gs, bs = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    lines = [f"# ncu --set full --clock-control none, one launch per kernel and shape, L2 flushed before the launch ({rep})",
             f"# {'kernel / shape':62s} {'us':>8s} {'DRAM rd MB':>10s} {'wr MB':>8s} {'GB/s':>7s} {'dram%':>6s} {'tensor%':>7s} {'XU%':>5s} {'issue%':>6s} {'warps%':>6s} {'regs':>4s}  grid x block"]
    traffic = {}
    li = 0
    for r in data:
        name = r[kn]
        label, key = name[:60], None
        while li < len(LABELS) and LABELS[li][0] not in name:
            li += 1
        if li < len(LABELS):
            label, key = LABELS[li][1], LABELS[li][2]
            li += 1
        us = to_us(r[col["t"]], units[col["t"]])
        rd, wr = to_bytes(r[col["rd"]], units[col["rd"]]), to_bytes(r[col["wr"]], units[col["wr"]])
        f = lambda k: float(r[col[k]].replace(",", ""))
        lines.append(f"  {label:62s} {us:8.1f} {rd / 1e6:10.2f} {wr / 1e6:8.2f} {(rd + wr) / us / 1e3:7.0f} {f('dram_pct'):6.1f} {f('tensor'):7.1f} "
                     f"{f('xu'):5.1f} {f('issue'):6.1f} {f('warps'):6.1f} {int(f('regs')):4d}  {r[gs]} x {r[bs]}")
        if key:
            traffic[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "duration_us_under_ncu": us,
                            "tensor_pipe_pct": f("tensor"), "source": rep.split("/")[-1]}
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if out_json:
        json.dump(traffic, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
