#!/usr/bin/env python
"""Condense an .ncu-rep into the handful of metrics the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv, subprocess, sys, io

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_op_dmma.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        # 5th-generation tensor cores (tcgen05): pipe activity, bf16 -> fp32 operations, tensor memory traffic
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.per_second",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    name_i = hdr.index("Kernel Name")
    for d in data:
        out.append(f"== {d[name_i][:70]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"  {k:88s} {d[i]:>18s} {units[i]}")
        st = [(float(d[i].replace(',', '')), h[len(STALL):].replace('_per_issue_active.ratio', ''))
              for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("per_issue_active.ratio")]
        out.append("  warp stall reasons (warps per issue slot): " +
                   ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")

main()
