"""Key metrics + top warp-stall reasons per captured launch from `ncu --page raw --csv` exports (tools/profile_r2.sh full).
usage: python tools/ncu_summary.py gpurun_out/r2_gemm.raw.csv [...]"""
import csv, sys
KEYS = [("gpu__time_duration.sum", "time"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefronts%"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "imma%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%"),
        ("smsp__inst_executed.sum", "inst"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("smsp__cycles_active.avg", "cycles")]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("==", f)
    for r in data:
        name = r[ix["Kernel Name"]][:70]
        out = []
        for k, lab in KEYS:
            if k in ix and r[ix[k]] != "":
                out.append(f"{lab}={r[ix[k]]}{units[ix[k]] if units[ix[k]] not in ('', '%') else ''}")
        print(name); print("   ", "  ".join(out))
        st = []
        for h, i in ix.items():   # warp-state sampling: share of the samples per stall reason
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i] not in ("", "n/a"):
                st.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        print("    warp states:", ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in st[:8]))
