"""Key metrics + top warp-stall reasons per captured launch from `ncu --page raw --csv` exports (tools/profile_r2.sh full).
usage: python tools/ncu_summary.py gpurun_out/r2_gemm.raw.csv [...]"""
import csv, sys
KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "imma%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
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
        st = [(float(r[i].replace(",", "")), h) for h, i in ix.items() if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct") and r[i] not in ("", "n/a")]
        st.sort(reverse=True)
        print("    stalls:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_warp')[0]} {v:.0f}%" for v, h in st[:6]))
