"""Per-kernel summary of an `ncu --set full` report (read with `ncu -i rep --page raw --csv`).
usage: ncu -i full.ncu-rep --page raw --csv | python experiments/ncu_summary.py"""
import csv, re, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("lts__t_sector_hit_rate.pct", "L2hit%"), ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%"),
     ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma%"),
     ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wf%"),
     ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc%"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu%"),
     ("sm__inst_executed.sum.per_cycle_elapsed", "ipc_gpu"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
     ("launch__registers_per_thread", "regs"), ("launch__block_size", "block"), ("launch__grid_size", "grid"),
     ("launch__shared_mem_per_block_dynamic", "dyn_smem")]
cols = [(ix[[h for h in hdr if h.endswith(k) or h == k][0]], n) for k, n in M if any(h.endswith(k) or h == k for h in hdr)]
print("%-44s " % "kernel" + " ".join("%11s" % n for _, n in cols))
print("%-44s " % "(units)" + " ".join("%11s" % units[i][:11] for i, _ in cols))
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("ukbb::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
    vals = []
    for i, _ in cols:
        v = r[i].replace(",", "")
        try:
            vals.append("%11.2f" % float(v))
        except ValueError:
            vals.append("%11s" % v[:11])
    print("%-44s " % name[:44] + " ".join(vals))
