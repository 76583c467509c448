"""tools/ncu_summary.py RAW.csv SOURCE.csv TITLE -- text summary of one ncu --set full capture (key metrics, stall reasons,
dynamic opcode mix) from `ncu -i X.ncu-rep --page raw --csv` and `--page source --csv`."""
import csv, collections, re, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__stack_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]

def main():
    raw, src, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = [title, "(ncu --set full --clock-control none; numbers under the profiler are not bench values)", ""]
    for k in WANT:
        if k in d: out.append(f"{k:75s} {d[k][0]:>20s} {d[k][1]}")
    out += ["", "stall reasons (warps per issue-active cycle):"]
    st = [(float(v.replace(",", "")), h) for h, v in zip(hdr, vals)
          if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    for v, h in sorted(st, reverse=True)[:10]:
        out.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:6.3f}")
    rows = list(csv.reader(open(src))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    op = collections.Counter(); tot = 0
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        if not m: continue
        parts = m.group(2).split("."); key = parts[0]
        if key == "IMAD": key = "IMAD." + (parts[1] if len(parts) > 1 else "")
        n = int(r[ix["Instructions Executed"]] or 0); op[key] += n; tot += n
    out += ["", f"dynamic opcode mix (warp instructions executed, total {tot:.3e}):"]
    for k, v in op.most_common(12): out.append(f"  {k:12s} {v / tot * 100:6.2f} %")
    out.append(f"\nIMAD.WIDE thread-instructions = {op['IMAD.WIDE'] * 32:.3e}")
    print("\n".join(out))

if __name__ == "__main__":
    main()
