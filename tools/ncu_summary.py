"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + hottest SASS lines.  usage: ncu_summary.py rep [out.md]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum"]
out = ["# ncu summary of " + rep, "", "| metric | unit | value |", "|---|---|---|"]
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        out.append(f"| {h} | {u} | {v} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S, I, SRC = ci["# Samples"], ci["Instructions Executed"], ci["Source"]
tot = sum(int(r[S]) for r in data)
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[SRC])
    o = m.group(2).split(".")[0] if m else "?"
    op[o] += int(r[I]); ops[o] += int(r[S])
out += ["", f"warp-level instructions executed by opcode (top 16): {op.most_common(16)}", "",
        f"stall samples by opcode (total {tot}): {ops.most_common(12)}", "", "hottest SASS lines (samples, executed, instruction, top stall reasons):", ""]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -int(r[S]))[:25]:
    st = sorted(((int(r[ci[h]]), h) for h in stall_cols), reverse=True)[:2]
    out.append(f"- {r[S]:>6} {r[I]:>9}  `{r[SRC].strip()[:80]}`  {st}")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
