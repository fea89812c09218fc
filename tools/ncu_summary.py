import csv, collections, subprocess, sys
rep = sys.argv[1]; trials = float(sys.argv[2]) if len(sys.argv) > 2 else 40.96e6
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum","launch__registers_per_thread","launch__block_size","launch__grid_size","sm__warps_active.avg.pct_of_peak_sustained_active",
"smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
"sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed","smsp__thread_inst_executed_per_inst_executed.ratio",
"dram__bytes_read.sum","dram__bytes_write.sum","sm__cycles_elapsed.avg","sm__inst_executed.avg.per_cycle_active","smsp__cycles_active.avg"]
for w in want:
    for i,h in enumerate(hdr):
        if h == w: print("%-75s %-10s %s" % (w, units[i], [r[i] for r in data]))
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; idx = {h:i for i,h in enumerate(hdr)}
blocks=[]
for r in rows:
    if r and r[0]=="Kernel Name": blocks.append([]); continue
    if r and r[0]=="Address": continue
    if blocks: blocks[-1].append(r)
b = blocks[0]
op = collections.Counter(); stall = collections.Counter(); samples=0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in b:
    n = int(r[idx["Instructions Executed"]]); src_ = r[idx["Source"]].split()
    o = src_[0] if not src_[0].startswith("@") else src_[1]
    op[o.split(".")[0]] += n
    for c in stall_cols: stall[c] += int(r[idx[c]])
    samples += int(r[idx["# Samples"]])
wt = trials/32
print("warp-instructions per warp-trial by opcode:")
print("  " + "  ".join("%s %.1f" % (o, n/wt) for o,n in op.most_common(22)))
print("  total %.1f" % (sum(op.values())/wt))
print("stall samples: " + "  ".join("%s %.1f%%" % (c[6:], 100*n/samples) for c,n in stall.most_common(9)))
