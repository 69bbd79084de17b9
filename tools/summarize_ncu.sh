#!/bin/bash
# Turn a .ncu-rep brought back in gpurun_out/ into the text summaries we commit under profiles/.
# usage: tools/summarize_ncu.sh gpurun_out/s1/fused_sgd.ncu-rep profiles/fused_sgd
REP=$1; OUT=$2
mkdir -p "$(dirname "$OUT")"
ncu -i "$REP" --page raw --csv > "$OUT.raw.csv" 2>/dev/null
python - "$OUT.raw.csv" "$OUT.summary.txt" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
if len(rows) < 3:
    open(sys.argv[2], "w").write("empty report\n"); sys.exit(0)
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "launch__occupancy_limit_registers"]
idx = {k: hdr.index(k) for k in keys if k in hdr}
with open(sys.argv[2], "w") as f:
    for r in rows[2:]:
        f.write("----\n")
        for k, i in idx.items():
            f.write(f"{k:70s} {r[i]} {units[i]}\n")
PY
ncu -i "$REP" --page source --csv > "$OUT.source.csv" 2>/dev/null
head -c 400000 "$OUT.source.csv" > "$OUT.source.head.csv"; rm -f "$OUT.source.csv"
echo "wrote $OUT.summary.txt"
