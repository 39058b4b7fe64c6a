#!/usr/bin/env bash
# Regenerates profiles/traffic.json (read by bench.py for roofline.traffic) on a B200 box:
#   gpurun --timeout 900 -- 'bash scripts/capture_traffic.sh r2'
# 1. the launch list of a short bench run (gpu__time_duration.sum, no clock control) -> gpurun_out/<tag>_launches.csv
# 2. one `ncu --set full` capture of two k_persistent_fused launches of the same command -> gpurun_out/<tag>_full.ncu-rep,
#    its raw page as CSV -> gpurun_out/<tag>_ncu_raw.csv
# 3. gpurun_out/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch and per pivot, stamped with
#    the hash of the kernel sources it was taken from (bench.py reports `traffic: null` when the sources changed since).
# Copy the outputs you want judged into profiles/.
set -u
TAG="${1:-r2}"
OUT=gpurun_out
mkdir -p "$OUT"
CMD="python bench.py --steps 1 --warmup 1 --pivots-per-step 128 --chunk 64 --no-full-solve --no-cpu-baseline --no-extras"
$CMD > "$OUT/${TAG}_prof_plain.json" 2> "$OUT/${TAG}_prof_plain.err" || { echo "bench failed without ncu"; tail -5 "$OUT/${TAG}_prof_plain.err"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches.csv" $CMD > "$OUT/${TAG}_ncu_list.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_persistent_fused -s 1 -c 2 -f -o "$OUT/${TAG}_full" $CMD > "$OUT/${TAG}_ncu_full.log" 2>&1
ncu -i "$OUT/${TAG}_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_ncu_raw.csv" 2> /dev/null
python - "$OUT/${TAG}_ncu_raw.csv" "$OUT/traffic.json" "$TAG" <<'PY'
import csv, hashlib, json, os, sys
raw, out, tag = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
units = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
def col(name):
    i = hdr.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0,
             "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(units[i], 1.0)
    return [float(r[i].replace(",", "")) * scale for r in data]
rd, wr, ms = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
n = len(rd)
root = os.path.dirname(os.path.dirname(os.path.abspath(raw)))
h = hashlib.sha256()
for f in ("sb200_fused.cuh", "sb200_persistent.cuh", "sb200_state.cuh", "sb200_phases.cuh"):
    h.update(open(os.path.join(root, "simplex_b200", "csrc", f), "rb").read())
piv = 64
t = {"source": "scripts/capture_traffic.sh %s: ncu --set full, %d launches of k_persistent_fused, %d pivots each, m=4096 n=16384" % (tag, n, piv),
     "kernel_source_sha256": h.hexdigest(),
     "dram_bytes_read_per_launch": sum(rd) / n, "dram_bytes_write_per_launch": sum(wr) / n, "pivots_per_launch": piv,
     "dram_bytes_per_pivot": (sum(rd) + sum(wr)) / n / piv, "kernel_ms_per_launch": sum(ms) / n,
     "dram_TBs": (sum(rd) + sum(wr)) / n / (sum(ms) / n * 1e-3) / 1e12, "model_bytes_per_pivot": 8 * (4096 * 16384 + 2 * 4096 * 4096)}
json.dump(t, open(out, "w"), indent=1)
print(json.dumps(t))
PY
