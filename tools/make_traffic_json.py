"""profiles/roofline_traffic.json from a `ncu --set full` capture of the flat kernels: DRAM bytes per launch of k_flat_words / k_flat_rows
and the hash of the CUDA sources the capture belongs to (bench.py reports `roofline.traffic` only when the hash matches its build).
    python tools/make_traffic_json.py capture.ncu-rep [raw.csv out] [note]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def col(r, name):
    i = hdr.index(name)
    return float(r[i]) * scale.get(units[i], 1)
out = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    for k in ("k_flat_words", "k_flat_rows"):
        if k in name:
            rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
            out[k] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "duration_us": float(r[hdr.index("gpu__time_duration.sum")])}
note = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none --cache-control none: dram__bytes_read.sum + dram__bytes_write.sum per launch, one launch = one 1,048,576-pair chunk"
json.dump({"source_hash": bench.source_hash(), "note": note, "kernels": out}, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out))
