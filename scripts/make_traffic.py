"""profiles/traffic.json from an ncu --set full capture of the sketching kernel (dram read+write bytes per launch)."""
import csv, io, json, subprocess, sys
rep, reads = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v = float(r[i]); u = units[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}[u]
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
out = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "reads_per_launch": reads,
       "kernel": r[hdr.index("Kernel Name")], "source": rep.split("/")[-1] + " (ncu --set full --clock-control none)"}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(out)
