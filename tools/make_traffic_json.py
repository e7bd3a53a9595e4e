"""profiles/traffic.json from ncu reports: dram bytes (read + write) of one rollout launch per problem size.
    python tools/make_traffic_json.py rep1.ncu-rep:envs:K [rep2.ncu-rep:envs:K ...]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
recs = []
for arg in sys.argv[1:]:
    rep, envs, K = arg.split(":")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, d = rows[0], rows[1], rows[2]
    def val(name):
        i = hdr.index(name)
        v, u = float(d[i]), units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    recs.append({"envs": int(envs), "K": int(K), "kernel": d[hdr.index("Kernel Name")],
                 "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                 "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                 "gpu_time_ms_under_ncu": float(d[hdr.index("gpu__time_duration.sum")]) *
                 {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].lower(), 1.0),
                 "report": os.path.basename(rep)})
json.dump({"launches": recs, "how": "ncu --set full --clock-control none, one launch each; dram__bytes_read.sum + dram__bytes_write.sum"},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(recs, indent=1))
