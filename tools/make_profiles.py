"""CPU side: turn the .ncu-rep captures a gpurun call brought back into the committed evidence under profiles/:
one text summary per kernel (tools/ncu_summary.py) and the per-launch DRAM traffic table bench.py reads
(profiles/traffic.json).
    python tools/make_profiles.py r2_final gpurun_out/r2f_gen_c5.ncu-rep:c5 gpurun_out/r2f_gen_c4.ncu-rep:c4 ..."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
outdir = os.path.join(ROOT, "profiles", tag)
os.makedirs(outdir, exist_ok=True)
tpath = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
detail = traffic.setdefault("_detail", {})
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}
for spec in sys.argv[2:]:
    rep, wl = spec.split(":")
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "500"], capture_output=True, text=True).stdout
    name = "%s_%s_ncu_summary.txt" % (wl, os.path.splitext(os.path.basename(rep))[0])
    open(os.path.join(outdir, name), "w").write(txt)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))

    def val(key):
        return float(d[key].replace(",", "")) * UNIT.get(u[key], 1.0)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    traffic[wl] = rd + wr
    detail[wl] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr, "kernel": d.get("Kernel Name", ""),
                  "duration_ns": val("gpu__time_duration.sum"), "report": os.path.basename(rep), "round": tag}
    print(wl, name, "dram %.3f GB" % ((rd + wr) / 1e9), "duration %.3f ms" % (val("gpu__time_duration.sum") / 1e6))
traffic["_note"] = ("per launch of the dominant kernel, from ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum); "
                    "see _detail[workload].round / .report for the capture each entry comes from. Output that is still dirty in "
                    "the 126 MB L2 when the kernel ends is not counted by ncu.")
json.dump(traffic, open(tpath, "w"), indent=1)
