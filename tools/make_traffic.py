"""profiles/dram_traffic.json from ncu --set full captures (gpurun_out/prof_r02_*.ncu-rep), stamped with the hash of the
sources the captured library was built from (bench.py ignores a file whose stamp differs from the sources next to it).
usage: python tools/make_traffic.py stage=report.ncu-rep[:kernel-substring] ..."""
import csv
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "semantic-gaussians_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
out = {"_src_sha256_16": b.source_hash(),
       "_source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r02_*.txt)"}
for arg in sys.argv[1:]:
    stage, rest = arg.split("=", 1)
    path, _, want = rest.partition(":")
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if want and want not in d.get("Kernel Name", ""):
            continue
        u = dict(zip(hdr, units))

        def gb(name):
            v = float(d[name].replace(",", ""))
            unit = u[name].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        out[stage] = int(gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"))
        break
json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
