#!/bin/bash
# One gpurun call: `ncu --set full` over scripts/ncu_driver.py (all hot kernels, one launch each) + the kmat kernel alone,
# exported on the box as small text files (the .ncu-rep files stay in /tmp: gpurun_out/ is limited to 64 MiB).
#   gpurun --timeout 1800 -- 'bash scripts/ncu_capture.sh TAG'
TAG=${1:-x}; O=gpurun_out/ncu_$TAG; mkdir -p $O
SCAML_NCU_ONCE=1 timeout 1000 ncu --set full --clock-control none --import-source on \
  -k regex:"scaml_(fit|fit8|predict|kmat|cond_prepare)" -f -o /tmp/kernels python scripts/ncu_driver.py > $O/driver.log 2>&1
tail -1 $O/driver.log
SCAML_NCU_ONLY=kmat SCAML_NCU_ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:scaml_kmat -s 2 -c 1 -f -o /tmp/kmat python scripts/ncu_driver.py > $O/driver_kmat.log 2>&1
tail -1 $O/driver_kmat.log
ncu -i /tmp/kernels.ncu-rep --page raw --csv > $O/kernels_raw.csv 2>/dev/null
ncu -i /tmp/kmat.ncu-rep --page raw --csv > $O/kmat_raw.csv 2>/dev/null
python - "$O" <<'PY'
import csv, subprocess, sys
O = sys.argv[1]
rows = list(csv.reader(open(f"{O}/kernels_raw.csv")))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    i, name = int(r[ix["ID"]]), r[ix["Kernel Name"]]
    tag = name.split("(")[0].replace("void ", "").replace("scaml::", "").replace("f8::", "").replace(" ", "")
    tag = tag.replace("<", "_").replace(">", "").replace(",", "_")
    src = f"/tmp/src_{i}.csv"
    with open(src, "w") as f:
        subprocess.run(["ncu", "-i", "/tmp/kernels.ncu-rep", "--page", "source", "--csv", "--print-source", "cuda,sass",
                        "--launch-skip", str(i), "--launch-count", "1"], stdout=f, stderr=subprocess.DEVNULL)
    out = subprocess.run([sys.executable, "profiles/ncu_lines.py", src, "40"], capture_output=True, text=True).stdout
    open(f"{O}/{i}_{tag}_hotspots.txt", "w").write(out)
    print(i, tag, out.splitlines()[0] if out else "EMPTY")
src = "/tmp/src_kmat.csv"
with open(src, "w") as f:
    subprocess.run(["ncu", "-i", "/tmp/kmat.ncu-rep", "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=f,
                   stderr=subprocess.DEVNULL)
out = subprocess.run([sys.executable, "profiles/ncu_lines.py", src, "40"], capture_output=True, text=True).stdout
open(f"{O}/kmat_hotspots.txt", "w").write(out)
PY
ls -la $O
