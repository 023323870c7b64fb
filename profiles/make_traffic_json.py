"""profiles/ncu_traffic.json from `ncu -i X.ncu-rep --page raw --csv` exports of the round's `ncu --set full` captures.
Every entry carries the DRAM bytes of ONE launch at the bench's shape and a hash of the kernel's source files at capture
time; bench.py recomputes the hash and reports `traffic_stale: true` when the sources have changed since.

  python profiles/make_traffic_json.py gpurun_out/r2_kernels_raw.csv [more.csv ...]
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scalable-meta-learning-with-gaussian-processes_b200", "csrc")
COMMON = ["scaml_device.cuh"]
# bench key -> (kernel-name substring, template filter, launch description, source files)
KERNELS = {
    "scaml_fit_kernel<RBF>": ("scaml_fit_kernel<0>", None, "config3: 4096 tasks x R=6 x n=256 x d=6", ["scaml_fit.cuh"]),
    # the same kernel at config 4's shape (same grid: 3 CTAs per SM at both shapes): told apart by the launch order of
    # scripts/ncu_driver.py -- config 3, the factorize-mode launch (< 10 ms), then config 4
    "scaml_fit_kernel<RBF> config4": ("scaml_fit_kernel<0>", "second", "config4 block: 2048 tasks x R=2 x n=512 x d=10",
                                      ["scaml_fit.cuh"]),
    "scaml_predict_kernel<RBF>": ("scaml_predict_kernel<0, 64, 0>", None, "4096 GPs x 18944 candidates",
                                  ["scaml_predict.cuh"]),
    "scaml_predict_kernel<RBF,64,CROSS>": ("scaml_predict_kernel<0, 64, 1>", None,
                                           "4096 GPs x 18944 candidates, n_t = 32", ["scaml_predict.cuh"]),
    "scaml_cond_prepare_kernel<RBF>": ("scaml_cond_prepare_kernel<0, 8>", None, "4096 GPs, 64-column panel (candidates)",
                                       ["scaml_cond.cuh", "scaml_predict.cuh"]),
    "scaml_kmat_kernel<RBF>": ("scaml_kmat_task_kernel<0>", None, "4096 tasks x 256 x 256", ["scaml_kmat.cuh"]),
}


def sha16(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(os.path.join(CSRC, p), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(paths):
    out = {"source": "ncu --set full --clock-control none, one launch each at the bench's shapes (scripts/ncu_driver.py; "
                     "summaries profiles/r2_*_ncu_summary.txt)"}
    fit_long = {}  # (capture, launch id) -> index among the long fit-kernel launches of that capture
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], dict(zip(rows[0], rows[1]))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            for key, (sub, _, launch, srcs) in KERNELS.items():
                if sub not in d["Kernel Name"]:
                    continue
                try:
                    rd = float(d["dram__bytes_read.sum"]) * scale(units["dram__bytes_read.sum"])
                    wr = float(d["dram__bytes_write.sum"]) * scale(units["dram__bytes_write.sum"])
                    ms = float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}[units["gpu__time_duration.sum"]]
                except (ValueError, KeyError):
                    continue
                if rd != rd or wr != wr:
                    continue  # pass not collected (nan)
                if key == "scaml_fit_kernel<RBF>" and ms < 10.0:
                    continue  # the factorize-mode launch of the same kernel
                if sub == "scaml_fit_kernel<0>":
                    if ms < 10.0:
                        continue  # the factorize-mode launch
                    nth = fit_long.setdefault((path, d["ID"]), len({k for k in fit_long if k[0] == path}))
                    if (key == "scaml_fit_kernel<RBF>") != (nth == 0):
                        continue  # first long launch of a capture = config 3, second = config 4
                files = COMMON + srcs
                out[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "duration_ms_under_ncu": ms, "launch": launch,
                            "sources": files, "source_sha16": sha16(files)}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        if isinstance(v, dict):
            print(f"{k:40s} read {v['dram_bytes_read'] / 1e9:8.3f} GB  write {v['dram_bytes_write'] / 1e9:8.3f} GB  "
                  f"{v['duration_ms_under_ncu']:9.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1:])
