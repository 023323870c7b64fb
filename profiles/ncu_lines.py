"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
total warp-stall samples, barrier-stall share and executed instructions.
usage: python profiles/ncu_lines.py <export.csv> [top_n]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file = None
    hdr = None
    out = []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r or r[0] == "" or r[0] == "Function Name":
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr, r))

        def num(k):
            try:
                return float(d.get(k, "0").replace(",", ""))
            except ValueError:
                return 0.0

        out.append((num("Warp Stall Sampling (All Samples)"), num("stall_barrier"), num("stall_long_sb"),
                    num("stall_short_sb"), num("stall_math"), num("stall_wait"), num("Instructions Executed"),
                    cur_file, line, r[1].strip()[:90]))
    tot = sum(o[0] for o in out) or 1.0
    print(f"total samples {tot:.0f}")
    print("  %smp  barrier  long_sb short_sb   math   wait   inst_exec  file:line  source")
    for o in sorted(out, reverse=True)[:top]:
        print(f"{100*o[0]/tot:6.2f} {100*o[1]/tot:7.2f} {100*o[2]/tot:7.2f} {100*o[3]/tot:7.2f} {100*o[4]/tot:7.2f} "
              f"{100*o[5]/tot:6.2f} {o[6]:11.0f}  {o[7]}:{o[8]}  {o[9]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
