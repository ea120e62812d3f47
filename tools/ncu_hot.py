"""Top SASS instructions of an `ncu --page source --csv` dump by stall samples and by
executed count.  Usage: python tools/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]  # two kernels in one report: skip the second header
ci = {n: i for i, n in enumerate(hdr)}
S, X = ci["# Samples"], ci["Instructions Executed"]
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot_s = sum(int(r[S] or 0) for r in body)
tot_x = sum(int(r[X] or 0) for r in body)
print(f"total samples {tot_s}, warp-instructions {tot_x}, sass lines {len(body)}")
def show(idx):
    r = body[idx]
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    return f"{idx:5d} {r[1].strip()[:70]:70s} smp={int(r[S] or 0):6d} ({100*int(r[S] or 0)/max(tot_s,1):4.1f}%) exe={int(r[X] or 0):9d} " + " ".join(f"{n}:{v}" for v, n in st if v)
print("--- by samples")
for i in sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:N]:
    print(show(i))
print("--- by executed")
for i in sorted(range(len(body)), key=lambda i: -int(body[i][X] or 0))[:N]:
    print(show(i))
