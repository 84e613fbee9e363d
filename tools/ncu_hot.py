"""Top stalled SASS instructions per kernel from `ncu -i rep --page source --csv` output.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_hot.py [N]"""
import csv, sys
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; secs.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for si, s in enumerate(secs):
    h = {n: i for i, n in enumerate(s["hdr"])}
    stall_cols = [n for n in s["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[h["# Samples"]]) for r in s["rows"])
    print("=== [{}] {}  total samples {}".format(si, s["name"][:60], tot))
    agg = {n: sum(int(r[h[n]]) for r in s["rows"]) for n in stall_cols}
    print("   by reason:", ", ".join("{}={:.0f}%".format(k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    order = sorted(range(len(s["rows"])), key=lambda i: -int(s["rows"][i][h["# Samples"]]))[:N]
    for i in sorted(order):
        r = s["rows"][i]
        smp = int(r[h["# Samples"]])
        top = max(stall_cols, key=lambda n: int(r[h[n]]))
        print("  {:5d} {:5.1f}% {:14s} {}".format(i, 100 * smp / max(tot, 1), top[6:], r[h["Source"]].strip()[:90]))
