"""Summarise where a warp-specialised kernel waits: reads `ncu -i REP --page source --csv` and prints the
mbarrier spin sites (SYNCS.PHASECHK.TRANS64.TRYWAIT + its back-edge), their share of the stall samples, and the
instruction mix of everything else.  usage: ncu_waits.py report.ncu-rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[k - 1][1][:100])
hdr, data = rows[k], rows[k + 1:]
iS, iX, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
n = lambda r, i: int(r[i] or 0)
tot = sum(n(r, iS) for r in data)
print("samples", tot, " warp-instructions", sum(n(r, iX) for r in data))
for i, r in enumerate(data):
    if "TRYWAIT" in r[iSrc]:
        s = n(r, iS) + (n(data[i + 1], iS) if i + 1 < len(data) else 0) + (n(data[i - 1], iS) if "YIELD" in data[i - 1][iSrc] else 0)
        if s * 50 >= tot:
            print("  wait %-58s samples %5d (%4.1f%%)  polls %8d" % (r[iSrc].split("TRYWAIT")[1].strip()[:58], s, 100.0 * s / tot, n(r, iX)))
mix = collections.Counter()
for r in data:
    src = r[iSrc].split()
    if not src: continue
    op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
    mix[op.split(".")[0]] += n(r, iX)
print("  mix:", ", ".join("%s %d" % kv for kv in mix.most_common(14)))
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in data:
    for i in stalls: agg[hdr[i]] += n(r, i)
print("  stalls:", ", ".join("%s %d" % kv for kv in agg.most_common(8)))
