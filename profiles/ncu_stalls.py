"""Top stalled SASS instructions of an .ncu-rep (needs -lineinfo + --import-source on): python profiles/ncu_stalls.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
# extra arguments select a kernel, e.g. --kernel-name regex:dkv --launch-skip 0 --launch-count 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sys.argv[3:], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(ends) > 1:
    rows = rows[:ends[1]]
print(rows[0][1][:120])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, ' instructions', len(data))
agg = {}
for r in data:
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h and r[ix[h]] not in ('', '0'):
            agg[h] = agg.get(h, 0) + int(r[ix[h]])
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
for i, r in sorted(enumerate(data), key=lambda kv: -int(kv[1][ix['# Samples']]))[:topn]:
    stalls = {h[6:]: int(r[ix[h]]) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h and r[ix[h]] not in ('', '0')}
    print(str(i).rjust(5), r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(8), r[ix['Source']].strip()[:64].ljust(64),
          dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:3]))
