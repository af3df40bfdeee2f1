"""Top SASS instructions by stall samples from an .ncu-rep.  Usage: python tools/ncu_hot.py rep kernel-regex [N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; out = []
for r in rows:
    if r and r[0] == 'Address': h = r; continue
    if h is None or len(r) < len(h): continue
    try: s = int(r[h.index('# Samples')]); e = int(r[h.index('Instructions Executed')])
    except ValueError: continue
    out.append((s, e, r[0], r[h.index('Source')]))
tot = sum(o[0] for o in out)
print("total samples", tot, "columns:", [c for c in h if 'stall' in c.lower()][:3])
for i, (s, e, addr, text) in enumerate(out):
    out[i] = (s, e, addr, text, i)
for s, e, addr, text, i in sorted(out, key=lambda o: -o[0])[:N]:
    print(f"{100*s/tot:5.1f}%  samples {s:6d}  exec {e:9d}  #{i:5d} {addr[-5:]}  {text[:90]}")
