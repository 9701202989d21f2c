"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py rep.ncu-rep [top] [function-name substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
def I(x):
    try: return int(x)
    except Exception: return 0
agg = {}
take, si, ii = want is None, None, None
for r in rows:
    if r and r[0] == "Function Name":
        take = want is None or want in r[1]
    elif r and r[0] == "Line No":
        si = r.index("# Samples") if "# Samples" in r else None
        ii = r.index("Instructions Executed")
    elif take and r and r[0].isdigit() and ii is not None and len(r) > ii:
        a = agg.setdefault(int(r[0]), [r[1], 0, 0]); a[1] += I(r[si]) if si is not None else 0; a[2] += I(r[ii])
tot = sum(a[1] for a in agg.values()) or 1; ti = sum(a[2] for a in agg.values()) or 1
print("samples", tot, "warp-instructions", ti)
for ln, (src, s, i) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{ln:4d} {100*s/tot:5.1f}% smp {100*i/ti:5.1f}% inst  {src.strip()[:120]}")
