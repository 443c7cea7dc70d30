"""Summarise `ncu --page source --csv --print-source sass,cuda` output: top source lines by samples."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = list(csv.reader(open(path, errors='replace')))
i = 0
sections = []
while i < len(rows):
    r = rows[i]
    if r and r[0] == 'File Path':
        fpath = r[1]; func = rows[i+1][1]; hdr = rows[i+2]; j = i+3; body = []
        while j < len(rows) and rows[j] and rows[j][0] not in ('File Path', 'Function Name') and len(rows[j]) == len(hdr):
            body.append(rows[j]); j += 1
        sections.append((fpath, func, hdr, body)); i = j
    else:
        i += 1
for fpath, func, hdr, body in sections:
    if not (fpath.endswith('.cu') or fpath.endswith('.cuh')): continue
    ci = {h: k for k, h in enumerate(hdr)}
    def col(r, name):
        try: return float(r[ci[name]].replace(',', ''))
        except Exception: return 0.0
    tot = sum(col(r, '# Samples') for r in body)
    if tot == 0: continue
    print(f"== {func[:60]} :: {fpath.split('/')[-1]} samples={tot:.0f}")
    body=[r for r in body if r[0].strip()]
    body.sort(key=lambda r: -col(r, '# Samples'))
    for r in body[:topn]:
        s = col(r, '# Samples')
        print(f"  L{r[0]:>4} {s/tot*100:5.1f}%  long_sb={col(r,'stall_long_sb'):.0f} short_sb={col(r,'stall_short_sb'):.0f} barrier={col(r,'stall_barrier'):.0f} mio={col(r,'stall_mio'):.0f} wait={col(r,'stall_wait'):.0f} inst={col(r,'Instructions Executed'):.0f} | {r[1].strip()[:90]}")
