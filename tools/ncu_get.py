import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; vals=rows[2:]
pats=sys.argv[2:]
for i,h in enumerate(hdr):
    if any(p in h for p in pats):
        print(f"{h:95s} {units[i]:14s} {[v[i] for v in vals]}")
