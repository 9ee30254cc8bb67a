import csv,sys,re
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
pat=re.compile(sys.argv[1])
for i,k in enumerate(h):
    if pat.search(k): print(k, r[i])
