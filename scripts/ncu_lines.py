"""Attribute the warp-stall samples of an ncu report to source lines of skinny.cu.
usage: ncu -i X.ncu-rep --page source --csv > /tmp/src.csv; cuobjdump -xelf all libmdbn_b200.so; nvdisasm -g skinny.sm_100a.cubin > /tmp/xelf/sk_lines.txt;
       python scripts/ncu_lines.py ILi10E 40   (kernel instantiation tag, rows to print)"""
import re,csv,collections,sys
tag=sys.argv[1]
line=None; off2line={}; infn=False
for l in open('/tmp/xelf/sk_lines.txt'):
    if l.lstrip().startswith('.section') and '.text.' in l:
        infn = tag in l
    if not infn: continue
    m=re.search(r'//## File "([^"]*)", line (\d+)',l)
    if m:
        line=int(m.group(2)) if m.group(1).endswith('skinny.cu') else -int(m.group(2)); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,6})\*/',l)
    if m: off2line[int(m.group(1),16)]=line
rows=list(csv.reader(open('/tmp/src.csv')))
hdr=rows[1]; data=rows[2:]
base=int(data[0][0],16)
si=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
cols=['stall_barrier','stall_long_sb','stall_short_sb','stall_wait','stall_no_inst','stall_not_selected','stall_selected','stall_math','stall_mio','stall_branch_resolving','stall_lg','stall_membar','stall_dispatch']
ci=[hdr.index(c) for c in cols]
agg=collections.defaultdict(lambda:[0,0]+[0]*len(cols))
for r in data:
    if r and r[0] in ('Kernel Name','Address'): 
        if r[0]=='Kernel Name': break
        continue
    if len(r)<len(hdr): continue
    ln=off2line.get(int(r[0],16)-base,-99999)
    a=agg[ln]; a[0]+=int(r[si]); a[1]+=int(r[ie])
    for k,i in enumerate(ci): a[2+k]+=int(r[i])
tot=sum(a[0] for a in agg.values()); print('total samples',tot)
src=open('/root/repo/mdbn_b200/csrc/skinny.cu').read().split('\n')
for ln,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[2])]:
    s=src[ln-1].strip()[:70] if ln>0 else '?'
    print(f"{ln:5d} {a[0]:5d} {100*a[0]/tot:5.1f}% inst={a[1]:8d} "+' '.join(f"{c[6:10]}={v}" for c,v in zip(cols,a[2:]) if v>0.1*a[0])+' | '+s)
