#!/usr/bin/env python
"""Group the source-page samples of one kernel by 'Instructions Executed' (= loop nest) and print
stall reasons per group.  usage: ncu_regions.py REP KERNEL_REGEX [top]"""
import csv, subprocess, io, collections, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kre],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=rows[1]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls=[(j,h) for j,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
seen=set(); order=[]
for r in rows[2:]:
    if len(r)<=max(i_s,i_ex) or r[0] in seen or r[0]=='Address': continue
    seen.add(r[0])
    try: int(r[i_s])
    except ValueError: continue
    order.append(r)
tot=sum(int(r[i_s]) for r in order)
groups=collections.OrderedDict()
for idx,r in enumerate(order):
    g=groups.setdefault(r[i_ex],{'n':0,'samples':0,'first':idx,'last':idx,'st':collections.Counter(),'ops':collections.Counter()})
    g['n']+=1; g['samples']+=int(r[i_s]); g['last']=idx
    t=r[i_src].split()
    g['ops'][t[1] if t[0].startswith('@') else t[0]]+=1
    for j,h in stalls:
        if r[j] not in('','0'): g['st'][h[6:]]+=int(r[j])
for k,g in sorted(groups.items(), key=lambda kv:-kv[1]['samples'])[:top]:
    print(f"ex={k:>8s} n={g['n']:5d} samples={g['samples']:5d} {100*g['samples']/tot:5.1f}% idx={g['first']}-{g['last']} stalls={g['st'].most_common(4)} ops={g['ops'].most_common(4)}")
print('total', tot)
