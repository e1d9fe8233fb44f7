#!/usr/bin/env python
"""Per-region / per-opcode instruction totals of an ncu SASS export (see tools_sass_lines.py)."""
import csv,re,collections,sys
src,dis,kname=sys.argv[1:4]
rows=list(csv.reader(open(src))); hdr=rows[1]
ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
insts=[(r[1].strip(), float(r[ie] or 0), float(r[isamp] or 0)) for r in rows[2:] if len(r)>ie]
lines=open(dis).read().split('\n')
start=next(i for i,l in enumerate(lines) if l.startswith('.text.'+kname+':'))
cur=('?',0); seq=[]
for l in lines[start+1:]:
    if (l.startswith('//-----') or l.startswith('.text.')) and seq: break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l): seq.append(cur)
srcl=open('/root/repo/icm_slam_b200/csrc/fused.cuh').read().split('\n')
marks=[(i+1,l) for i,l in enumerate(srcl) if '// ----' in l or l.startswith('__device__') or l.startswith('__global__') or l.startswith('template')]
def region(f,ln):
    if f!='fused.cuh': return f
    name='top'
    for i,l in marks:
        if i<=ln: name='%d:%s'%(i,l.strip()[:60])
    return name
reg=collections.OrderedDict(); opc=collections.Counter(); tot=0; tots=0
for (txt,n,s),(f,ln) in zip(insts,seq):
    k=region(f,ln); a=reg.setdefault(k,[0,0]); a[0]+=n; a[1]+=s; tot+=n; tots+=s
    opc[txt.split()[0] if not txt.startswith('@') else txt.split()[1]]+=n
print('total %.1f M warp-inst'%(tot/1e6))
for k,v in reg.items(): print('%-72s %6.1f M  %5.1f%% inst %5.1f%% samples'%(k,v[0]/1e6,100*v[0]/tot,100*v[1]/tots))
print()
print(' '.join('%s:%.1f%%'%(k,100*v/tot) for k,v in opc.most_common(24)))
