import csv,re,sys,collections,os
fn=sys.argv[1]  # mangled function name
sass_csv=sys.argv[2]
lines=open(os.environ.get('NTM_SASS','gpurun_out/elf/all.sass')).read().split('\n')
# locate function text section
start=None
for i,l in enumerate(lines):
    if l.strip().startswith('.section') and ('.text.'+fn) in l: start=i; break
assert start is not None
cur=('?',0); instrs=[]
for l in lines[start+1:]:
    if l.strip().startswith('.section'): break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: instrs.append((int(m.group(1),16),cur,m.group(2)))
rows=list(csv.reader(open(sass_csv)))
hdr=rows[1]; data=[]
for r in rows[2:]:
    if r and r[0]=='Kernel Name': break
    data.append(r)
ia=hdr.index('Address'); ie=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); ismp=hdr.index('# Samples')
base=int(data[0][ia],16)
byoff={o:(c,t) for o,c,t in instrs}
agg=collections.Counter(); smp=collections.Counter(); tot=0; tots=0
opagg=collections.Counter()
for r in data:
    off=int(r[ia],16)-base; n=int(r[ie]); s=int(r[ismp])
    c,t=byoff.get(off,(('?',0),''))
    agg[c]+=n; smp[c]+=s; tot+=n; tots+=s
    op=r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
    opagg[op.split('.')[0]]+=n
print('total inst',tot,'samples',tots)
print('--- by opcode'); 
for k,v in opagg.most_common(25): print(f'{k:12s} {v/tot:6.3f}')
print('--- by source line')
for (f,l),v in agg.most_common(60): print(f'{f}:{l:4d} inst={v/tot:6.3f} samples={smp[(f,l)]/tots:6.3f}')
# by function region (device.cuh line ranges)
print('--- by region')
# by function: line -> nearest preceding definition at column 0 carrying __device__/__global__ in the current sources
import os
SRC=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'mpc-ntm-control_b200','csrc')
starts={}
for f in ('ntm_device.cuh','ntm_kernels.cu'):
    lst=[]
    for n,l in enumerate(open(os.path.join(SRC,f)),1):
        if ('__device__' in l or '__global__' in l) and '(' in l and not l.startswith((' ','\t','//')):
            m=re.search(r'(\w+)\s*\(',l[l.index('__'):].replace('__launch_bounds__',''))
            if m: lst.append((n,m.group(1)))
        elif re.match(r'^(\w[\w\s\*&:<>]*\s)?(\w+)\(.*[,{]\s*$',l) and lst and n-lst[-1][0]<=2 and lst[-1][1] in ('void','__launch_bounds__'):
            lst[-1]=(lst[-1][0],re.match(r'^(\w[\w\s\*&:<>]*\s)?(\w+)\(',l).group(2))
    for n,l in enumerate(open(os.path.join(SRC,f)),1):
        if l.startswith('struct Group'): lst.append((n,'Group primitives'))
        if l.startswith('struct Work'): lst.append((n,'(work area)'))
    starts[f]=sorted(lst)
def region(f,l):
    if f in starts:
        name='(top)'
        for n,nm in starts[f]:
            if n<=l: name=nm
            else: break
        return f.split('.')[0][4:]+':'+name
    return f
ra=collections.Counter(); rs=collections.Counter()
for k,v in agg.items(): ra[region(*k)]+=v
for k,v in smp.items(): rs[region(*k)]+=v
for k,v in ra.most_common(): print(f'{k:28s} inst={v/tot:6.3f} ({v/13107200:8.1f} warp-inst/inner-iter) samples={rs[k]/tots:6.3f}')
