import csv,re,sys,collections
fn=sys.argv[1]  # mangled function name
sass_csv=sys.argv[2]
lines=open('/tmp/elf/all.sass').read().split('\n')
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
def region(f,l):
    if f=='ntm_device.cuh':
        if 501<=l<=555: return 'build_GF_toeplitz'
        if 555<l<=620: return 'build_GF_dense'
        if 327<=l<=440: return 'qp_solve'
        if 269<=l<=305: return 'ldl_solve'
        if 452<=l<=490: return 'aff scan'
        if 50<=l<=95: return 'schedule'
        if 95<l<=215: return 'group prims'
        return 'device other'
    if f=='ntm_kernels.cu':
        if 150<=l<=175: return 'rollout'
        return 'run_scenario other'
    return f
ra=collections.Counter(); rs=collections.Counter()
for k,v in agg.items(): ra[region(*k)]+=v
for k,v in smp.items(): rs[region(*k)]+=v
for k,v in ra.most_common(): print(f'{k:28s} inst={v/tot:6.3f} ({v/13107200:8.1f} warp-inst/inner-iter) samples={rs[k]/tots:6.3f}')
