import csv,subprocess,sys,re,collections,os,tempfile
rep=sys.argv[1]; lib=sys.argv[2]
src=open('/root/repo/hybrid-drt_b200/csrc/qphb_warp.cuh').read().split('\n')
# function line ranges
starts=[]
for i,l in enumerate(src,1):
    if re.match(r'(__device__|__global__|template <int J0)',l) or l.startswith('qphb_warp_kernel'):
        nm=re.search(r'(\w+)\(',l)
        if l.startswith('template <int J0'): continue
        if nm: starts.append((i,nm.group(1)))
def fn(line):
    name='?'
    for s,n in starts:
        if s<=line: name=n
    return name
tmp=tempfile.mkdtemp()
subprocess.run(['cuobjdump','-xelf','all',lib],cwd=tmp,check=True,stdout=subprocess.DEVNULL)
line_of={}
for f in os.listdir(tmp):
    if not f.endswith('.cubin'): continue
    txt=subprocess.run(['nvdisasm','--print-line-info-inline',os.path.join(tmp,f)],capture_output=True,text=True).stdout
    infunc=False; cur=None; stack=[]
    for ln in txt.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            infunc='qphb_warp_kernel' in ln; continue
        if not infunc: continue
        m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',ln)
        if m:
            f_=os.path.basename(m.group(1)); l_=int(m.group(2))
            if 'inlined at' in m.group(3):
                # take the outermost warp.cuh location in chain: handled by following lines
                pass
            cur=(f_,l_); chain=[cur]; continue
        m=re.search(r'//## File "([^"]+)", line (\d+) inlined at',ln)
        m2=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(\S.*);',ln)
        if m2: line_of[int(m2.group(1),16)]=cur
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr=rows[1]
ia,isamp,iexec=hdr.index('Address'),hdr.index('# Samples'),hdr.index('Instructions Executed')
base=None; recs=[]
for r in rows[2:]:
    if len(r)<=isamp: continue
    a=int(r[ia],16)
    if base is None: base=a
    recs.append((a-base,int(r[isamp] or 0),int(r[iexec] or 0)))
# assign function by nearest preceding instruction whose line is in qphb_warp.cuh
agg=collections.Counter(); aggx=collections.Counter(); curfn='?'
for off,s,x in recs:
    k=line_of.get(off)
    if k and k[0]=='qphb_warp.cuh': curfn=fn(k[1])
    agg[curfn]+=s; aggx[curfn]+=x
tot=sum(agg.values()); totx=sum(aggx.values())
for k,v in agg.most_common(): print(f'{100*v/tot:6.1f}% samples  {100*aggx[k]/totx:6.1f}% instr  {aggx[k]/5920:10.0f} instr/fit  {k}')
