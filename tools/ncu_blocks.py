import csv,sys,subprocess
rep=sys.argv[1]; minfrac=float(sys.argv[2]) if len(sys.argv)>2 else 0.006
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]
ia=hdr.index('Address'); isrc=hdr.index('Source'); ismp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
data=[]
for r in rows[2:]:
    try: data.append((int(r[ia],16), r[isrc].strip(), int(r[ismp]), int(r[iex])))
    except: pass
base=data[0][0]
tot_ex=sum(d[3] for d in data); tot_s=sum(d[2] for d in data)
print('total warp instr', tot_ex, 'samples', tot_s)
prev=None; start=None; n=0; ops=[]; smp=0
def flush():
    if prev is not None and prev*n>minfrac*tot_ex:
        print(f"{start:#7x} x{n:3d} instr  exec/instr {prev:8d} ({prev*n/tot_ex*100:4.1f}% ex, {smp/tot_s*100:4.1f}% smp) : {' '.join(ops[:16])}")
for a,s,sm,ex in data:
    if ex!=prev:
        flush(); prev=ex; start=a-base; n=0; ops=[]; smp=0
    n+=1; smp+=sm; ops.append((s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0])
flush()
