import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
# first kernel only
k=0; out=[]
hdr=None
for r in rows:
    if r and r[0]=='Kernel Name':
        k+=1
        if k>1: break
        continue
    if r and r[0]=='Address': hdr=r; continue
    if hdr and len(r)>=len(hdr)-2: out.append(r)
si=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
tot=sum(int(r[si]) for r in out)
print('total samples',tot,'instr rows',len(out))
# print program regions with cumulative samples: top 45 lines by samples
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
top=sorted(range(len(out)),key=lambda i:-int(out[i][si]))[:int(sys.argv[2]) if len(sys.argv)>2 else 40]
for i in sorted(top):
    r=out[i]
    st=sorted(((int(r[c]),hdr[c]) for c in stall_cols),reverse=True)[:2]
    print(i, r[1].strip()[:60].ljust(60), r[si], r[ie], st)
# cumulative by region: print running sum every 50 instr
acc=0
for i,r in enumerate(out):
    acc+=int(r[si])
    if i%40==39: print('upto',i,'cum %.1f%%'%(100*acc/tot))
print('--- executed warp-instr per warp by region')
ex=[int(r[ie]) for r in out]
import itertools
for a in range(0,len(out),100):
    print(a, '%.1f'%(sum(ex[a:a+100])/32768.0), 'samples %.1f%%'%(100*sum(int(r[si]) for r in out[a:a+100])/tot))
