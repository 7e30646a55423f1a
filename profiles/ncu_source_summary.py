import csv,sys,collections,re
def toi(v):
    try: return int(v)
    except Exception: return 0
rows=list(csv.reader(open(sys.argv[1])))
kernel=sys.argv[2] if len(sys.argv)>2 else '(int)0'
secs=[i for i,r in enumerate(rows) if r and r[0]=='File Path']+[len(rows)]
byop=collections.Counter(); sampop=collections.Counter(); stall=collections.Counter(); byline=collections.Counter(); sampline=collections.Counter()
tot=0; tots=0
for a,b in zip(secs[:-1],secs[1:]):
    f=rows[a][1]; fn=rows[a+1][1]
    if kernel not in fn: continue
    hdr=rows[a+2]; idx={h:i for i,h in enumerate(hdr)}
    stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    line=None; src=''
    for r in rows[a+3:b]:
        if len(r)<len(hdr): continue
        if r[0]: line=r[0]; src=r[1]; continue
        if not r[2].startswith('0x'): continue
        m=re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3]); op=m.group(1) if m else '?'
        ie=toi(r[idx['Instructions Executed']]); ns=toi(r[idx['# Samples']])
        tot+=ie; tots+=ns; byop[op]+=ie; sampop[op]+=ns
        key=(f.split('/')[-1],line,src.strip()[:70]); byline[key]+=ie; sampline[key]+=ns
        for c in stall_cols: stall[c]+=toi(r[idx[c]])
print('total warp-inst',tot,'samples',tots)
for op,c in byop.most_common(30): print(f'  {op:10s} {c:11d} {100*c/tot:5.1f}%  samples {100*sampop[op]/max(tots,1):5.1f}%')
print('stalls %:', [(k[6:],round(100*v/tots,1)) for k,v in stall.most_common(14)])
print('top lines by samples:')
for k,v in sampline.most_common(40): print(f'  {100*v/tots:5.1f}% samp {100*byline[k]/tot:5.1f}% inst  {k}')
