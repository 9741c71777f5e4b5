import csv,sys
f=sys.argv[1]
rows=list(csv.reader(open(f)))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
def num(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(num(r,'# Samples') for r in data)
texec=sum(num(r,'Instructions Executed') for r in data)
print('total samples',tot,'inst executed',texec)
# top 40 by samples
top=sorted(data,key=lambda r:-num(r,'# Samples'))[:int(sys.argv[2]) if len(sys.argv)>2 else 40]
for r in sorted(top,key=lambda r:int(r[ix['Address']],16) if r[ix['Address']].startswith('0x') else int(r[ix['Address']])):
    st={k:num(r,k) for k in ['stall_long_sb','stall_short_sb','stall_wait','stall_math','stall_barrier','stall_not_selected','stall_selected','stall_mio','stall_lg','stall_branch_resolving','stall_dispatch','stall_no_inst']}
    s=' '.join(f"{k[6:]}={int(v)}" for k,v in st.items() if v>0.02*num(r,'# Samples'))
    print(r[ix['Address']][-5:], f"{100*num(r,'# Samples')/tot:5.2f}%", f"{r[ix['Source']][:60]:60s}", s)
