import csv, collections, re, sys
path=sys.argv[1]; n=int(sys.argv[2]) if len(sys.argv)>2 else 185
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    val=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit=='ns': val/=1e3
    elif unit=='ms': val*=1e3
    seq.append((re.sub(r'\(.*','',row['Kernel Name']).replace('void ','').replace('ukbb::',''),val))
first=[i for i,(k,v) in enumerate(seq) if k.startswith('conv0') or k.startswith('conv_first')]
a=first[0]; b=first[1] if len(first)>1 else len(seq)
print('--- one sub-batch ---')
for k,v in seq[a:b]: print(f'{k[:58]:58s} {v:9.1f} us')
tot=collections.Counter(); cnt=collections.Counter()
for k,v in seq[:n]: tot[k]+=v; cnt[k]+=1
T=sum(tot.values()); print('--- per subject: total us',round(T,1))
for k,v in tot.most_common(): print(f'{k[:58]:58s} {v:10.1f} us {100*v/T:5.1f}% n={cnt[k]}')
