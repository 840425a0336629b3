import csv,re,collections,sys
lines=[l for l in open(sys.argv[1]) if l.startswith('"')]
d=collections.OrderedDict()
for r in csv.DictReader(lines):
    k=int(r['ID']); d.setdefault(k,{'name':re.sub(r'.*rrqr_blocked_kernel','',r['Kernel Name'])[:14],'grid':int(r['Grid Size'].strip('()').split(',')[0])})
    d[k][r['Metric Name']]=float(r['Metric Value'].replace(',',''))
items=list(d.items()); half=len(items)//2
agg=collections.defaultdict(lambda:[0,0.0,0.0,0])
tot=0
for i,(k,v) in enumerate(items[half:]):
    t=v['gpu__time_duration.sum']/1e3
    key=v['name']
    agg[key][0]+=1; agg[key][1]+=t; agg[key][2]+=v['smsp__inst_executed.sum']/1e6; agg[key][3]+=v['grid']
    tot+=t
    if len(sys.argv)>2: print(k, v['name'], v['grid'], 'smem', int(v['launch__shared_mem_per_block_dynamic']), 't(us)', round(t,1), 'Minst', round(v['smsp__inst_executed.sum']/1e6,1))
print("total ms", tot/1e3)
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): print(k, 'launches',v[0],'ms',round(v[1]/1e3,2),'Ginst',round(v[2]/1e3,2),'CTAs',v[3])
