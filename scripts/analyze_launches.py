import csv, collections, re, sys
f = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines=[l for l in open(f) if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0])
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    unit=row['Metric Unit']; ns = v*1e3 if unit=='us' else (v*1e6 if unit=='ms' else v)
    short=re.sub(r'^void ','',re.sub(r'\(.*','',row['Kernel Name']))
    agg[short][0]+=1; agg[short][1]+=ns
tot=sum(v[1] for v in agg.values())
print("total %.2f ms, %d launches"%(tot/1e6, sum(v[0] for v in agg.values())))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print("%-62s n=%4d  %8.3f ms  %5.1f%%"%(k[:62],v[0],v[1]/1e6,100*v[1]/tot))
if len(sys.argv) > 2:
    rows=list(csv.DictReader(open(sys.argv[2])))
    agg=collections.OrderedDict()
    for r in rows:
        k=(int(r['M']),int(r['N']),int(r['K']),int(r['ksize']),int(r['stride']),int(r['BN']),int(r['m_tiles']),int(r['n_tiles']),int(r['stages']),int(r['grid'])*10+int(r.get('cg',1)))
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r['ms'])
    print("conv_tc total (event-timed, warm): %.2f ms"%sum(a[1] for a in agg.values()))
    print("     M     N      K ks s  BN  mt  nt st grid |  n   ms_tot  ms_each  TF/s  | ideal_us  memMB")
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[3]) if len(sys.argv)>3 else 24]:
        M,N,K=k[0],k[1],k[2]; fl=2*M*N*K
        byts=(M*K/(9 if k[3]==3 else 1)+M*N+N*K)*2/1e6
        print("%6d %5d %6d %2d %d %3d %4d %3d %2d %4d | %2d  %7.3f  %7.3f  %5.0f | %7.1f %7.1f"%(k+(a[0],a[1],a[1]/a[0],fl/(a[1]/a[0]*1e-3)/1e12, fl/1385e12*1e6, byts)))
