cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "1 5 1" "1 10 1" "2 10 1"; do set -- $cfg
NWCWT_STREAMS=$1 NWCWT_RING_MB=$2 NWCWT_NO_L2_PERSIST=$3 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:nwcwt_pass.2_ -s 100 -c 40 --csv --log-file gpurun_out/dram_$1_$2_$3.csv python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/ncu_d.log 2>&1
python - <<PY
import csv,collections
lines=[l for l in open('gpurun_out/dram_$1_$2_$3.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in csv.DictReader(lines):
    k='A' if 'passA2' in r['Kernel Name'] else 'B'
    agg[k][r['Metric Name']].append(float(r['Metric Value'].replace(',',''))*{'Gbyte':1e3,'Mbyte':1,'Kbyte':1e-3,'byte':1e-6}.get(r['Metric Unit'],1))
for k,v in agg.items():
    print("streams $1 ring $2", k, {m.replace('lts__t_sectors_srcunit_tex_op_','').replace('.sum','').replace('dram__bytes_','dram_'):round(sum(x)/len(x),2) for m,x in v.items()}, 'n', len(v['gpu__time_duration.sum']))
PY
done
