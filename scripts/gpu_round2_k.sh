# 1 GPU, release build: steady-state DRAM bytes of back-to-back launches (ncu keeps L2, one pass), GPU tests, smoke, bench
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for snake in 0 1; do
  for w in free rocket falcon9 cube_sat; do
    B200_SNAKE=$snake ncu --cache-control none --clock-control none --metrics $M -k regex:body_fast_spec --launch-skip 2 --launch-count 6 \
      --csv --log-file gpurun_out/r02_steady_${w}_snake${snake}.csv python scripts/rocket_kernel_run.py $w > /dev/null 2>&1
  done
done
python - <<'PY'
import csv, glob, collections
for f in sorted(glob.glob('gpurun_out/r02_steady_*_snake*.csv')):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    i_id, i_m, i_v = hdr.index('ID'), hdr.index('Metric Name'), hdr.index('Metric Value')
    d = collections.defaultdict(dict)
    for r in rows: d[r[i_id]][r[i_m]] = float(r[i_v].replace(',', ''))
    print(f, [(round(v.get('dram__bytes_read.sum', 0) / 1e6, 1), round(v.get('dram__bytes_write.sum', 0) / 1e6, 1), round(v.get('gpu__time_duration.sum', 0) / 1e3, 1)) for v in d.values()])
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_f.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02_bench_n1_f.json').read().strip().splitlines()[-1])
print('value', b['value'], 'frac', b['roofline']['frac'], 'e2e', b['e2e']['value'])
print({k:(round(v['frac'],3), round(v['us_per_tick'],1)) for k,v in b['effector_sets'].items()}, b.get('telemetry_every_tick'), b.get('egm08_degree_64'))
PY
