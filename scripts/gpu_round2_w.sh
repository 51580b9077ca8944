# 1 GPU, final code: ncu --set full of the default free-body FAST kernel and of the EXACT kernel, then tests, smoke, bench
for w in free; do
  ncu --set full --clock-control none --import-source on -k regex:body_fast_spec --launch-skip 6 --launch-count 1 -o gpurun_out/r02_final_$w python scripts/rocket_kernel_run.py $w > /dev/null 2>&1
  ncu -i gpurun_out/r02_final_$w.ncu-rep --page details --csv > gpurun_out/r02_final_${w}_ncu_details.csv 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:body_exact --launch-skip 3 --launch-count 1 -o gpurun_out/r02_final_exact python scripts/exact_kernel_run.py > /dev/null 2>&1
ncu -i gpurun_out/r02_final_exact.ncu-rep --page details --csv > gpurun_out/r02_final_exact_ncu_details.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_n1_h.json 2> gpurun_out/r02_bench_n1_h.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_h.err
python - <<'PY'
import json, csv
b=json.loads(open('gpurun_out/r02_bench_n1_h.json').read().strip().splitlines()[-1])
print('value', b['value'], 'frac', b['roofline']['frac'], 'e2e', b['e2e']['value'], 'exact', b['exact_math']['value'], 'launches', b['gpu_launches'], b['clocks'])
for f in ('gpurun_out/r02_final_free_ncu_details.csv','gpurun_out/r02_final_exact_ncu_details.csv'):
    rows=list(csv.reader(open(f))); h=rows[0]; i_n=h.index('Metric Name'); i_v=h.index('Metric Value')
    print(f, {r[i_n]: r[i_v] for r in rows[1:] if r[i_n] in ('Duration','DRAM Throughput','Memory Throughput','Compute (SM) Throughput','Registers Per Thread','Achieved Occupancy','Executed Ipc Active')})
PY
