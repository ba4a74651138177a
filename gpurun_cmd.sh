T=${TAG:-a}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_$T.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$T.log
SCGPU_LIB=$PWD/sc-gameengine_b200/libscgpu_checked.so timeout 300 python -m pytest tests/test_gpu_churn.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_${T}_checked.log 2>&1; echo "checked rc=$?"; tail -2 gpurun_out/t_${T}_checked.log
for C in 25; do
SCGPU_CHURN_COHORTS=$C timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_resolve_parents|k_build_windows|k_scan_tiles|k_flatten_windows|k_update_win$" -c 60 --csv --log-file gpurun_out/topo_$C.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-partial --churn-frames 6 > gpurun_out/topo_$C.log 2>&1; echo "ncu rc=$?"
done
python - <<PY
import csv,io,collections,re
for C in (25,):
    t=open('gpurun_out/topo_%d.csv'%C,errors='replace').read(); t=t[t.index('"ID"'):]
    by=collections.defaultdict(list)
    for r in csv.DictReader(io.StringIO(t)):
        by[re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","")].append(float(r["Metric Value"])/1e3)
    print('cohorts',C,{k:[round(x,1) for x in v[-6:]] for k,v in by.items()})
PY
