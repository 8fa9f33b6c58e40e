cd $GRAFT_REPO_ROOT; export DSAC_SKIP_BUILD=1; mkdir -p gpurun_out
NB=1 REPS=3 WRITE_DM=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c16_n1.csv python tools/prof_driver.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c16_n1.csv')) if len(r)>10 and r[0].isdigit()]
n=len(rows)//3
for r in rows[-n:]: print("%-40s grid %-14s block %-12s %8.1f us"%(r[4].split('(')[0][:40], r[7], r[8], float(r[-1])/1e3))
PY
