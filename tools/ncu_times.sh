#!/bin/bash
# per-kernel device time of a few steady-state frames: tools/ncu_times.sh <regex> <skip> <count> <cmd...>
re=$1; skip=$2; cnt=$3; shift 3
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"$re" -s $skip -c $cnt --csv --log-file /tmp/ncu_times.csv "$@" > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("/tmp/ncu_times.csv")) if len(r) > 5]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print("%-34s %-55s %s" % (d["Kernel Name"][:34], d["Metric Name"], d["Metric Value"]))
PY
