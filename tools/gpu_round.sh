#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of every library kernel of one step.
# usage (from repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_eager.csv \
    python tools/one_step.py 3 > $OUT/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
# step 3 of 3 eager steps: skip the launches of the first two (counted from the launch list)
SKIP=$(python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("$OUT/${TAG}_launches_eager.csv") if l.startswith('"'))]
names=[r[4] for r in rows[1:]]
idx=[i for i,n in enumerate(names) if "k_mt_generate_chunks" in n or "k_mt_generate(" in n]
mine=[i for i,n in enumerate(names) if "tipb::" in n]
start=idx[-2]+1 if len(idx)>=2 else 0
print(sum(1 for i in mine if i<start))
PY
)
echo "skipping $SKIP library launches"
TIPB_DUMP_WORKLOAD=$OUT/${TAG}_workload.json timeout 1200 ncu --set full --clock-control none --import-source on \
    -k 'regex:^k_' -s $SKIP -c 80 -f -o $OUT/${TAG}_full python tools/one_step.py 3 > $OUT/${TAG}_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | tail -12
