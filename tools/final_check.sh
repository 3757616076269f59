#!/bin/bash
# End-of-round check on the GPU box: parity suite, smoke, the reference arm and the three bench lines, then the
# Blokus captures (launch list + full captures of the legal and step kernels).
tag=${1:-final}
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
python bench.py --impl reference > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
python bench.py > $out/bench_default_$tag.json 2> $out/bench_default_$tag.err
python bench.py --workload ttt4 > $out/bench_ttt4_$tag.json 2>/dev/null
python bench.py --workload blokus > $out/bench_blokus_$tag.json 2>/dev/null
python - <<PY
import json
for n in ("ref","default","ttt4","blokus"):
    d=json.load(open("$out/bench_%s_$tag.json"%n))
    r=d.get("roofline") or {}
    print(n, round(d["value"]/1e6,1), d["ms_per_step"], r.get("frac"), (r.get("single_chain") or {}).get("ms_per_step"), round(d["e2e"]["value"]/1e6,1), (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
PY
python bench.py --workload blokus --steps 24 --warmup 3 --no-cpu > $out/plain_blokus_$tag.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_blokus_$tag.csv \
    python bench.py --workload blokus --steps 24 --warmup 3 --no-cpu > $out/ncu_launches_blokus_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blokus_legal --launch-skip 60 --launch-count 1 -o $out/prof_blokus_legal_$tag -f \
    python tools/blokus_probe.py 30 > $out/ncu_full_blokus_legal_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blokus_step -s 20 -c 1 -o $out/prof_blokus_step_$tag -f \
    python bench.py --workload blokus --steps 24 --warmup 3 --no-cpu > $out/ncu_full_blokus_step_$tag.log 2>&1
ls $out/*_$tag* | wc -l
