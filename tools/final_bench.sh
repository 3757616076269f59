python bench.py --impl reference > gpurun_out/bench_ref_r1t.json 2> gpurun_out/bench_ref_r1t.err
python bench.py > gpurun_out/bench_default_r1t.json 2> gpurun_out/bench_default_r1t.err
python bench.py --workload ttt4 > gpurun_out/bench_ttt4_r1t.json 2>/dev/null
python bench.py --workload blokus > gpurun_out/bench_blokus_r1t.json 2>/dev/null
python - <<EOF
import json
for n in ("ref","default","ttt4","blokus"):
    d=json.load(open("gpurun_out/bench_%s_r1t.json"%n))
    r=d.get("roofline") or {}
    print(n, round(d["value"]/1e6,1), d["ms_per_step"], r.get("frac"), (r.get("single_chain") or {}).get("ms_per_step"), round(d["e2e"]["value"]/1e6,1), (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
EOF
