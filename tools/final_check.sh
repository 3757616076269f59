#!/bin/bash
# End-of-round check on the GPU box (gpurun -- bash tools/final_check.sh): build, parity suite, smoke, both bench arms at the
# driver's flags.  Profiles are refreshed separately by tools/r2_profile.sh + tools/r2_summarise.py.
tag=${1:-final}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "reference arm rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python tools/show_bench.py $out/bench_$tag.json $out/bench_ref_$tag.json
