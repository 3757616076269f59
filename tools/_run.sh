python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tron.csv python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tron_step -s 60 -c 2 -o gpurun_out/prof_tron_r1b python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blokus_legal -s 40 -c 1 -o gpurun_out/prof_blokus_legal_r1b python bench.py --workload blokus --steps 24 --warmup 3 --no-cpu > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blokus_step -s 40 -c 1 -o gpurun_out/prof_blokus_step_r1b python bench.py --workload blokus --steps 24 --warmup 3 --no-cpu > gpurun_out/ncu5.log 2>&1
tail -c 300 gpurun_out/ncu2.log
