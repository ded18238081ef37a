TAG=v7y
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/${TAG}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 57 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_shares.txt
cat gpurun_out/${TAG}_tests.txt; tail -1 gpurun_out/${TAG}_smoke.txt; tail -c 300 gpurun_out/${TAG}_bench.json; echo; head -14 gpurun_out/${TAG}_launch_shares.txt
