set -x
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log; tail -1 gpurun_out/r02_bench_n1.log
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.log; cat gpurun_out/r02_bench_ref.json | cut -c1-600
ITERS=3 FRAMES=5 TOP=30 python tools/kernel_timeline.py gpurun_out/r02_timeline_insitu.json > gpurun_out/r02_timeline_insitu.txt 2>&1; grep -E "^==" gpurun_out/r02_timeline_insitu.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_iter_bs3_plus_5frames.csv python tools/profile_iter.py 2>&1 | tail -2
ncu --profile-from-start off --set full --clock-control none --import-source on -o /tmp/top python tools/prof_final.py 2>&1 | tail -1
ncu -i /tmp/top.ncu-rep --page raw --csv > gpurun_out/r02_top_kernels_raw.csv 2>/dev/null; wc -c gpurun_out/r02_top_kernels_raw.csv
