set -x
python bench.py > gpurun_out/bench_r02_g.json 2> gpurun_out/bench_r02_g.err
python tools/shard_time.py > gpurun_out/shard_r02_g.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_shard_r02_g.csv python tools/shard_time.py once > gpurun_out/ncu_shard.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gb_fourier_stage2|gb_stage1_tab|gb_pack" -s 3 -c 3 -f -o gpurun_out/synth240_r02_g python tools/shard_once.py 240 3 > gpurun_out/ncu240.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gb_fourier_stage2|gb_stage1_tab|gb_pack" -s 3 -c 3 -f -o gpurun_out/synth30_r02_g python tools/shard_once.py 30 3 > gpurun_out/ncu30.log 2>&1
tail -c 600 gpurun_out/bench_r02_g.json | head -c 300; tail -2 gpurun_out/shard_r02_g.log | head -1
