set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_j.log 2>&1; tail -2 gpurun_out/t_j.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_j.log 2>&1; tail -1 gpurun_out/smoke_j.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -c 200 gpurun_out/bench_j.json
python bench.py --steps 5 --warmup 3 --no-net --no-cpu-baseline --no-config5 --no-selfplay-e2e > gpurun_out/b_j.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_j.csv python bench.py --steps 5 --warmup 3 --no-net --no-cpu-baseline --no-config5 --no-selfplay-e2e > gpurun_out/ncu_lj.log 2>&1
python tools/kbench.py --selfplay 10 --check 0 > gpurun_out/k_j.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_play_e0 -s 3 -c 1 -o gpurun_out/prof_r02h -f python tools/kbench.py --selfplay 10 --check 0 > gpurun_out/ncu_j.log 2>&1; tail -1 gpurun_out/k_j.log; tail -1 gpurun_out/ncu_j.log
export GMZ_LIB=$PWD/datou_gomoku_muzero_b200/libgmz_verify.so
rm -f gpurun_out/soak_final.jsonl
(timeout 400 python tools/soak_parity.py --batches 25 --divs 0 --accum float64 | tail -1 >> gpurun_out/soak_final.jsonl) 2>> gpurun_out/soak_final.err
(timeout 400 python tools/soak_parity.py --batches 25 --divs 0,16 --accum float32 | tail -1 >> gpurun_out/soak_final.jsonl) 2>> gpurun_out/soak_final.err
(timeout 400 python tools/soak_parity.py --batches 13 --divs 0,16 --accum float32 --mode MuZero | tail -1 >> gpurun_out/soak_final.jsonl) 2>> gpurun_out/soak_final.err
(timeout 400 python tools/soak_parity.py --batches 12 --divs 2,4,16 --accum float64 | tail -1 >> gpurun_out/soak_final.jsonl) 2>> gpurun_out/soak_final.err
(timeout 400 python tools/soak_parity.py --batches 4 --divs 2,0 --accum float64 --sims 1600 --games 1024 | tail -1 >> gpurun_out/soak_final.jsonl) 2>> gpurun_out/soak_final.err
(timeout 400 python tools/soak_parity.py --batches 4 --divs 0,16 --accum float32 --board 19 --top 32 --sims 200 --games 2048 | tail -1 >> gpurun_out/soak_final.jsonl) 2>> gpurun_out/soak_final.err
cut -c1-400 gpurun_out/soak_final.jsonl
# optional (6 min): the production-mode soak, 1 024 000 searches -> last line of profiles/r02_soak_parity.jsonl
if [ -n "$GMZ_LONG_SOAK" ]; then
(timeout 1200 python tools/soak_parity.py --batches 250 --divs 0,0,16 --accum float32 | tail -1 >> gpurun_out/soak_final.jsonl) 2>> gpurun_out/soak_final.err
fi
