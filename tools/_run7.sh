mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; head -c 400 gpurun_out/r1e_bench.json; echo
LMRS_B200_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 340 --csv --log-file gpurun_out/r1e_decode_launches.csv python tools/ncu_decode.py llama-3.2-1b 1 512 3 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
LMRS_B200_GRAPH=0 timeout 300 ncu --set full --clock-control none -k regex:attn_cluster -s 17 -c 2 -o /tmp/attc_full python tools/ncu_decode.py llama-3.2-1b 1 512 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 120 ncu -i /tmp/attc_full.ncu-rep --page raw --csv > gpurun_out/r1e_attn_cluster.raw.csv 2>/dev/null; ls -la gpurun_out | tail -8
