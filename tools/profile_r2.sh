#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are copied to profiles/).
#   tools/profile_r2.sh launches   per-kernel launch lists (gpu__time_duration) of prefill and decode
#   tools/profile_r2.sh full       `ncu --set full` captures of the GEMM tiles, the fused prefill attention and the decode kernels
#   tools/profile_r2.sh sanitize   compute-sanitizer memcheck / racecheck / synccheck on the tiny-model smoke
set -u
O=gpurun_out
mkdir -p $O
case "${1:-launches}" in
launches)
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_prefill_launches.csv python tools/ncu_prefill.py llama-3.2-1b 512 > $O/r2_prefill_launches.log 2>&1
  LMRS_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 340 --csv --log-file $O/r2_decode_launches.csv python tools/ncu_decode.py llama-3.2-1b 1 512 3 > $O/r2_decode_launches.log 2>&1
  ;;
full)
  ncu --set full --clock-control none -k regex:gemm_q8_kernel -s 4 -c 4 -f -o $O/r2_gemm python tools/ncu_prefill.py llama-3.2-1b 512 > $O/r2_full_gemm.log 2>&1
  ncu --set full --clock-control none -k regex:prefill_attn_fused -s 1 -c 1 -f -o $O/r2_pfattn python tools/ncu_prefill.py llama-3.2-1b 512 > $O/r2_full_pfattn.log 2>&1
  LMRS_B200_GRAPH=0 ncu --set full --clock-control none -k regex:"lmrs_q_matvec|attn_cluster" -s 86 -c 5 -f -o $O/r2_decode python tools/ncu_decode.py llama-3.2-1b 1 512 3 > $O/r2_full_decode.log 2>&1
  for r in r2_gemm r2_pfattn r2_decode; do      # the reports are 10-40 MB each: keep the CSV pages, drop the reports (gpurun_out is capped at 64 MiB)
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
    ncu -i $O/$r.ncu-rep --page details --csv > $O/$r.details.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  done
  ;;
prefillfull)
  ncu --set full --clock-control none -k regex:gemm_q8_kernel -s 6 -c 2 -f -o $O/r2b_gemm python tools/ncu_prefill.py llama-3.2-1b 512 > $O/r2b_full_gemm.log 2>&1
  ncu --set full --clock-control none -k regex:prefill_attn_fused -s 1 -c 1 -f -o $O/r2b_pfattn python tools/ncu_prefill.py llama-3.2-1b 512 > $O/r2b_full_pfattn.log 2>&1
  for r in r2b_gemm r2b_pfattn; do
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  done
  ;;
sanitize)
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py 2 > $O/r2_sanitize_$tool.log 2>&1
    echo "exit $?" >> $O/r2_sanitize_$tool.log
  done
  ;;
esac
