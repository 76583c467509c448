# tools/ncu_kernels.sh -- one `ncu --set full` capture per big kernel of the aggregate-verify step (run on the GPU box):
#   gpurun -- 'bash tools/ncu_kernels.sh'   -> gpurun_out/r2_<kernel>.ncu-rep ; summarise here with tools/ncu_summary.py
for k in k_hash_to_g2 k_g2_decode k_rlc_scale k_mask_aggregate_serial; do
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:^$k -c 1 -o gpurun_out/r2_$k python tools/profile_target.py 303104 1 > gpurun_out/r2_ncu_$k.log 2>&1
done
HBLS_COOP_MAX=100000 HBLS_RLC_MIN=1000000 timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_pairing_coop -c 1 -o gpurun_out/r2_k_pairing_coop python tools/profile_target.py 1332 1 > gpurun_out/r2_ncu_coop.log 2>&1
