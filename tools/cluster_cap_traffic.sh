for c in 74 70 66 60 50; do
  echo "== max clusters $c"
  CPB200_DEBUG_MAX_CLUSTERS=$c python tools/time_models.py 1048576 cmb_TT_NN 2>&1 | grep rows=
  CPB200_DEBUG_MAX_CLUSTERS=$c ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_mlp -s 3 -c 1 --csv python tools/time_models.py 1048576 cmb_TT_NN 2>/dev/null | grep -E "dram__|lts__" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
