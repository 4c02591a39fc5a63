for B in 148 111 99 74; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:bootstrap_kernel_v4 -c 1 --csv --log-file gpurun_out/r2_traffic_b$B.csv python bench.py --n 1024 --batch $B --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
  echo "batch $B"; grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/r2_traffic_b$B.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
