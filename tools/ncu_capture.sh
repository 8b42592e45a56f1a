#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU):  tools/ncu_capture.sh <tag> [dram] [full] [c3dram]
#   dram   : every launch of ONE C2 build with its DRAM bytes and duration (whole-step traffic + launch list)
#   c3dram : the same for ONE C3 build (long)
#   full   : `--set full` capture of the first (= largest, level 0) launch of each kernel in KERNELS
tag=${1:-r02}; shift
mkdir -p gpurun_out
KERNELS=${KERNELS:-"k_sa_chained k_wave_rounds k_rs_scatter k_rs_hist k_merge_rounds k_cut_find k_build_clusters_warp k_update_quadrics k_write_pages_warp k_partition_chained k_pivot_large k_rank"}
for what in "$@"; do
  case $what in
    dram)
      ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${tag}_c2_ncu_launches.csv python tools/one_build.py C2 1 > gpurun_out/${tag}_c2_ncu_launches.log 2>&1
      gzip -f gpurun_out/${tag}_c2_ncu_launches.csv ;;
    c3dram)
      ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${tag}_c3_ncu_launches.csv python tools/one_build.py C3 1 > gpurun_out/${tag}_c3_ncu_launches.log 2>&1
      gzip -f gpurun_out/${tag}_c3_ncu_launches.csv ;;
    full)
      for k in $KERNELS; do
        ncu --set full --clock-control none --import-source on -k regex:^${k} -c 1 -f -o gpurun_out/${tag}_${k}_full python tools/one_build.py C2 1 > gpurun_out/${tag}_${k}_full.log 2>&1
      done ;;
  esac
done
ls -la gpurun_out | tail -30
