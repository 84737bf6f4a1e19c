cd /root/repo/dlwp_benchmark_b200/csrc
for cfg in "4 160 128" "8 96 128" "4 160 256" "8 96 256" "2 288 128" "4 192 64"; do
  set -- $cfg
  rm -f analysis_fused.o
  make -j8 BRINGUP=1 EXTRA="-DAF_KYT_V=$1 -DAF_CTHREADS_V=$2 -DAF_RTHREADS_V=$3" > /tmp/mk.log 2>&1 || { tail -5 /tmp/mk.log; continue; }
  echo "KYT=$1 CTHREADS=$2 RTHREADS=$3: $(cd /root/repo && python scripts/kb_analysis.py | tail -1)"
done
