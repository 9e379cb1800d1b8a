# gpurun (1 GPU): per-kernel timeline of to_csc in its partition form
O=gpurun_out/r2csx; mkdir -p $O
TCHGEO_CSX_SORT=partition timeout 100 python tools/profile_timeline.py to_csc > $O/to_csc_partition_timeline.txt 2> $O/timeline.err; echo "rc=$?"; cat $O/to_csc_partition_timeline.txt | cut -c1-110
