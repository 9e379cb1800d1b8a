# gpurun (1 GPU): ncu capture of the partition form's scatter and bucket kernels (second to_csc call of the process)
O=gpurun_out/r2csx; mkdir -p $O
cat > /tmp/one_csc.py <<'PY'
import os, sys, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "tch-geometric_b200")]
import tch_geometric as thg
from tools import synth
ei, n = synth.products_like("cuda:0")
for _ in range(2):
    out = thg.to_csc(ei, n)
torch.cuda.synchronize()
PY
TCHGEO_CSX_SORT=partition timeout 150 ncu --set full --import-source on --clock-control none -k regex:'pt_scatter_kernel|pt_bucket_kernel' --launch-skip 2 -c 2 -f -o $O/r2_to_csc_partition python /tmp/one_csc.py > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu.log; ls -la $O/*.ncu-rep
