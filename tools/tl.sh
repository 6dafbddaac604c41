# development: timeline build + run, then normal rebuild + quick checks
cd /root/repo; rm -f spurfies_b200/csrc/mlp_tc2.o; SPF_TIMELINE=1 python -c "
from spurfies_b200.build import build_library; build_library(force=False)" > /dev/null 2>&1; python tools/timeline_sdf.py > gpurun_out/timeline.log 2>&1; head -4 gpurun_out/timeline.log
rm -f spurfies_b200/csrc/mlp_tc2.o; python -c "
from spurfies_b200.build import build_library; build_library(force=False)" > /dev/null 2>&1
bash tools/gpu_quick.sh ${1:-g2c}
