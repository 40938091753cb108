"""Print the tcgen05 / mbarrier hand-off latencies measured by grl_tc_latency_probe (SM clock cycles).
usage: python tools/tc_latency_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geometry_rl_b200 import _lib as L  # noqa: E402

NAMES = ["1 MMA (M128 N64 K16) + commit -> wait, issuing thread", "4 MMAs + commit -> wait", "8 MMAs + commit -> wait",
         "16 MMAs + commit -> wait", "2 x st.shared.v4 + fence.proxy.async", "mbarrier.arrive -> try_wait in another warp",
         "tcgen05.ld 32x32b.x16 + wait::ld", "round trip: arrive -> MMA warp issues 4 MMAs + commit -> wait",
         "round trip with st.shared + fence.proxy.async + tcgen05 fences before the arrive",
         "16 unrolled MMAs N64, A K-major B K-major", "16 unrolled MMAs N64, A K-major B MN-major",
         "16 unrolled MMAs N64, A MN-major B MN-major", "16 unrolled MMAs N80, A MN-major B MN-major",
         "8 unrolled MMAs N128, A K-major B K-major"]
out = torch.zeros(64, dtype=torch.int64, device="cuda")
for _ in range(2):
    L.call("grl_tc_latency_probe", L.ptr(out))
torch.cuda.synchronize()
o = out.cpu().tolist()
res = {n: {"min_cycles": o[2 * i], "mean_cycles": o[2 * i + 1]} for i, n in enumerate(NAMES)}
print(json.dumps(res, indent=1))
