cd $GRAFT_REPO_ROOT
timeout 600 python - <<'PY' 2> gpurun_out/enc_bench.err
import argparse, json, sys, torch
sys.path.insert(0, ".")
import bench
from geometry_rl_b200 import ops
args = argparse.Namespace(no_graph=False, no_dp_graph=False, warmup=3, repeats=3, steps=20)
dev = torch.device("cuda")
name = "rigid_insertion_two_agents_multi_transformer_trpl_cfg"
for B in (1000, 8192):
    r = bench.measure(args, name, B, "bf16", dev, None, 0, 1, 0, 20, full=False)
    orig = ops.encoder_layer_supported
    ops.encoder_layer_supported = lambda x, l: False
    r2 = bench.measure(args, name, B, "bf16", dev, None, 0, 1, 0, 20, full=False)
    ops.encoder_layer_supported = orig
    print(json.dumps({"B": B, "kernel_ms": r["ms_per_step"], "kernel_samples_s": r["value"], "library_ms": r2["ms_per_step"], "library_samples_s": r2["value"]}))
PY
tail -3 gpurun_out/enc_bench.err
